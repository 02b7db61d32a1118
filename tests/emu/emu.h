// emu.h -- minimal cooperative-fiber SIMT emulator (TEST INFRASTRUCTURE ONLY).
//
// Runs the 32 lanes of one warp as 32 user-space fibers on one host thread. Every warp collective
// of ba_warp.cuh is built on exchange(): each lane deposits a 32-bit value, parks until all 32
// have arrived, then reads what it needs. This lets tests run the *unmodified* device source of
// block_aligner_b200/csrc/ba_kernel.cuh on a machine without a GPU and compare it bit for bit with
// the CPU oracle. Nothing in the product library references this file.
#pragma once
#include <stdint.h>

namespace emu {
int lane();
const uint32_t* exchange(uint32_t v);
void run_warp(void (*fn)(void*), void* arg);
}
