// ba_launch.cu -- the __global__ entry point of the alignment kernel family and its dispatch.
//
// ba_align_kernel<SCORING, FLAGS, FM> is instantiated for every scoring kind x {TRACE, X_DROP, EXT} x fast-phase
// mode (80 kernels). Compiled as one translation unit that takes almost four minutes, so the product build compiles
// this file once per (scoring kind, fast-phase group) with -DBA_TU_S=<kind> -DBA_TU_G=<group> in parallel
// (__graft_entry__.build) and links the objects; each unit defines ba_launch_tu_<kind>_<group> /
// ba_occupancy_tu_<kind>_<group> for its share. With -DBA_TU_TOP the file defines only the dispatch over those
// functions. Without either macro (the BA_EMU test build, tools/build_variant.sh) everything lives in one unit.
#include <stdint.h>
#include <stdio.h>
#include <vector>

#include "../../include/block_aligner_b200.h"
#include "ba_dev.h"
#ifndef BA_TU_TOP
#include "ba_kernel.cuh"
using namespace ba;
#endif

#if !defined(BA_TU_S) && !defined(BA_TU_TOP)
#define BA_TU_ALL 1
#define BA_TU_S (-1)
#define BA_TU_G (-1)
#endif

// fast-phase group of a kernel: 0 = generic phase only, 1 = packed fast phase (FM 18 / 19), 2 = the same with the live
// borders in global memory (FM 34 / 35)
#define BA_TU_GROUP(R) ((R) == 0 ? 0 : ((R) < 32 ? 1 : 2))

#ifndef BA_TU_TOP
#ifdef BA_EMU
struct EmuLaunch { const Params* P; unsigned char* smem; uint32_t wg; };
template <int SCORING, int FLAGS, int FR>
static void emu_warp_entry(void* arg) {
  auto* l = (EmuLaunch*)arg;
  warp_main<SCORING, FLAGS, FR>(*l->P, l->smem, 0, l->wg);
}
template <int SCORING, int FLAGS, int FR>
static int launch_align(const Params& P, int blocks, int wpb, size_t smem_bytes, dev_stream_t) {
  (void)wpb;
  for (int b = 0; b < blocks; b++) {
    std::vector<unsigned char> smem(smem_bytes + 64, 0xAB);
    const int nm = SCORING == kNuc ? 128 : (SCORING == kAA ? 864 : 0);
    for (int i = 0; i < nm; i++) smem[i] = (unsigned char)P.matrix[i];
    stage_tables<SCORING>(smem.data(), 0, 1);
    EmuLaunch l{&P, smem.data(), (uint32_t)b};
    emu::run_warp(&emu_warp_entry<SCORING, FLAGS, FR>, &l);
    // guard: the device code must stay inside the shared memory it was given
    for (size_t i = smem_bytes; i < smem_bytes + 64; i++)
      if (smem[i] != 0xAB) { fprintf(stderr, "emu: shared memory overrun at byte %zu (limit %zu)\n", i, smem_bytes); abort(); }
  }
  return 0;
}
template <int SCORING, int FLAGS, int FR>
static int occupancy(int, size_t, int* blocks_per_sm) { *blocks_per_sm = 1; return 0; }
#else
// 128 threads per block, at least 4 blocks per SM: caps the kernel at 128 registers/thread. Measured on
// B200 (C2 workload): uncapped (207 regs, 8 warps/SM) 476 GCUPS, 160 regs 580, 128 regs 634, 96 regs 596.
#ifndef BA_LB_BLOCKS
#define BA_LB_BLOCKS 4
#endif
// TRACE kernels carry the trace-word accumulators on top of everything else: at 128 registers ptxas spills 390 bytes
// per thread (local-memory round trips inside the column loop); BA_LB_BLOCKS_TRACE = 3 gives them 168.
#ifndef BA_LB_BLOCKS_TRACE
#define BA_LB_BLOCKS_TRACE BA_LB_BLOCKS
#endif
// Protein kernels without X-drop or trace (C3) are held by instruction-cache misses and latency, not by the ALU pipe: a
// fifth block per SM (102 registers, ~150 spill instructions, 20 warps) is worth +5 % there (753 -> 791 GCUPS; 3 blocks
// 759, 6: 751, 8: 727), while the nucleotide X-drop kernel of C2 loses 7 % with it (1343 -> 1249).
#ifndef BA_LB_BLOCKS_AA
#define BA_LB_BLOCKS_AA 5
#endif
#ifndef BA_LB_BLOCKS_PROF
#define BA_LB_BLOCKS_PROF BA_LB_BLOCKS
#endif
template <int SCORING, int FLAGS> constexpr int lb_blocks() {
  return (FLAGS & kTrace) ? BA_LB_BLOCKS_TRACE : (SCORING == kProfile ? BA_LB_BLOCKS_PROF :
         ((SCORING == kAA && (FLAGS & (kTrace | kXDrop | kExt)) == 0) ? BA_LB_BLOCKS_AA : BA_LB_BLOCKS));
}
template <int SCORING, int FLAGS, int FR>
__global__ void __launch_bounds__(128, lb_blocks<SCORING, FLAGS>()) ba_align_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) unsigned char ba_smem[];
  const int nm = SCORING == kNuc ? 128 : (SCORING == kAA ? 864 : 0);
  for (int i = threadIdx.x; i < nm; i += blockDim.x) ba_smem[i] = (unsigned char)P.matrix[i];
  __syncthreads();
  stage_tables<SCORING>(ba_smem, (int)threadIdx.x, (int)blockDim.x);
  __syncthreads();
  const int wib = threadIdx.x >> 5;
  warp_main<SCORING, FLAGS, FR>(P, ba_smem, wib, blockIdx.x * (blockDim.x >> 5) + wib);
}
template <int SCORING, int FLAGS, int FR>
static int launch_align(const Params& P, int blocks, int wpb, size_t smem_bytes, dev_stream_t st) {
  CK(cudaFuncSetAttribute(ba_align_kernel<SCORING, FLAGS, FR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  ba_align_kernel<SCORING, FLAGS, FR><<<blocks, wpb * 32, smem_bytes, st>>>(P);
  CK(cudaGetLastError());
  return 0;
}
template <int SCORING, int FLAGS, int FR>
static int occupancy(int wpb, size_t smem_bytes, int* blocks_per_sm) {
  CK(cudaFuncSetAttribute(ba_align_kernel<SCORING, FLAGS, FR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ba_align_kernel<SCORING, FLAGS, FR>, wpb * 32, smem_bytes));
  return 0;
}
#endif
#endif  // !BA_TU_TOP

// kernel instantiations: scoring x flags x fast-phase mode (0 = generic phase only)
// flags 4..7 = kExt | {TRACE, X_DROP}: LOCAL_START / FREE_QUERY_*_GAPS selected at run time, generic phase only
// fast-phase modes 18 / 19: packed fast phase for min block 32 / 64; 34 / 35: the same with the live borders in global
// memory (max block size >= 1024)
#define BA_FOR_SEQ(X, S) X(S, 0, 0) X(S, 1, 0) X(S, 2, 0) X(S, 3, 0) X(S, 4, 0) X(S, 5, 0) X(S, 6, 0) X(S, 7, 0) \
                         X(S, 0, 18) X(S, 1, 18) X(S, 2, 18) X(S, 3, 18) X(S, 0, 19) X(S, 1, 19) X(S, 2, 19) X(S, 3, 19) \
                         X(S, 0, 34) X(S, 1, 34) X(S, 2, 34) X(S, 3, 34) X(S, 0, 35) X(S, 1, 35) X(S, 2, 35) X(S, 3, 35)
#ifdef BA_MINIMAL   // tuning builds (tools/build_variant.sh): only the kernels of the C2, C3, C4 and C5 workloads
#define BA_FOR_KERNELS(X) X(kNuc, 2, 0) X(kNuc, 2, 18) X(kNuc, 3, 19) X(kNuc, 3, 35) X(kAA, 0, 0) X(kAA, 0, 18) X(kProfile, 2, 0) X(kProfile, 2, 18)
#else
#define BA_FOR_KERNELS(X) BA_FOR_SEQ(X, kNuc) BA_FOR_SEQ(X, kAA) BA_FOR_SEQ(X, kByte) \
  X(kProfile, 0, 0) X(kProfile, 1, 0) X(kProfile, 2, 0) X(kProfile, 3, 0) \
  X(kProfile, 4, 0) X(kProfile, 5, 0) X(kProfile, 6, 0) X(kProfile, 7, 0) \
  X(kProfile, 0, 18) X(kProfile, 2, 18) X(kProfile, 0, 19) X(kProfile, 2, 19)
#endif

#ifndef BA_TU_TOP
#define BA_TU_MATCH(S, R) ((BA_TU_S < 0 || (S) == BA_TU_S) && (BA_TU_G < 0 || BA_TU_GROUP(R) == BA_TU_G))
// (templates with a dummy parameter: `if constexpr` discards -- i.e. does not instantiate -- only inside a template)
template <int Z>
static int tu_launch(int scoring, int flags, int fr, const Params& P, int blocks, int wpb, size_t smem, dev_stream_t st) {
#define X(S, F, R) if constexpr (BA_TU_MATCH(S, R)) { if (scoring == S && flags == F && fr == R) return launch_align<S + Z, F, R>(P, blocks, wpb, smem, st); }
  BA_FOR_KERNELS(X)
#undef X
  return fail(BA_ERR_ARG, "unsupported scoring/flags combination");
}
template <int Z>
static int tu_occupancy(int scoring, int flags, int fr, int wpb, size_t smem, int* bps) {
#define X(S, F, R) if constexpr (BA_TU_MATCH(S, R)) { if (scoring == S && flags == F && fr == R) return occupancy<S + Z, F, R>(wpb, smem, bps); }
  BA_FOR_KERNELS(X)
#undef X
  return fail(BA_ERR_ARG, "unsupported scoring/flags combination");
}
#endif

#define BA_CAT3_(a, b, c) a##b##_##c
#define BA_CAT3(a, b, c) BA_CAT3_(a, b, c)
#if defined(BA_TU_ALL)
int ba_launch_dispatch(int scoring, int flags, int fr, const Params& P, int blocks, int wpb, size_t smem, dev_stream_t st) {
  return tu_launch<0>(scoring, flags, fr, P, blocks, wpb, smem, st);
}
int ba_occupancy_dispatch(int scoring, int flags, int fr, int wpb, size_t smem, int* bps) { return tu_occupancy<0>(scoring, flags, fr, wpb, smem, bps); }
#elif !defined(BA_TU_TOP)
int BA_CAT3(ba_launch_tu_, BA_TU_S, BA_TU_G)(int scoring, int flags, int fr, const Params& P, int blocks, int wpb, size_t smem, dev_stream_t st) {
  return tu_launch<0>(scoring, flags, fr, P, blocks, wpb, smem, st);
}
int BA_CAT3(ba_occupancy_tu_, BA_TU_S, BA_TU_G)(int scoring, int flags, int fr, int wpb, size_t smem, int* bps) { return tu_occupancy<0>(scoring, flags, fr, wpb, smem, bps); }
#else
// the (scoring kind, group) units the product build compiles: sequence kinds 0..2 x groups 0..2, profiles (3) x groups 0..1
#define BA_FOR_TUS(Y) Y(0, 0) Y(0, 1) Y(0, 2) Y(1, 0) Y(1, 1) Y(1, 2) Y(2, 0) Y(2, 1) Y(2, 2) Y(3, 0) Y(3, 1)
#define Y(S, G) int ba_launch_tu_##S##_##G(int, int, int, const ba::Params&, int, int, size_t, dev_stream_t); \
                int ba_occupancy_tu_##S##_##G(int, int, int, int, size_t, int*);
BA_FOR_TUS(Y)
#undef Y
int ba_launch_dispatch(int scoring, int flags, int fr, const ba::Params& P, int blocks, int wpb, size_t smem, dev_stream_t st) {
  const int g = BA_TU_GROUP(fr);
#define Y(S, G) if (scoring == S && g == G) return ba_launch_tu_##S##_##G(scoring, flags, fr, P, blocks, wpb, smem, st);
  BA_FOR_TUS(Y)
#undef Y
  return fail(BA_ERR_ARG, "unsupported scoring/flags combination");
}
int ba_occupancy_dispatch(int scoring, int flags, int fr, int wpb, size_t smem, int* bps) {
  const int g = BA_TU_GROUP(fr);
#define Y(S, G) if (scoring == S && g == G) return ba_occupancy_tu_##S##_##G(scoring, flags, fr, wpb, smem, bps);
  BA_FOR_TUS(Y)
#undef Y
  return fail(BA_ERR_ARG, "unsupported scoring/flags combination");
}
#endif
