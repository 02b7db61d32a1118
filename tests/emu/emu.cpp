#include "emu.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

extern "C" void emu_switch(void** from_sp, void* to_sp);
// x86-64 SysV context switch: callee-saved registers + stack pointer.
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
)");

namespace emu {
namespace {
constexpr int kLanes = 32;
constexpr size_t kStack = 512 * 1024;
struct WarpCtx {
  void* sp[kLanes];
  void* main_sp;
  bool done[kLanes];
  int cur;
  int ndone;
  uint32_t vals[2][kLanes];
  int arrived;
  unsigned gen;
  void (*fn)(void*);
  void* arg;
  char* stacks;
};
thread_local WarpCtx* g = nullptr;

void switch_to_next_from(int me, bool finished) {
  if (finished && g->ndone == kLanes) {
    void* dummy;
    emu_switch(&dummy, g->main_sp);
    abort();
  }
  int nxt = me;
  for (int t = 0; t < kLanes; t++) {
    nxt = (nxt + 1) % kLanes;
    if (!g->done[nxt]) break;
  }
  if (nxt == me && !finished) return;
  g->cur = nxt;
  if (finished) { void* dummy; emu_switch(&dummy, g->sp[nxt]); abort(); }
  emu_switch(&g->sp[me], g->sp[nxt]);
}

void fiber_entry() {
  const int me = g->cur;
  g->fn(g->arg);
  g->done[me] = true;
  g->ndone++;
  switch_to_next_from(me, true);
  abort();
}
}  // namespace

int lane() { return g->cur; }

const uint32_t* exchange(uint32_t v) {
  const int me = g->cur;
  const unsigned gen = g->gen;
  const int buf = gen & 1;
  g->vals[buf][me] = v;
  g->arrived++;
  if (g->arrived == kLanes - g->ndone) {
    g->arrived = 0;
    g->gen++;
  } else {
    while (g->gen == gen) switch_to_next_from(me, false);
  }
  return g->vals[buf];
}

void run_warp(void (*fn)(void*), void* arg) {
  WarpCtx ctx;
  memset(&ctx, 0, sizeof(ctx));
  ctx.fn = fn; ctx.arg = arg;
  ctx.stacks = (char*)aligned_alloc(64, kStack * kLanes);
  for (int l = 0; l < kLanes; l++) {
    uintptr_t top = ((uintptr_t)(ctx.stacks + kStack * (l + 1))) & ~(uintptr_t)15;
    void** sp = (void**)(top - 8 * sizeof(void*));
    for (int k = 0; k < 6; k++) sp[k] = nullptr;   // r15 r14 r13 r12 rbx rbp
    sp[6] = (void*)&fiber_entry;                    // return address of emu_switch
    sp[7] = nullptr;                                // fake caller return address
    ctx.sp[l] = sp;
  }
  WarpCtx* saved = g;
  g = &ctx;
  ctx.cur = 0;
  emu_switch(&ctx.main_sp, ctx.sp[0]);
  g = saved;
  free(ctx.stacks);
}
}  // namespace emu
