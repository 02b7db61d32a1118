// ba_runtime.cu -- host side of the batch API (include/block_aligner_b200.h, Part 2) and the
// __global__ entry points. One BaAligner per GPU; no CPU execution path exists in the product build.
//
// The same file is compiled by g++ with -DBA_EMU into tests/emu/libba_emu.so, where "device
// memory" is host memory and kernels run under the fiber emulator. That artefact is test
// infrastructure (it lets `pytest -m "not gpu"` check the device source and this host logic
// against the oracle); it is never loaded by the product.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/block_aligner_b200.h"
#include "ba_host.h"
#include "ba_dev.h"
#include "ba_kernel.cuh"

using namespace ba;

namespace ba { std::string& last_error_ref() { static thread_local std::string e; return e; } }
#define g_last_error (ba::last_error_ref())

// -------------------------------------------------------------------------------------------------
// kernels
// -------------------------------------------------------------------------------------------------
// PaddedBytes::from_bytes (scan_block.rs:1829-1836): [NULL] + convert(bytes) + pad x [NULL]
struct PackArgs {
  const uint8_t* raw_q; const uint64_t* raw_q_off;
  const uint8_t* raw_r; const uint64_t* raw_r_off;   // raw_r may be null (profiles)
  uint8_t* seq; const uint64_t* pq_off; const uint64_t* pr_off;
  uint32_t n; uint32_t pad; int32_t scoring; uint32_t* err;
  uint32_t rev_q, rev_r;      // PaddedBytes::set_bytes_rev (scan_block.rs:1815-1822)
  uint32_t nuc4;              // BA_INPUT_NUC4: raw arenas hold 4-bit codes, raw offsets count nibbles (high nibble first)
};
// BAM / htslib nibble codes -> the ASCII letter the reference would have been given ("=ACMGRSVTWYHKDBN"); code 0 ('=') is
// not a base and is reported as a bad character
BA_HD uint8_t nuc4_letter(uint32_t code) {
  const uint64_t lo = 0x565352474d434100ull, hi = 0x4e42444b48595754ull;   // bytes: \0 A C M G R S V | T W Y H K D B N
  return (uint8_t)(((code & 8u) ? hi : lo) >> (8u * (code & 7u)));
}

// One padded device profile (ProfBuild, ba_types.h). `tid` / `nthreads`: the calling thread's share of the profile.
// Layout written: pos_aa [curr_len][32] i8, then gap_open_C, gap_close_C, gap_open_R [curr_len] i16.
BA_HD void prof_build_one(const ProfBuildArgs& a, uint32_t k, uint32_t tid, uint32_t nthreads, bool& bad) {
  const ProfBuild d = a.desc[k];
  uint8_t* base = a.arena + d.dst_off;
  const uint32_t cl = d.curr_len;
  uint32_t* w = (uint32_t*)base;
  int16_t* oc = (int16_t*)(base + (uint64_t)cl * 32);
  int16_t* cc = oc + cl;
  int16_t* orr = cc + cl;
  if (a.kind == 0) {
    const uint32_t* src = (const uint32_t*)(a.stage + d.src_off);
    const uint32_t nw = d.np * 8;
    for (uint32_t t = tid; t < cl * 8; t += nthreads) w[t] = t < nw ? src[t] : 0x80808080u;
    const int16_t* g = (const int16_t*)(a.stage + d.src_off + (uint64_t)d.np * 32);
    for (uint32_t t = tid; t < cl; t += nthreads) {
      const bool in = t < d.np;
      oc[t] = in ? g[t] : (int16_t)-128;
      cc[t] = in ? g[d.np + t] : (int16_t)-128;
      orr[t] = in ? g[2 * d.np + t] : (int16_t)-128;
    }
  } else {
    const uint32_t len = d.np;
    const int8_t* sc = a.scores + d.src_off;
    for (uint32_t t = tid; t < cl * 8; t += nthreads) {
      const uint32_t pos = t >> 3, b0 = (t & 7u) * 4;
      uint32_t word = 0x80808080u;
      if (pos >= 1 && pos <= len) {
        // set_all walks positions 1..=len, set_all_rev len..=1, consuming one score row each (scores.rs:686-712)
        const uint32_t n = a.rev ? len - pos : pos - 1;
        const int8_t* row = sc + (uint64_t)n * a.order_len;
        word = 0;
        for (int e = 0; e < 4; e++) {
          const int j = a.inv[b0 + e];
          const int8_t v = j >= 0 ? (int8_t)((int8_t)(row[j] << a.left_shift) >> a.right_shift) : (int8_t)-128;
          word |= (uint32_t)(uint8_t)v << (8 * e);
        }
      }
      w[t] = word;
    }
    for (uint32_t t = tid; t < cl; t += nthreads) {
      int g0 = -128, g1 = -128, g2 = -128;
      if (t <= len) {
        g0 = a.gap_open_C ? (int)a.gap_open_C[d.gap_src_off + t] : a.all_open_C;
        g1 = a.gap_close_C ? (int)a.gap_close_C[d.gap_src_off + t] : a.all_close_C;
        g2 = a.gap_open_R ? (int)a.gap_open_R[d.gap_src_off + t] : a.all_open_R;
        if (g0 >= 0 || g2 >= 0) bad = true;     // "Gap open cost must be negative!" (scores.rs:549, 560)
      }
      oc[t] = (int16_t)g0; cc[t] = (int16_t)g1; orr[t] = (int16_t)g2;
    }
  }
}

// byte offsets of the derived layouts inside a profile's arena slot, and the slot size
BA_HD uint64_t prof_tlen(uint64_t cl) { return (cl + 15) & ~(uint64_t)15; }
BA_HD uint64_t prof_gp_off(uint64_t cl) { return (cl * 38 + 15) & ~(uint64_t)15; }
BA_HD uint64_t prof_tp_off(uint64_t cl) { return prof_gp_off(cl) + cl * 16; }
BA_HD uint64_t prof_slot_bytes(uint64_t cl, bool derive) { return ((derive ? prof_tp_off(cl) + 32 * prof_tlen(cl) : cl * 38) + 63) & ~(uint64_t)63; }
// second pass over a built profile (after a barrier): the transposed scores and the packed per-position gap operands
BA_HD void prof_derive_one(const ProfBuildArgs& a, uint32_t k, uint32_t tid, uint32_t nthreads) {
  const ProfBuild d = a.desc[k];
  if (!d.derive) return;
  uint8_t* base = a.arena + d.dst_off;
  const uint32_t cl = d.curr_len, tlen = (uint32_t)prof_tlen(cl);
  const int8_t* pos_aa = (const int8_t*)base;
  const int16_t* oc = (const int16_t*)(base + (uint64_t)cl * 32);
  const int16_t* cc = oc + cl;
  const int16_t* orr = cc + cl;
  uint32_t* gp = (uint32_t*)(base + prof_gp_off(cl));
  int8_t* tp = (int8_t*)(base + prof_tp_off(cl));
  for (uint32_t t = tid; t < cl; t += nthreads) {
    gp[4 * t + 0] = (uint32_t)(((int)oc[t] + d.gap_extend) * 65537);
    gp[4 * t + 1] = (uint32_t)((int)cc[t] * 65537);
    gp[4 * t + 2] = (uint32_t)((int)orr[t] * 65537);
    gp[4 * t + 3] = 0u;
  }
  // one 32-bit word = 4 consecutive positions of one residue
  for (uint32_t t = tid; t < 32 * (tlen / 4); t += nthreads) {
    const uint32_t res = t / (tlen / 4), p0 = (t % (tlen / 4)) * 4;
    uint32_t word = 0;
    for (uint32_t e = 0; e < 4; e++) {
      const uint32_t pos = p0 + e;
      const int8_t v = pos < cl ? pos_aa[(uint64_t)pos * 32 + res] : (int8_t)-128;
      word |= (uint32_t)(uint8_t)v << (8 * e);
    }
    ((uint32_t*)tp)[t] = word;
  }
}

#ifdef BA_EMU
static void prof_build_all(const ProfBuildArgs& a) {
  bool bad = false;
  for (uint32_t k = 0; k < a.n; k++) { prof_build_one(a, k, 0, 1, bad); prof_derive_one(a, k, 0, 1); }
  if (bad) *a.err = 1;
}
static void pack_all(const PackArgs& a) {
  const int nseq = a.raw_r ? 2 : 1;
  for (uint32_t k = 0; k < a.n; k++)
    for (int which = 0; which < nseq; which++) {
      const uint8_t* src = which ? a.raw_r + a.raw_r_off[k] : a.raw_q + a.raw_q_off[k];
      const uint64_t len = which ? a.raw_r_off[k + 1] - a.raw_r_off[k] : a.raw_q_off[k + 1] - a.raw_q_off[k];
      uint8_t* dst = a.seq + (which ? a.pr_off[k] : a.pq_off[k]);
      const uint8_t nul = host::null_code(a.scoring);
      dst[0] = nul;
      const bool rev = which ? a.rev_r != 0 : a.rev_q != 0;
      if (a.nuc4) {
        const uint8_t* arena = which ? a.raw_r : a.raw_q;
        const uint64_t o0 = which ? a.raw_r_off[k] : a.raw_q_off[k];
        for (uint64_t t = 0; t < len; t++) {
          const uint64_t nb = o0 + (rev ? len - 1 - t : t);
          const uint32_t code = (nb & 1) ? (arena[nb >> 1] & 15u) : (arena[nb >> 1] >> 4);
          dst[1 + t] = nuc4_letter(code);
          if (code == 0) *a.err = 1;
        }
      } else
      for (uint64_t t = 0; t < len; t++) { bool ok; dst[1 + t] = host::convert_char(a.scoring, src[rev ? len - 1 - t : t], &ok); if (!ok) *a.err = 1; }
      for (uint32_t t = 0; t < a.pad; t++) dst[1 + len + t] = nul;
    }
}
struct EmuTb { const Params* P; uint32_t pair, qi, rj; int eq; DevResult* out1; };
static void emu_tb_entry(void* arg) {
  auto* t = (EmuTb*)arg;
  static uint8_t lut[128];
  for (uint32_t i = 0; i < 128; i++) lut[i] = tb_entry(i);
  warp_traceback(*t->P, lut, t->pair, t->qi, t->rj, t->eq != 0, t->out1);
}
static int launch_traceback(const Params& P, uint32_t pair, uint32_t qi, uint32_t rj, int eq, DevResult* out1, dev_stream_t) {
  EmuTb t{&P, pair, qi, rj, eq, out1};
  emu::run_warp(&emu_tb_entry, &t);
  return 0;
}
struct EmuTbBatch { const Params* P; const uint32_t* list; uint32_t n, warp; };
static void emu_tb_batch_entry(void* arg) {
  auto* t = (EmuTbBatch*)arg;
  static uint8_t lut[128];
  for (uint32_t i = 0; i < 128; i++) lut[i] = tb_entry(i);
  // BA_TB_STAGED: the shape of the staged launch (eight lanes per walk, four walks per warp)
  static uint32_t stage[4 * 2 * kTbStageWords];
  const uint32_t lane = (uint32_t)wp::lane_id();
  if (!getenv("BA_TB_STAGED")) lanes_traceback<false>(*t->P, lut, t->list, t->n, t->warp * 32 + lane);
  else lanes_traceback<true>(*t->P, lut, t->list, t->n, (lane & 7u) ? 0xffffffffu : t->warp * 4 + (lane >> 3), stage);
}
static int launch_traceback_batch(const Params& P, const uint32_t* list, uint32_t n, dev_stream_t) {
  const uint32_t per = !getenv("BA_TB_STAGED") ? 32u : 4u;
  for (uint32_t wi = 0; wi * per < n; wi++) { EmuTbBatch t{&P, list, n, wi}; emu::run_warp(&emu_tb_batch_entry, &t); }
  return 0;
}
#else
__global__ void ba_pack_kernel(PackArgs a) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t nseq = a.raw_r ? 2 * a.n : a.n;
  const uint8_t nul = a.scoring == kNuc ? (uint8_t)'Z' : (a.scoring == kByte ? (uint8_t)0 : (uint8_t)26);
  bool bad = false;
  for (uint32_t s = warp; s < nseq; s += nwarps) {
    const uint32_t k = a.raw_r ? (s >> 1) : s;
    const bool which = a.raw_r ? (s & 1) : false;
    const uint8_t* src = which ? a.raw_r + a.raw_r_off[k] : a.raw_q + a.raw_q_off[k];
    const uint64_t len = which ? a.raw_r_off[k + 1] - a.raw_r_off[k] : a.raw_q_off[k + 1] - a.raw_q_off[k];
    uint8_t* dst = a.seq + (which ? a.pr_off[k] : a.pq_off[k]);
    if (lane == 0) dst[0] = nul;
    const bool rev = which ? a.rev_r != 0 : a.rev_q != 0;
    if (a.nuc4) {
      const uint8_t* arena = which ? a.raw_r : a.raw_q;
      const uint64_t o0 = which ? a.raw_r_off[k] : a.raw_q_off[k];
      for (uint64_t t = lane; t < len; t += 32) {
        const uint64_t nb = o0 + (rev ? len - 1 - t : t);
        const uint32_t byte = arena[nb >> 1];
        const uint32_t code = (nb & 1) ? (byte & 15u) : (byte >> 4);
        if (code == 0) bad = true;
        dst[1 + t] = nuc4_letter(code);
      }
    } else
    for (uint64_t t = lane; t < len; t += 32) {
      uint8_t c = src[rev ? len - 1 - t : t];
      if (a.scoring != kByte) {
        if (c >= 'a' && c <= 'z') c -= 32;                      // to_ascii_uppercase
        if (a.scoring == kNuc) { if (!(c >= 'A' && c <= 'Z')) bad = true; }          // scores.rs:212-216
        else { if (!(c >= 'A' && c <= 'A' + 26)) bad = true; c = (uint8_t)(c - 'A'); } // scores.rs:130-134
      }
      dst[1 + t] = c;
    }
    for (uint32_t t = lane; t < a.pad; t += 32) dst[1 + len + t] = nul;
  }
  if (bad) atomicOr(a.err, 1u);
}

// one CTA per profile (grid-stride): coalesced 32-bit stores over the whole padded profile
__global__ void ba_profile_build_kernel(ProfBuildArgs a) {
  bool bad = false;
  for (uint32_t k = blockIdx.x; k < a.n; k += gridDim.x) {
    prof_build_one(a, k, threadIdx.x, blockDim.x, bad);
    __syncthreads();
    prof_derive_one(a, k, threadIdx.x, blockDim.x);
  }
  if (bad) atomicOr(a.err, 1u);
}

__global__ void ba_traceback_kernel(const __grid_constant__ Params P, uint32_t pair, uint32_t qi, uint32_t rj, int eq, DevResult* out1) {
  __shared__ uint8_t lut[128];
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) lut[i] = tb_entry(i);
  __syncthreads();
  warp_traceback(P, lut, pair, qi, rj, eq != 0, out1);
}
static int launch_traceback(const Params& P, uint32_t pair, uint32_t qi, uint32_t rj, int eq, DevResult* out1, dev_stream_t st) {
  ba_traceback_kernel<<<1, 32, 0, st>>>(P, pair, qi, rj, eq, out1);
  CK(cudaGetLastError());
  return 0;
}
// one walk per lane over the pairs of a launch whose traces are resident (list = the launch's work list)
// `stride`: every stride-th lane walks (32 / stride walks per warp, see launch_traceback_batch)
// ten blocks per SM (48 registers): with 63 registers a launch of two walks per warp does not fit the GPU in one wave
#ifndef BA_TB_LB
#define BA_TB_LB 10
#endif
__global__ void __launch_bounds__(128, BA_TB_LB) ba_traceback_batch_kernel(const __grid_constant__ Params P, const uint32_t* list, uint32_t n, uint32_t stride) {
  __shared__ uint8_t lut[128];
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) lut[i] = tb_entry(i);
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  lanes_traceback<false>(P, lut, list, n, (t % stride) ? 0xffffffffu : t / stride);
}
// eight lanes per walk; the seven lanes that do not walk help to stage the next rectangle's trace words in shared memory
__global__ void __launch_bounds__(128) ba_traceback_staged_kernel(const __grid_constant__ Params P, const uint32_t* list, uint32_t n) {
  __shared__ uint8_t lut[128];
  __shared__ __align__(16) uint32_t stage[4 * 4 * 2 * kTbStageWords];     // 4 warps x 4 walks x 2 buffers
  for (uint32_t i = threadIdx.x; i < 128; i += blockDim.x) lut[i] = tb_entry(i);
  __syncthreads();
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t ngroups = (gridDim.x * blockDim.x) >> 3;
  // the grid holds fewer groups than there are walks (launch_traceback_batch): every group takes walks t, t + ngroups, ...
  for (uint32_t base = 0; base < n; base += ngroups)
    lanes_traceback<true>(P, lut, list, n, (t & 7u) ? 0xffffffffu : base + (t >> 3), stage + (threadIdx.x >> 5) * (4 * 2 * kTbStageWords));
}
static int launch_traceback_batch(const Params& P, const uint32_t* list, uint32_t n, dev_stream_t st) {
  if (!n) return 0;
  // Walks per warp. Every walk is a chain of dependent loads through its own pages (trace arena, sequences, records,
  // run scratch), and the walks of one warp share every load instruction: measured on B200 with 10 k walks of ~90 k
  // steps (C5), 32 walks per warp take 106 ms, 16: 87, 8: 66, 4: 55, 2: 74 (more resident warps per SM than the memory
  // system serves well), 1: 84. So: four walks per warp while the launch still fits the GPU at once, more as the batch grows.
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  // BA_TB_STAGED=1: eight lanes per walk, the seven idle ones copy the next rectangle's trace words into shared memory
  // with asynchronous copies (ba_traceback_staged_kernel), BA_TB_WALKS = walks in flight. Kept for the record, OFF: on
  // B200 / C5 the staged walk takes 57.8 ms against 55.2 ms, and with fewer walks in flight the time grows in proportion
  // (4096: +76 ms, 2048: +160 ms) -- a walk is not held by the latency of its trace words but by its own chain of ~90
  // dependent instructions per cell step (ncu r02: 23 G warp instructions, issue slots 40 % busy with 4 warps per
  // scheduler; profiles/r02_traceback_experiments.txt).
  if (getenv("BA_TB_STAGED")) {
    uint32_t walks = 1u << 20;
    if (const char* e = getenv("BA_TB_WALKS")) walks = (uint32_t)std::max(16, atoi(e));
    const uint64_t threads8 = (uint64_t)std::min<uint32_t>(n, walks) * 8;
    ba_traceback_staged_kernel<<<(unsigned)((threads8 + 127) / 128), 128, 0, st>>>(P, list, n);
    CK(cudaGetLastError());
    return 0;
  }
  // (round 2: the 74 ms of two walks per warp were a second wave -- at 63 registers 5000 warps do not fit at once. With
  // the kernel capped at 48 registers they do: two walks per warp 45 ms, four 49 ms; C5 step 206 -> 196 ms.)
  uint32_t stride = (uint32_t)std::min<uint64_t>(16, std::max<uint64_t>(1, (uint64_t)sms * 1280 / n));
  if (const char* e = getenv("BA_TB_STRIDE")) stride = (uint32_t)std::max(1, atoi(e));
  const uint64_t threads = (uint64_t)n * stride;
  ba_traceback_batch_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(P, list, n, stride);
  CK(cudaGetLastError());
  return 0;
}

#endif

// -------------------------------------------------------------------------------------------------
// objects
// -------------------------------------------------------------------------------------------------
// one stream + events per batch in flight, so that the H2D copies of one batch overlap the kernel of another
struct StreamSet {
  dev_stream_t stream = 0;
#ifndef BA_EMU
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp0 = nullptr, evp1 = nullptr;
#endif
};

struct BaAligner {
  int device = 0;
  dev_stream_t stream = 0;
  int sm_count = 1;
  size_t smem_optin = 48 * 1024;
  size_t mem_total = 0;
#ifndef BA_EMU
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
#endif
  double budget_frac = 0.55;   // share of device memory one batch may take for per-slot trace arenas
  int emu_warps = 3;   // emulation: number of (sequentially executed) warps, to exercise per-warp scratch
  // Device-buffer pool: cudaMalloc / cudaFree cost tens to hundreds of milliseconds for GB-sized buffers
  // (measured: 26-257 ms and 8-613 ms per call on B200), far more than the 40 ms H2D copy of a 2 GB batch,
  // so buffers of freed batches are kept and handed to the next batch.
  std::vector<struct StreamSet> streams_free;
  std::vector<std::pair<void*, size_t>> pool_free;
  std::vector<std::pair<void*, size_t>> pool_live;
  size_t pool_cached = 0;
  // pinned host staging for data the library has to re-pack before the upload (host AAProfile objects); grow-only
  uint8_t* h_stage = nullptr; size_t h_stage_cap = 0;
#ifndef BA_EMU
  // pinned bounce buffers for sequence arenas that arrive in ordinary (pageable) host memory, two per copy thread
  // (h2d_seq); allocated on first use
  static constexpr int kStageThreads = 8;      // upper bound; BA_STAGE_THREADS (default 4) of them are used
  static constexpr size_t kStageBlock = (size_t)8 << 20;
  uint8_t* ring[kStageThreads][2] = {};
  cudaEvent_t ring_ev[kStageThreads][2] = {};
#endif
  size_t pool_live_bytes = 0, pool_peak = 0;   // bytes handed out now / high-water mark since the cache was last trimmed
  // Every public entry point that touches an aligner (or one of its batches) holds this lock for its whole body, so
  // calls on one BaAligner from several host threads are serialised instead of racing on the pool / stream / staging
  // bookkeeping (the Part 1 calls of all threads share one aligner). Recursive: ba_align_batch calls ba_batch_upload.
  std::recursive_mutex mu;
};
typedef std::lock_guard<std::recursive_mutex> AlLock;

static int stage_reserve(BaAligner* al, size_t n) {
  if (n <= al->h_stage_cap) return 0;
#ifdef BA_EMU
  free(al->h_stage);
  al->h_stage = (uint8_t*)malloc(n);
  if (!al->h_stage) { al->h_stage_cap = 0; return 1; }
#else
  if (al->h_stage) cudaFreeHost(al->h_stage);
  al->h_stage = nullptr; al->h_stage_cap = 0;
  CK(cudaHostAlloc((void**)&al->h_stage, n, cudaHostAllocDefault));
#endif
  al->h_stage_cap = n;
  return 0;
}

// Host -> device copy of a sequence arena. From pinned memory: one cudaMemcpyAsync. From pageable memory the runtime's
// own staging is a single-threaded bounce at ~5 GB/s that blocks the calling thread (C2, 1.05 GB: 228 ms per batch
// against 83 ms from pinned buffers); here four threads copy 8 MB blocks into pinned double buffers and issue the DMA
// from there, so that a caller with malloc'ed reads -- the usual drop-in case -- gets most of the pinned throughput.
#ifdef BA_EMU
static int h2d_seq(BaAligner*, void* d, const uint8_t* s, size_t n, dev_stream_t st) { return h2d(d, s, n, st); }
#else
static int h2d_seq(BaAligner* al, void* d, const uint8_t* s, size_t n, dev_stream_t st) {
  size_t min_bytes = (size_t)4 << 20;
  if (const char* e = getenv("BA_STAGE_MIN_BYTES")) min_bytes = (size_t)atoll(e);
  if (n < min_bytes) return h2d(d, s, n, st);
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, s) != cudaSuccess) { cudaGetLastError(); return h2d(d, s, n, st); }
  if (at.type != cudaMemoryTypeUnregistered) return h2d(d, s, n, st);     // pinned / managed: the DMA engine reads it directly
  int T = 4;
  if (const char* e = getenv("BA_STAGE_THREADS")) T = std::max(1, std::min(atoi(e), (int)BaAligner::kStageThreads));
  constexpr size_t B = BaAligner::kStageBlock;
  for (int t = 0; t < T; t++) for (int k = 0; k < 2; k++) if (!al->ring[t][k]) {
    CK(cudaHostAlloc((void**)&al->ring[t][k], B, cudaHostAllocDefault));
    CK(cudaEventCreateWithFlags(&al->ring_ev[t][k], cudaEventDisableTiming));
  }
  const size_t nb = (n + B - 1) / B;
  int err[BaAligner::kStageThreads] = {};
  std::thread th[BaAligner::kStageThreads];
  for (int t = 0; t < T; t++) th[t] = std::thread([&, t]() {
    if (cudaSetDevice(al->device) != cudaSuccess) { err[t] = 1; return; }
    for (size_t j = (size_t)t, k = 0; j < nb; j += T, k++) {
      const int slot = (int)(k & 1);
      const size_t off = j * B, len = std::min(B, n - off);
      if (cudaEventSynchronize(al->ring_ev[t][slot]) != cudaSuccess) { err[t] = 1; return; }    // the copy that last used this buffer
      memcpy(al->ring[t][slot], s + off, len);
      if (cudaMemcpyAsync((uint8_t*)d + off, al->ring[t][slot], len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
          cudaEventRecord(al->ring_ev[t][slot], st) != cudaSuccess) { err[t] = 1; return; }
    }
  });
  for (int t = 0; t < T; t++) th[t].join();
  for (int t = 0; t < T; t++) if (err[t]) return cuda_fail(cudaGetLastError(), "staged host -> device copy");
  return 0;
}
#endif

static int streams_acquire(BaAligner* al, StreamSet* ss) {
  if (!al->streams_free.empty()) { *ss = al->streams_free.back(); al->streams_free.pop_back(); return 0; }
#ifndef BA_EMU
  CK(cudaStreamCreateWithFlags(&ss->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&ss->ev0)); CK(cudaEventCreate(&ss->ev1)); CK(cudaEventCreate(&ss->evp0)); CK(cudaEventCreate(&ss->evp1));
#endif
  return 0;
}
static void streams_release(BaAligner* al, const StreamSet& ss) { al->streams_free.push_back(ss); }

static void pool_note_live(BaAligner* al, size_t sz) {
  al->pool_live_bytes += sz;
  if (al->pool_live_bytes > al->pool_peak) al->pool_peak = al->pool_live_bytes;
}
static int pool_alloc(BaAligner* al, void** p, size_t n) {
  if (n == 0) n = 1;
  int best = -1;
  for (size_t i = 0; i < al->pool_free.size(); i++) {
    const size_t sz = al->pool_free[i].second;
    if (sz >= n && sz <= 2 * n + (1 << 20) && (best < 0 || sz < al->pool_free[best].second)) best = (int)i;
  }
  if (best >= 0) {
    *p = al->pool_free[best].first;
    al->pool_live.push_back(al->pool_free[best]);
    al->pool_cached -= al->pool_free[best].second;
    pool_note_live(al, al->pool_free[best].second);
    al->pool_free.erase(al->pool_free.begin() + best);
    return 0;
  }
  int rc = dmalloc(p, n);
  if (rc) {   // out of memory: drop the cache and retry once
    for (auto& e : al->pool_free) dfree(e.first);
    al->pool_free.clear(); al->pool_cached = 0;
#ifndef BA_EMU
    cudaGetLastError();
#endif
    rc = dmalloc(p, n);
    if (rc) return rc;
  }
  al->pool_live.push_back({*p, n});
  pool_note_live(al, n);
  return 0;
}
static void pool_release(BaAligner* al, void* p) {
  if (!p) return;
  for (size_t i = 0; i < al->pool_live.size(); i++) {
    if (al->pool_live[i].first == p) {
      const size_t sz = al->pool_live[i].second;
      al->pool_live.erase(al->pool_live.begin() + i);
      al->pool_live_bytes -= std::min(al->pool_live_bytes, sz);
      // The cache is bounded by what the caller's recent batches actually used: twice the high-water mark of live
      // bytes (+ 64 MB), never more than 90 % of the device. A process that aligns one pair at a time keeps a few MB;
      // C5 (trace arenas = 55 % of the device; with a 50 % cap some of them were cudaFree'd and cudaMalloc'ed again on
      // every call, 400 ms each time) keeps its arenas. ba_trim() empties the cache and resets the mark.
      const size_t cap = std::min<size_t>(al->mem_total / 10 * 9, 2 * al->pool_peak + ((size_t)64 << 20));
      if (al->pool_cached + sz <= cap) { al->pool_free.push_back({p, sz}); al->pool_cached += sz; }
      else dfree(p);
      return;
    }
  }
  dfree(p);   // not from the pool
}

struct BaBatch {
  BaAligner* al = nullptr;
  BaConfig cfg;
  size_t n = 0;
  uint32_t min_size = 0, max_size = 0;
  uint32_t max_pair_len = 0;
  // device
  uint8_t* d_seq = nullptr; uint64_t* d_qoff = nullptr; uint64_t* d_roff = nullptr;
  uint32_t* d_qlen = nullptr; uint32_t* d_rlen = nullptr; uint32_t* d_order = nullptr;
  int8_t* d_matrix = nullptr; ProfileDev* d_profiles = nullptr; uint8_t* d_prof_arena = nullptr;
  DevResult* d_out = nullptr; uint32_t* d_ticket = nullptr;
  int16_t* d_ckpt = nullptr; uint32_t* d_trace = nullptr; Rect* d_rects = nullptr; uint32_t* d_runs = nullptr;
  uint32_t* d_cigar = nullptr; unsigned long long* d_cigar_used = nullptr; uint64_t cigar_cap = 0;
  StepLog* d_steplog = nullptr; uint32_t* d_steplog_n = nullptr;
  uint32_t* d_overflow_list = nullptr; uint32_t* d_overflow_n = nullptr;
  uint32_t* d_zwords = nullptr;     // zero masks (TRACE && LOCAL_START)
  // TRACE: every pair of a launch owns one trace arena; a wave = the pairs of one launch (their arenas fit the budget)
  struct Wave { uint32_t t0 = 0, n = 0; uint64_t words = 0; uint32_t rects = 0, runs = 0; };
  std::vector<Wave> waves;              // first pass, over the processing order
  Wave cur;                             // geometry of the wave that ran last (its traces are the resident ones)
  std::vector<uint32_t> h_order;        // processing order (pair ids, longest first)
  std::vector<uint32_t> h_arena_of;     // pair -> arena index inside its wave
  std::vector<uint8_t> resident;        // pairs whose trace is still in an arena (ba_batch_traceback)
  uint32_t* d_arena_of = nullptr; bool h_arena_of_dirty = false;
  uint64_t cap_trace = 0, cap_rects = 0, cap_runs = 0;   // bytes allocated for the arenas
  std::vector<uint32_t> h_qlen, h_rlen;   // lengths, kept for the bounds check of ba_batch_traceback
  bool arena_forced = false;
  int kflags = 0;                   // template FLAGS of the kernel: (flags & 3) | kExt
  int pk_smax = 0; uint32_t pk_enable = 0;   // packed 2 x i16 path (ba_packed.cuh): largest matrix entry, on/off
  bool prof_fast = false; int prof_ge = -1;  // profile batch served by the packed fast phase; its (uniform) gap_extend
  uint64_t mem_budget = 0; uint64_t max_blocks_hw = 1;
  // launch geometry
  int blocks = 0, wpb = 0; size_t smem_bytes = 0; uint32_t slots_per_warp = 1; int fast_rows = 0; bool gb = false;
  int16_t* d_gborders = nullptr;
  uint32_t* d_trace_pool = nullptr; uint32_t* d_trace_pool_cursor = nullptr; uint64_t trace_pool_units = 0;
  // host results
  std::vector<DevResult> h_out;
  std::vector<uint32_t> h_cigar;
  uint32_t* cigar_dst = nullptr; uint64_t cigar_dst_cap = 0; uint64_t cigar_dst_used = 0;   // ba_align_batch_cigar: caller's buffer
  bool cigar_skip = false;   // ba_align_batch (no CIGAR output): do not copy the stream back at all
  std::vector<uint32_t> h_tb;      // runs of the last ba_batch_traceback
  DevResult* d_tb_res = nullptr;
  float pack_ms = 0;
  bool downloaded = false;
  StreamSet ss; bool has_ss = false;
  bool launched = false; uint32_t launches = 0;
  // ba_align_batch_exp*: the work list of the current round (pair ids still below their target score)
  const uint32_t* d_active = nullptr; uint32_t n_active = 0; bool use_active = false;
};

static bool pow2(uint64_t x) { return x && !(x & (x - 1)); }

extern "C" const char* ba_error_string(int code) {
  switch (code) {
    case BA_OK: return "ok";
    case BA_ERR_CUDA: return "CUDA error (no usable device or a runtime failure)";
    case BA_ERR_GAPS: return "Gap costs must be negative and gap open must cost more than gap extend!";
    case BA_ERR_SIZE: return "Block sizes must be powers of two (and not exceed the supported maximum)!";
    case BA_ERR_XDROP: return "X-drop threshold amount must be nonnegative!";
    case BA_ERR_ARG: return "invalid argument";
    case BA_ERR_CHAR: return "sequence byte outside the alphabet of the scoring matrix";
    case BA_ERR_NOMEM: return "out of device memory";
    case BA_ERR_OVERFLOW: return "a per-warp trace or CIGAR buffer overflowed";
  }
  return "unknown";
}
extern "C" const char* ba_last_error_message(void) { return g_last_error.c_str(); }

extern "C" int ba_create(int device, BaAligner** out) {
  if (!out) return fail(BA_ERR_ARG, "out is null");
  BaAligner* a = new BaAligner();
  a->device = device;
#ifndef BA_EMU
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { delete a; return fail(BA_ERR_CUDA, "no CUDA device available; this library has no CPU fallback"); }
  if (device < 0 || device >= ndev) { delete a; return fail(BA_ERR_CUDA, "device index out of range"); }
  if (cudaSetDevice(device) != cudaSuccess) { delete a; return fail(BA_ERR_CUDA, "cudaSetDevice failed"); }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete a; return fail(BA_ERR_CUDA, "cudaGetDeviceProperties failed"); }
  a->sm_count = prop.multiProcessorCount;
  a->smem_optin = prop.sharedMemPerBlockOptin;
  a->mem_total = prop.totalGlobalMem;
  if (cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking) != cudaSuccess) { delete a; return fail(BA_ERR_CUDA, "cudaStreamCreate failed"); }
  cudaEventCreate(&a->ev0); cudaEventCreate(&a->ev1); cudaEventCreate(&a->ev2);
#else
  a->smem_optin = 227 * 1024;
  a->mem_total = (size_t)8 << 30;
  if (const char* e = getenv("BA_EMU_WARPS")) a->emu_warps = std::max(1, atoi(e));
#endif
  *out = a;
  return BA_OK;
}

extern "C" void ba_destroy(BaAligner* a) {
  if (!a) return;
#ifndef BA_EMU
  cudaSetDevice(a->device);
#endif
#ifndef BA_EMU
  for (auto& ss : a->streams_free) {
    cudaStreamDestroy(ss.stream); cudaEventDestroy(ss.ev0); cudaEventDestroy(ss.ev1); cudaEventDestroy(ss.evp0); cudaEventDestroy(ss.evp1);
  }
#endif
  a->streams_free.clear();
#ifdef BA_EMU
  free(a->h_stage);
#else
  if (a->h_stage) cudaFreeHost(a->h_stage);
  cudaDeviceSynchronize();
  for (int t = 0; t < BaAligner::kStageThreads; t++) for (int k = 0; k < 2; k++) {
    if (a->ring[t][k]) cudaFreeHost(a->ring[t][k]);
    if (a->ring_ev[t][k]) cudaEventDestroy(a->ring_ev[t][k]);
  }
#endif
  for (auto& e : a->pool_free) dfree(e.first);
  for (auto& e : a->pool_live) dfree(e.first);   // batches must be freed before the aligner; be forgiving
  a->pool_free.clear(); a->pool_live.clear();
#ifndef BA_EMU
  if (a->ev0) cudaEventDestroy(a->ev0);
  if (a->ev1) cudaEventDestroy(a->ev1);
  if (a->ev2) cudaEventDestroy(a->ev2);
  if (a->stream) cudaStreamDestroy(a->stream);
#endif
  delete a;
}

// Give the cached device buffers back to the driver (they are kept between batches because cudaMalloc / cudaFree of
// GB-sized buffers cost more than the copies they serve) and forget the high-water mark that sizes the cache.
extern "C" int ba_trim(BaAligner* a) {
  if (!a) return fail(BA_ERR_ARG, "aligner is null");
  AlLock lk(a->mu);
#ifndef BA_EMU
  if (cudaSetDevice(a->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
#endif
  for (auto& e : a->pool_free) dfree(e.first);
  a->pool_free.clear(); a->pool_cached = 0; a->pool_peak = a->pool_live_bytes;
  return BA_OK;
}

extern "C" void ba_batch_free(BaBatch* b) {
  if (!b) return;
  AlLock lk(b->al->mu);
#ifndef BA_EMU
  cudaSetDevice(b->al->device);
#endif
  BaAligner* al = b->al;
  void* bufs[] = {b->d_seq, b->d_qoff, b->d_roff, b->d_qlen, b->d_rlen, b->d_order, b->d_matrix, b->d_profiles, b->d_prof_arena,
                  b->d_out, b->d_ticket, b->d_ckpt, b->d_trace, b->d_rects, b->d_runs, b->d_cigar, b->d_cigar_used,
                  b->d_steplog, b->d_steplog_n, b->d_tb_res, b->d_overflow_list, b->d_overflow_n, b->d_zwords,
                  b->d_gborders, b->d_trace_pool, b->d_trace_pool_cursor, b->d_arena_of};
  for (void* q : bufs) pool_release(al, q);
  if (b->has_ss) { dsync(b->ss.stream); streams_release(al, b->ss); }
  delete b;
}

// argument rules of Block::align / align_profile (scan_block.rs:849-862, 944-954)
static int check_config(const BaConfig* cfg, uint32_t* mn, uint32_t* mx) {
  if (!cfg) return fail(BA_ERR_ARG, "cfg is null");
  if (cfg->scoring < 0 || cfg->scoring > 3) return fail(BA_ERR_ARG, "bad scoring kind");
  if (cfg->flags & ~(BA_TRACE | BA_XDROP | BA_LOCAL_START | BA_FREE_QUERY_START_GAPS | BA_FREE_QUERY_END_GAPS | BA_REV_QUERY | BA_REV_REFERENCE | BA_INPUT_NUC4))
    return fail(BA_ERR_ARG, "unsupported flags");
  if ((cfg->flags & BA_INPUT_NUC4) && cfg->scoring != BA_SCORING_NUC) return fail(BA_ERR_ARG, "BA_INPUT_NUC4 needs BA_SCORING_NUC");
  if ((cfg->flags & BA_XDROP) && (cfg->flags & BA_FREE_QUERY_END_GAPS))
    return fail(BA_ERR_ARG, "Cannot set both X_DROP and FREE_QUERY_END_GAPS!");   // scan_block.rs:861
  if ((cfg->flags & BA_LOCAL_START) && (cfg->flags & BA_FREE_QUERY_START_GAPS))
    return fail(BA_ERR_ARG, "Cannot set both LOCAL_START and FREE_QUERY_START_GAPS!");   // scan_block.rs:860
  if (cfg->scoring != BA_SCORING_PROFILE) {
    if (!cfg->matrix) return fail(BA_ERR_ARG, "matrix is null");
    if (!(cfg->gaps.open < 0 && cfg->gaps.extend < 0)) return fail(BA_ERR_GAPS, "Gap costs must be negative!");
    if (!(cfg->gaps.open < cfg->gaps.extend)) return fail(BA_ERR_GAPS, "Gap open must cost more than gap extend!");
  }
  uint64_t a = cfg->size.min < (uint64_t)kL ? kL : cfg->size.min, b = cfg->size.max < (uint64_t)kL ? kL : cfg->size.max;
  if (!(a < 65535 && b < 65535)) return fail(BA_ERR_SIZE, "Block sizes must be smaller than 2^16 - 1!");
  if (!(pow2(a) && pow2(b))) return fail(BA_ERR_SIZE, "Block sizes must be powers of two!");
  if (b > (uint64_t)kMaxBlock) return fail(BA_ERR_SIZE, "max block size above 16384 is not supported by this build");
  if ((cfg->flags & BA_XDROP) && cfg->x_drop < 0) return fail(BA_ERR_XDROP, "X-drop threshold amount must be nonnegative!");
  if (cfg->scoring == BA_SCORING_PROFILE && cfg->cigar_eq) return fail(BA_ERR_ARG, "cigar_eq needs a reference sequence");
  // our limit: the lane-ordered argmax of this mode is computed with the whole block in one 256-row chunk
  if ((cfg->flags & BA_FREE_QUERY_END_GAPS) && a > 256) return fail(BA_ERR_SIZE, "FREE_QUERY_END_GAPS supports min block sizes up to 256");
  *mn = (uint32_t)a; *mx = (uint32_t)b;
  return BA_OK;
}

// Launch geometry and per-slot scratch of a resident batch for min block size `mn` (the max size is fixed by the
// upload: sequences are padded for it). Called by the upload and again by ba_align_batch_exp* for every doubled min
// size, so whatever an earlier configuration allocated is released first.
#define CFG_TRY(x) do { int _r = (x); if (_r) return _r == 1 ? BA_ERR_CUDA : _r; } while (0)
// worst-case trace words of one alignment of total length `len` = the reference's Trace::new (scan_block.rs:1364-1369)
static uint64_t trace_bound_words(int flags, uint32_t mn, uint32_t mx, uint64_t len) {
  uint64_t words = 2 * (uint64_t)(mx / 16) * (len + 2 * (uint64_t)mx);
  if (mn == 16) words *= 2;   // 16-row rectangles still occupy a 32-lane word group
  // FREE_QUERY_END_GAPS: every rectangle is laid out with 8 rows per lane (one 256-row chunk): 32 words per column
  if (flags & BA_FREE_QUERY_END_GAPS) words = std::max<uint64_t>(words, 32 * (len + 2 * (uint64_t)mx));
  return words + 64;
}
static int batch_configure(BaBatch* b, uint32_t mn) {
  BaAligner* al = b->al;
  const BaConfig* cfg = &b->cfg;
  const uint32_t mx = b->max_size;
  const size_t n = b->n;
  const bool prof = cfg->scoring == BA_SCORING_PROFILE;
  b->min_size = mn;
  {
    void** old[] = {(void**)&b->d_ckpt, (void**)&b->d_gborders, (void**)&b->d_trace_pool, (void**)&b->d_trace_pool_cursor, (void**)&b->d_trace,
                    (void**)&b->d_zwords, (void**)&b->d_rects, (void**)&b->d_runs, (void**)&b->d_arena_of};
    for (void** q : old) { pool_release(al, *q); *q = nullptr; }
    b->trace_pool_units = 0;
  }
  // launch geometry and per-slot scratch
  // fast phase: four alignments per warp while the block sits at its minimum size (32 or 64)
  const bool ext = (cfg->flags & (BA_LOCAL_START | BA_FREE_QUERY_START_GAPS | BA_FREE_QUERY_END_GAPS)) != 0;
  b->kflags = (cfg->flags & 3) | (ext ? kExt : 0);
  // fast-phase mode (third kernel template argument): 4 / 8 = s32 rows per lane (TRACE), 16 + LGT = packed, 0 = none
  b->fast_rows = (!prof && !ext && !getenv("BA_NO_FAST")) ? (mn == 32 ? 4 : (mn == 64 ? 8 : 0)) : 0;
  if (prof && b->prof_fast) b->fast_rows = mn == 32 ? 4 : (mn == 64 ? 8 : 0);
  // the per-step debug log (BA_STEP_LOG, single pair) is only written by the generic phase: the fast step's code has to
  // stay small (instruction cache), and the sequence of steps does not depend on which phase executes them
  if (getenv("BA_STEP_LOG") && n == 1) b->fast_rows = 0;
  b->slots_per_warp = b->fast_rows ? 4 : 1;
  if (b->fast_rows && !b->pk_enable && !prof) b->fast_rows = 0;
  if (b->fast_rows) {
    const int lgt = mn == 32 ? 2 : 3;
    b->fast_rows = 16 + lgt;
    b->slots_per_warp = 32u >> lgt;
  }
  // max block >= 1024: the four live borders (8-32 KB per warp) move to global memory so that shared memory does not
  // cap the SM at 8 warps or fewer (C5: 64..=2048)
  const bool gb = b->fast_rows >= 16 && mx >= 1024 && !prof && !getenv("BA_NO_GLOBAL_BORDERS");
  if (gb) b->fast_rows += 16;
  b->gb = gb;
  const size_t wbytes = warp_smem_bytes(mx, gb);
  int wpb = 4;
  while (wpb > 1 && kSmemHeader + wpb * wbytes > al->smem_optin - 1024) wpb >>= 1;
  if (kSmemHeader + wpb * wbytes > al->smem_optin) { return fail(BA_ERR_SIZE, "max block size does not fit in shared memory"); }
  b->wpb = wpb; b->smem_bytes = kSmemHeader + wpb * wbytes;
  int bps = 1;
  CFG_TRY(ba_occupancy_dispatch(prof ? (int)kProfile : cfg->scoring, b->kflags, b->fast_rows, wpb, b->smem_bytes, &bps));
  if (bps < 1) bps = 1;
  uint64_t max_blocks = (uint64_t)al->sm_count * bps;
#ifdef BA_EMU
  max_blocks = al->emu_warps; b->wpb = wpb = 1;
#endif
  const bool trace = (cfg->flags & BA_TRACE) != 0;
  const size_t ms = mx < 32 ? 32 : mx;
  const uint64_t spw = b->slots_per_warp;
  b->max_blocks_hw = max_blocks;
  uint64_t want = (n + wpb * spw - 1) / (wpb * spw);
  if (want < 1) want = 1;
  b->blocks = (int)std::min<uint64_t>(want, max_blocks);
  const uint64_t nwarps = (uint64_t)b->blocks * wpb;
  const uint64_t nslots = nwarps * spw;
  CFG_TRY(pool_alloc(al, (void**)&b->d_ckpt, nslots * 4 * ms * sizeof(int16_t)));
  if (b->gb) CFG_TRY(pool_alloc(al, (void**)&b->d_gborders, nwarps * 4 * ms * sizeof(int16_t)));
  if (trace) {
    // Trace memory. Every pair of a launch owns one arena (trace words + rectangle records + run scratch), indexed by
    // arena_of[pair], so the traces of a whole launch are resident when its alignment kernel ends and the traceback
    // runs as a second kernel with one walk per lane. The worst case per alignment is the reference's Trace::new
    // (scan_block.rs:1364-1369: the block sits at its maximum size all the time); real alignments spend most steps
    // at the minimum size, so the first pass gives every pair twice its all-minimum-size path and lets the rare
    // rectangle that does not fit borrow from a batch-wide overflow pool (trace_push). Pairs that still overflow are
    // re-run with worst-case arenas (batch_wait). The processing order is longest-first, so a wave (the pairs of one
    // launch) is a run of similar lengths and uniform arenas sized by its first pair waste little.
    const uint64_t zmul = (cfg->flags & BA_LOCAL_START) ? 2 : 1;
    b->mem_budget = (uint64_t)(al->mem_total * al->budget_frac);
    const bool want_pool = !(cfg->flags & BA_LOCAL_START) && !getenv("BA_NO_TRACE_POOL") && !getenv("BA_TRACE_WORST_CASE") &&
                           !(cfg->flags & BA_FREE_QUERY_END_GAPS);
    auto pair_len = [&](uint32_t t) { return (uint64_t)b->h_qlen[b->h_order[t]] + b->h_rlen[b->h_order[t]]; };
    b->waves.clear();
    b->arena_forced = getenv("BA_TRACE_ARENA_WORDS") != nullptr;
    uint64_t spill_max = 0;
    for (uint32_t t = 0; t < n;) {
      const uint64_t len = (((pair_len(t) >> 6) + 1) << 6) + 2;      // the order is sorted in buckets of 64
      const uint64_t bound = trace_bound_words(cfg->flags, mn, mx, len);
      if (bound >= ((uint64_t)1 << 32)) return fail(BA_ERR_SIZE, "trace arena of one alignment exceeds 2^32 words");
      uint64_t words = 2 * len * (uint64_t)std::max<uint32_t>(mn, 32) / 8 + 4096 + (want_pool ? 0 : (uint64_t)mx * mx / 4);
      if (!want_pool && (getenv("BA_TRACE_WORST_CASE") || (cfg->flags & BA_FREE_QUERY_END_GAPS))) words = bound;
      if (const char* e = getenv("BA_TRACE_ARENA_WORDS")) words = std::max<uint64_t>(64, (uint64_t)atoll(e));   // tests: tiny arenas
      words = std::min(words, bound);
      BaBatch::Wave wv;
      wv.t0 = t; wv.words = words;
      wv.rects = (uint32_t)std::min<uint64_t>(words < bound ? len / 4 + 1024 : len * 2 + 8, 0x7fffffffu);
      wv.runs = (uint32_t)std::min<uint64_t>(len + 8, 0x7fffffffu);
      const uint64_t per = zmul * words * 4 + (uint64_t)wv.rects * sizeof(Rect) + (uint64_t)wv.runs * 4;
      const uint64_t room = b->mem_budget - (want_pool ? b->mem_budget / 10 : 0);
      wv.n = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n - t, room / std::max<uint64_t>(per, 1)));
      if (b->waves.empty() && wv.n == n && !b->arena_forced && words < bound) {
        // one wave holds everything: the arenas take what is left of the budget, up to 4x the all-minimum-size path
        // (an overflow retry costs far more than the memory)
        const uint64_t fixed = (uint64_t)n * ((uint64_t)wv.rects * sizeof(Rect) + (uint64_t)wv.runs * 4);
        const uint64_t cap = std::min<uint64_t>(bound, 4 * len * (uint64_t)std::max<uint32_t>(mn, 32) / 8 + 4096);
        if (room > fixed) wv.words = std::max<uint64_t>(words, std::min<uint64_t>(cap, (room - fixed) / ((uint64_t)n * zmul * 4)));
      }
      wv.words = (wv.words + 31) & ~(uint64_t)31;    // arenas start on 128-byte lines (16-byte asynchronous copies of the batch traceback)
      spill_max = std::max<uint64_t>(spill_max, (uint64_t)wv.n * (bound - std::min(bound, wv.words)));
      b->cap_trace = std::max<uint64_t>(b->cap_trace, (uint64_t)wv.n * wv.words * 4);
      b->cap_rects = std::max<uint64_t>(b->cap_rects, (uint64_t)wv.n * wv.rects * sizeof(Rect));
      b->cap_runs = std::max<uint64_t>(b->cap_runs, (uint64_t)wv.n * wv.runs * 4);
      b->waves.push_back(wv);
      t += wv.n;
    }
    b->h_arena_of.assign(n, 0);
    for (const BaBatch::Wave& wv : b->waves)
      for (uint32_t t = 0; t < wv.n; t++) b->h_arena_of[b->h_order[wv.t0 + t]] = t;
    if (want_pool && spill_max) {
      // a tenth of the budget, but never more than the alignments of a wave could spill: a one-pair Part 1 call must
      // not allocate (and then cache) gigabytes
      uint64_t pool_bytes = std::min<uint64_t>(b->mem_budget / 10, spill_max * 4 + 4096);
#ifdef BA_EMU
      pool_bytes = std::min<uint64_t>(pool_bytes, (uint64_t)8 << 20);
#endif
      if (const char* e = getenv("BA_TRACE_POOL_BYTES")) pool_bytes = (uint64_t)atoll(e);   // tests: force exhaustion
      b->trace_pool_units = pool_bytes / 64;
      if (b->trace_pool_units) {
        CFG_TRY(pool_alloc(al, (void**)&b->d_trace_pool, b->trace_pool_units * 64));
        CFG_TRY(pool_alloc(al, (void**)&b->d_trace_pool_cursor, 4));
      }
    }
    CFG_TRY(pool_alloc(al, (void**)&b->d_trace, b->cap_trace));
    if (cfg->flags & BA_LOCAL_START) CFG_TRY(pool_alloc(al, (void**)&b->d_zwords, b->cap_trace));
    CFG_TRY(pool_alloc(al, (void**)&b->d_rects, b->cap_rects));
    CFG_TRY(pool_alloc(al, (void**)&b->d_runs, b->cap_runs));
    CFG_TRY(pool_alloc(al, (void**)&b->d_arena_of, std::max<size_t>(n, 1) * 4));
    CFG_TRY(h2d(b->d_arena_of, b->h_arena_of.data(), n * 4, b->ss.stream));
    CFG_TRY(dsync(b->ss.stream));
    if (!b->d_cigar) {
      uint64_t cap = 0;
      for (size_t k = 0; k < n; k++) cap += (uint64_t)b->h_qlen[k] + b->h_rlen[k] + 5;
      const uint64_t limit = (uint64_t)(al->mem_total * 0.15) / 4;
      b->cigar_cap = std::min<uint64_t>(cap, std::max<uint64_t>(limit, 1024));
      CFG_TRY(pool_alloc(al, (void**)&b->d_cigar, b->cigar_cap * 4));
      CFG_TRY(pool_alloc(al, (void**)&b->d_cigar_used, 8));
      CFG_TRY(pool_alloc(al, (void**)&b->d_overflow_list, std::max<size_t>(n, 1) * 4));
      CFG_TRY(pool_alloc(al, (void**)&b->d_overflow_n, 4));
    }
    b->resident.assign(n, 0);
    if (!b->waves.empty()) b->cur = b->waves[0];
  }
  if (getenv("BA_STEP_LOG") && n == 1 && !b->d_steplog) {
    CFG_TRY(pool_alloc(al, (void**)&b->d_steplog, (size_t)(1 << 20) * sizeof(StepLog)));
    CFG_TRY(pool_alloc(al, (void**)&b->d_steplog_n, 4));
  }
  return BA_OK;
}

#define TRY(x) do { int _r = (x); if (_r) { ba_batch_free(b); return _r == 1 ? BA_ERR_CUDA : _r; } } while (0)

static int upload_common(BaAligner* al, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                         const uint8_t* r_bytes, const uint64_t* r_off, const AAProfile* const* profiles, const BaPssmBatch* pssm,
                         BaBatch** out) {
  if (!al || !out) return fail(BA_ERR_ARG, "null aligner/out");
  AlLock lk(al->mu);
  uint32_t mn = 0, mx = 0;
  int rc = check_config(cfg, &mn, &mx);
  if (rc) return rc;
  // The reference does not check min <= max (its scratch is sized by max, so min > max walks off the end of it); here
  // that is an argument error.
  if (mn > mx) return fail(BA_ERR_SIZE, "min block size is larger than max block size");
  const bool prof = cfg->scoring == BA_SCORING_PROFILE;
  if (prof != (profiles != nullptr || pssm != nullptr)) return fail(BA_ERR_ARG, "profiles must be given exactly for BA_SCORING_PROFILE");
  if (pssm) {
    if (!pssm->order || pssm->order_len == 0 || pssm->order_len > 32) return fail(BA_ERR_ARG, "BaPssmBatch: order must hold 1..32 residues");
    if (n && (!pssm->scores || !pssm->score_off)) return fail(BA_ERR_ARG, "BaPssmBatch: null scores");
    if (pssm->gap_extend >= 0) return fail(BA_ERR_GAPS, "Gap extend cost must be negative!");
    const bool per_pos = pssm->gap_open_C || pssm->gap_close_C || pssm->gap_open_R;
    if (per_pos && !(pssm->gap_open_C && pssm->gap_close_C && pssm->gap_open_R && pssm->gap_off))
      return fail(BA_ERR_ARG, "BaPssmBatch: give all three per-position gap arrays and gap_off, or none");
    if (!per_pos && (pssm->all_gap_open_C >= 0 || pssm->all_gap_open_R >= 0)) return fail(BA_ERR_GAPS, "Gap open cost must be negative!");
  }
  if (n >= ((size_t)1 << 31)) return fail(BA_ERR_ARG, "too many pairs");
  if (n && (!q_off || (!prof && !r_off))) return fail(BA_ERR_ARG, "null offsets");
#ifndef BA_EMU
  if (cudaSetDevice(al->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
#endif
  const bool timing = getenv("BA_TIMING") != nullptr;
  const auto tu0 = std::chrono::steady_clock::now();
  BaBatch* b = new BaBatch();
  b->al = al; b->cfg = *cfg; b->n = n; b->min_size = mn; b->max_size = mx;
  if (streams_acquire(al, &b->ss)) { delete b; return BA_ERR_CUDA; }
  b->has_ss = true;
  dev_stream_t st = b->ss.stream;
  const uint32_t pad = mx + 32;

  // host pass: lengths, padded offsets, processing order (longest first)
  std::vector<uint64_t> pq(n), pr(n);
  std::vector<uint32_t> ql(n), rl(n), order(n);
  uint64_t pos = 0, max_len = 0;
  for (size_t k = 0; k < n; k++) {
    const uint64_t a = q_off[k + 1] - q_off[k];
    uint64_t c;
    if (pssm) {
      const uint64_t ns = pssm->score_off[k + 1] - pssm->score_off[k];
      if (ns % pssm->order_len) { ba_batch_free(b); return fail(BA_ERR_ARG, "BaPssmBatch: scores of a profile are not a multiple of order_len"); }
      c = ns / pssm->order_len;
    } else if (prof) {
      if (!profiles[k]) { ba_batch_free(b); return fail(BA_ERR_ARG, "null profile"); }
      c = host::profile_len(profiles[k]);
      if (host::profile_gap_extend(profiles[k]) >= 0) { ba_batch_free(b); return fail(BA_ERR_GAPS, "Gap extend cost must be negative!"); }
      if (host::profile_curr_len(profiles[k]) < c + mx + 1) { ba_batch_free(b); return fail(BA_ERR_SIZE, "profile was created with a smaller block size than max block size"); }
    } else {
      c = r_off[k + 1] - r_off[k];
    }
    if (a >= ((uint64_t)1 << 31) || c >= ((uint64_t)1 << 31)) { ba_batch_free(b); return fail(BA_ERR_ARG, "sequence too long"); }
    if ((cfg->flags & BA_FREE_QUERY_END_GAPS) && !(mn > a)) {   // scan_block.rs:862
      ba_batch_free(b); return fail(BA_ERR_SIZE, "Min block size must be larger than the query length for FREE_QUERY_END_GAPS!");
    }
    ql[k] = (uint32_t)a; rl[k] = (uint32_t)c;
    pq[k] = pos; pos += (1 + a + pad + 15) & ~(uint64_t)15;
    if (!prof) { pr[k] = pos; pos += (1 + c + pad + 15) & ~(uint64_t)15; }
    max_len = std::max<uint64_t>(max_len, a + c);
  }
  b->max_pair_len = (uint32_t)std::min<uint64_t>(max_len, 0xffffffffu);
  {  // counting sort by total length, descending (long alignments first -> short tail)
    const int SH = 6;
    const size_t nb = (size_t)(max_len >> SH) + 2;
    std::vector<uint32_t> cnt(nb + 1, 0);
    for (size_t k = 0; k < n; k++) cnt[nb - 1 - (((uint64_t)ql[k] + rl[k]) >> SH)]++;
    uint32_t run = 0;
    for (size_t i = 0; i < nb; i++) { uint32_t c = cnt[i]; cnt[i] = run; run += c; }
    for (size_t k = 0; k < n; k++) order[cnt[nb - 1 - (((uint64_t)ql[k] + rl[k]) >> SH)]++] = (uint32_t)k;
  }
  const uint64_t seq_bytes = pos + 512;   // slack: the fast phase touches up to B + 136 bytes past the position it reads (pk_fast_step)
  const auto tu1 = std::chrono::steady_clock::now();

  // BA_INPUT_NUC4: offsets count nibbles; the bytes that hold nibbles [off[0], off[n]) are copied and the offsets are
  // rebased to the first copied byte (off[0] rounded down to even)
  const bool nuc4 = (cfg->flags & BA_INPUT_NUC4) != 0;
  const uint64_t qb0 = n ? (nuc4 ? q_off[0] >> 1 : q_off[0]) : 0, rb0 = (!prof && n) ? (nuc4 ? r_off[0] >> 1 : r_off[0]) : 0;
  const uint64_t qraw = n ? (nuc4 ? ((q_off[n] + 1) >> 1) - qb0 : q_off[n] - q_off[0]) : 0;
  const uint64_t rraw = (!prof && n) ? (nuc4 ? ((r_off[n] + 1) >> 1) - rb0 : r_off[n] - r_off[0]) : 0;
  uint8_t *d_rawq = nullptr, *d_rawr = nullptr; uint64_t *d_rawqoff = nullptr, *d_rawroff = nullptr; uint32_t* d_err = nullptr;
  auto free_tmp = [&]() { pool_release(al, d_rawq); pool_release(al, d_rawr); pool_release(al, d_rawqoff); pool_release(al, d_rawroff); pool_release(al, d_err); };
#define TRY2(x) do { int _r = (x); if (_r) { free_tmp(); ba_batch_free(b); return _r == 1 ? BA_ERR_CUDA : _r; } } while (0)
  TRY2(pool_alloc(al, (void**)&b->d_seq, seq_bytes));
  TRY2(pool_alloc(al, (void**)&b->d_qoff, n * 8)); TRY2(pool_alloc(al, (void**)&b->d_qlen, n * 4));
  TRY2(pool_alloc(al, (void**)&b->d_roff, n * 8)); TRY2(pool_alloc(al, (void**)&b->d_rlen, n * 4));
  TRY2(pool_alloc(al, (void**)&b->d_order, n * 4));
  TRY2(pool_alloc(al, (void**)&b->d_out, n * sizeof(DevResult)));
  TRY2(pool_alloc(al, (void**)&b->d_ticket, 4));
  TRY2(pool_alloc(al, (void**)&d_rawq, qraw)); TRY2(pool_alloc(al, (void**)&d_rawqoff, (n + 1) * 8));
  if (!prof) { TRY2(pool_alloc(al, (void**)&d_rawr, rraw)); TRY2(pool_alloc(al, (void**)&d_rawroff, (n + 1) * 8)); }
  TRY2(pool_alloc(al, (void**)&d_err, 4));
  TRY2(dzero(d_err, 4, st));
  const auto tu2 = std::chrono::steady_clock::now();
  // offsets are rebased so that the raw arenas start at 0
  std::vector<uint64_t> qo(n + 1), ro(n + 1);
  for (size_t k = 0; k <= n && n; k++) {
    qo[k] = q_off[k] - (nuc4 ? qb0 * 2 : q_off[0]);
    if (!prof) ro[k] = r_off[k] - (nuc4 ? rb0 * 2 : r_off[0]);
  }
  if (n) {
    TRY2(h2d_seq(al, d_rawq, q_bytes + qb0, qraw, st)); TRY2(h2d(d_rawqoff, qo.data(), (n + 1) * 8, st));
    if (!prof) { TRY2(h2d_seq(al, d_rawr, r_bytes + rb0, rraw, st)); TRY2(h2d(d_rawroff, ro.data(), (n + 1) * 8, st)); }
    TRY2(h2d(b->d_qoff, pq.data(), n * 8, st)); TRY2(h2d(b->d_qlen, ql.data(), n * 4, st));
    if (!prof) TRY2(h2d(b->d_roff, pr.data(), n * 8, st));
    TRY2(h2d(b->d_rlen, rl.data(), n * 4, st));
    TRY2(h2d(b->d_order, order.data(), n * 4, st));
  }
  // scoring data
  if (!prof) {
    const size_t mb = cfg->scoring == BA_SCORING_NUC ? 128 : (cfg->scoring == BA_SCORING_AA ? 864 : 2);
    int smax = 0;
    for (size_t i = 0; i < mb; i++) smax = std::max(smax, (int)((const int8_t*)cfg->matrix)[i]);
    b->pk_smax = smax;
    b->pk_enable = getenv("BA_NO_PACKED") ? 0u : 1u;
    TRY2(pool_alloc(al, (void**)&b->d_matrix, mb));
    TRY2(h2d(b->d_matrix, cfg->matrix, mb, st));
  } else {
    // One device arena: per profile pos_aa [curr_len][32] then three i16 arrays [curr_len]. Only the positions that
    // can differ from AAProfile::new's defaults cross the bus (a host AAProfile's leading positions, or the raw
    // PSSM rows of a BaPssmBatch); ba_profile_build_kernel expands them and writes the -128 padding on the device.
    std::vector<ProfileDev> pd(n);
    std::vector<ProfBuild> bd(n);
    // The packed fast phase serves profile batches without TRACE / extended modes at min block 32 / 64 when every
    // profile of the batch has the same gap_extend (the constants of the packed scan are per batch); it needs the
    // derived layouts (ProfileDev::tp / gp), built on the device behind the base arrays.
    bool prof_fast = !(cfg->flags & (BA_TRACE | BA_LOCAL_START | BA_FREE_QUERY_START_GAPS | BA_FREE_QUERY_END_GAPS)) && (mn == 32 || mn == 64) &&
                     !getenv("BA_NO_FAST") && !getenv("BA_NO_PACKED") && n > 0;
    if (prof_fast && !pssm) {
      const int ge0 = host::profile_gap_extend(profiles[0]);
      for (size_t k = 1; k < n && prof_fast; k++) prof_fast = host::profile_gap_extend(profiles[k]) == ge0;
    }
    b->prof_fast = prof_fast;
    b->prof_ge = n ? (pssm ? (int)pssm->gap_extend : host::profile_gap_extend(profiles[0])) : -1;
    uint64_t ppos = 0, spos = 0;
    for (size_t k = 0; k < n; k++) {
      const uint64_t cl = pssm ? (uint64_t)rl[k] + mx + 1 : host::profile_curr_len(profiles[k]);
      bd[k].dst_off = ppos; bd[k].curr_len = (uint32_t)cl;
      bd[k].gap_extend = pssm ? (int32_t)pssm->gap_extend : host::profile_gap_extend(profiles[k]);
      bd[k].derive = prof_fast ? 1u : 0u;
      ppos += prof_slot_bytes(cl, prof_fast);
      if (pssm) {
        bd[k].np = rl[k];
        bd[k].src_off = pssm->score_off[k] - pssm->score_off[0];
        bd[k].gap_src_off = pssm->gap_off ? pssm->gap_off[k] - pssm->gap_off[0] : 0;
        if (pssm->gap_off && pssm->gap_off[k + 1] - pssm->gap_off[k] != (uint64_t)rl[k] + 1) {
          free_tmp(); ba_batch_free(b); return fail(BA_ERR_ARG, "BaPssmBatch: gap arrays need len + 1 entries per profile");
        }
      } else {
        const uint64_t np = host::profile_used_len(profiles[k]);
        bd[k].np = (uint32_t)np; bd[k].src_off = spos; bd[k].gap_src_off = 0;
        spos += (np * 38 + 63) & ~(uint64_t)63;
      }
    }
    TRY2(pool_alloc(al, (void**)&b->d_prof_arena, ppos + 64));
    for (size_t k = 0; k < n; k++) {
      const uint64_t cl = bd[k].curr_len;
      uint8_t* dbase = b->d_prof_arena + bd[k].dst_off;
      pd[k].pos_aa = (const int8_t*)dbase;
      pd[k].gap_open_C = (const int16_t*)(dbase + cl * 32);
      pd[k].gap_close_C = (const int16_t*)(dbase + cl * 34);
      pd[k].gap_open_R = (const int16_t*)(dbase + cl * 36);
      pd[k].len = rl[k];
      pd[k].gap_extend = bd[k].gap_extend;
      pd[k].tp = prof_fast ? (const int8_t*)(dbase + prof_tp_off(cl)) : nullptr;
      pd[k].gp = prof_fast ? (const uint32_t*)(dbase + prof_gp_off(cl)) : nullptr;
      pd[k].tlen = (uint32_t)prof_tlen(cl); pd[k].pad_ = 0;
    }
    ProfBuildArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.n = (uint32_t)n; ba.arena = b->d_prof_arena; ba.err = d_err;
    ProfBuild* d_bd = nullptr; uint8_t* d_src = nullptr; int8_t* d_gaps = nullptr;
    auto free_prof_tmp = [&]() { pool_release(al, d_bd); pool_release(al, d_src); pool_release(al, d_gaps); };
#define TRY3(x) do { int _r = (x); if (_r) { free_prof_tmp(); free_tmp(); ba_batch_free(b); return _r == 1 ? BA_ERR_CUDA : _r; } } while (0)
    TRY3(pool_alloc(al, (void**)&d_bd, std::max<size_t>(n, 1) * sizeof(ProfBuild)));
    TRY3(h2d(d_bd, bd.data(), n * sizeof(ProfBuild), st));
    ba.desc = d_bd;
    if (pssm) {
      ba.kind = 1;
      const uint64_t sbytes = n ? pssm->score_off[n] - pssm->score_off[0] : 0;
      TRY3(pool_alloc(al, (void**)&d_src, sbytes));
      if (n) TRY3(h2d_seq(al, d_src, (const uint8_t*)(pssm->scores + pssm->score_off[0]), sbytes, st));   // pageable caller memory: threaded staging copy
      ba.scores = (const int8_t*)d_src;
      if (pssm->gap_off) {
        const uint64_t gb = n ? pssm->gap_off[n] - pssm->gap_off[0] : 0;
        TRY3(pool_alloc(al, (void**)&d_gaps, 3 * gb));
        if (n) {
          TRY3(h2d(d_gaps, pssm->gap_open_C + pssm->gap_off[0], gb, st));
          TRY3(h2d(d_gaps + gb, pssm->gap_close_C + pssm->gap_off[0], gb, st));
          TRY3(h2d(d_gaps + 2 * gb, pssm->gap_open_R + pssm->gap_off[0], gb, st));
        }
        ba.gap_open_C = d_gaps; ba.gap_close_C = d_gaps + gb; ba.gap_open_R = d_gaps + 2 * gb;
      }
      ba.all_open_C = pssm->all_gap_open_C; ba.all_close_C = pssm->all_gap_close_C; ba.all_open_R = pssm->all_gap_open_R;
      ba.order_len = (uint32_t)pssm->order_len;
      ba.left_shift = (uint32_t)pssm->left_shift & 7u; ba.right_shift = (uint32_t)pssm->right_shift & 7u;
      ba.rev = pssm->rev ? 1u : 0u;
      for (int e = 0; e < 32; e++) ba.inv[e] = -1;
      for (size_t j = 0; j < pssm->order_len; j++) {
        const uint8_t c = host::upper(pssm->order[j]);
        if (!(c >= 'A' && c <= 'Z' + 1)) { free_prof_tmp(); free_tmp(); ba_batch_free(b); return fail(BA_ERR_CHAR, "BaPssmBatch: order byte out of range"); }
        ba.inv[c - 'A'] = (int8_t)j;
      }
    } else {
      ba.kind = 0;
      if (stage_reserve(al, spos + 64)) { free_prof_tmp(); free_tmp(); ba_batch_free(b); return fail(BA_ERR_NOMEM, "pinned staging allocation failed"); }
      uint8_t* stage = al->h_stage;
      auto export_range = [&](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; k++) {
          const size_t np = bd[k].np;
          uint8_t* sb = stage + bd[k].src_off;
          host::profile_export(profiles[k], np, (int8_t*)sb, (int16_t*)(sb + np * 32), (int16_t*)(sb + np * 34), (int16_t*)(sb + np * 36));
        }
      };
      const size_t nthr = spos >= ((size_t)8 << 20) ? std::min<size_t>(8, std::max(1u, std::thread::hardware_concurrency())) : 1;
      if (nthr <= 1) export_range(0, n);
      else {
        std::vector<std::thread> th;
        for (size_t t = 0; t < nthr; t++) th.emplace_back(export_range, n * t / nthr, n * (t + 1) / nthr);
        for (auto& t : th) t.join();
      }
      TRY3(pool_alloc(al, (void**)&d_src, spos + 64));
      TRY3(h2d(d_src, stage, spos, st));
      ba.stage = d_src;
    }
    if (n) {
#ifdef BA_EMU
      prof_build_all(ba);
#else
      const int pblocks = (int)std::min<uint64_t>(n, (uint64_t)al->sm_count * 16);
      ba_profile_build_kernel<<<pblocks, 128, 0, st>>>(ba);
      if (cudaGetLastError() != cudaSuccess) { free_prof_tmp(); free_tmp(); ba_batch_free(b); return fail(BA_ERR_CUDA, "profile build kernel launch failed"); }
#endif
    }
    TRY3(pool_alloc(al, (void**)&b->d_profiles, n * sizeof(ProfileDev)));
    TRY3(h2d(b->d_profiles, pd.data(), n * sizeof(ProfileDev), st));
    TRY3(dsync(st));     // the staging buffers (host and device) are reused by the next upload
    free_prof_tmp();
    uint32_t perr = 0;
    TRY2(d2h(&perr, d_err, 4, st));
    TRY2(dsync(st));
    if (perr) { free_tmp(); ba_batch_free(b); return fail(BA_ERR_GAPS, "Gap open cost must be negative!"); }
  }
  // convert + pad on the device
  PackArgs pa;
  pa.raw_q = d_rawq; pa.raw_q_off = d_rawqoff; pa.raw_r = d_rawr; pa.raw_r_off = d_rawroff;
  pa.seq = b->d_seq; pa.pq_off = b->d_qoff; pa.pr_off = b->d_roff; pa.n = (uint32_t)n; pa.pad = pad;
  pa.scoring = prof ? (int)kAA : cfg->scoring; pa.err = d_err;
  pa.rev_q = (cfg->flags & BA_REV_QUERY) ? 1u : 0u; pa.rev_r = (cfg->flags & BA_REV_REFERENCE) ? 1u : 0u;
  pa.nuc4 = nuc4 ? 1u : 0u;
  if (n) {
#ifdef BA_EMU
    pack_all(pa);
#else
    const int threads = 256;
    const uint64_t nseq = prof ? n : 2 * n;
    const int blocks = (int)std::min<uint64_t>((nseq * 32 + threads - 1) / threads, (uint64_t)al->sm_count * 16);
    CUDA_OK(cudaEventRecord(b->ss.evp0, st));
    ba_pack_kernel<<<blocks, threads, 0, st>>>(pa);
    if (cudaGetLastError() != cudaSuccess) { free_tmp(); ba_batch_free(b); return fail(BA_ERR_CUDA, "pack kernel launch failed"); }
    CUDA_OK(cudaEventRecord(b->ss.evp1, st));
#endif
  }
  uint32_t herr = 0;
  TRY2(d2h(&herr, d_err, 4, st));
  TRY2(dsync(st));
#ifndef BA_EMU
  if (n) cudaEventElapsedTime(&b->pack_ms, b->ss.evp0, b->ss.evp1);
#endif
  const auto tu3 = std::chrono::steady_clock::now();
  free_tmp();
  const auto tu4 = std::chrono::steady_clock::now();
  if (timing) {
    auto d = [](std::chrono::steady_clock::time_point x, std::chrono::steady_clock::time_point y) { return std::chrono::duration<double, std::milli>(y - x).count(); };
    fprintf(stderr, "upload: host pass %.1f ms, cudaMalloc %.1f ms, h2d+pack %.1f ms, free tmp %.1f ms\n", d(tu0, tu1), d(tu1, tu2), d(tu2, tu3), d(tu3, tu4));
  }
  if (herr) { ba_batch_free(b); return fail(BA_ERR_CHAR, "sequence byte outside the alphabet of the scoring matrix"); }

  b->h_qlen = ql; b->h_rlen = rl; b->h_order = order;
  TRY(batch_configure(b, mn));
  *out = b;
  return BA_OK;
}

extern "C" int ba_batch_upload(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                               const uint8_t* r_bytes, const uint64_t* r_off, BaBatch** out) {
  if (cfg && cfg->scoring == BA_SCORING_PROFILE) return fail(BA_ERR_ARG, "use ba_batch_upload_profiles");
  return upload_common(a, cfg, n, q_bytes, q_off, r_bytes, r_off, nullptr, nullptr, out);
}
extern "C" int ba_batch_upload_profiles(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                        const AAProfile* const* profiles, BaBatch** out) {
  if (cfg && cfg->scoring != BA_SCORING_PROFILE) return fail(BA_ERR_ARG, "scoring must be BA_SCORING_PROFILE");
  if (!profiles) return fail(BA_ERR_ARG, "profiles is null");
  return upload_common(a, cfg, n, q_bytes, q_off, nullptr, nullptr, profiles, nullptr, out);
}
extern "C" int ba_batch_upload_pssm(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                    const BaPssmBatch* pssm, BaBatch** out) {
  if (cfg && cfg->scoring != BA_SCORING_PROFILE) return fail(BA_ERR_ARG, "scoring must be BA_SCORING_PROFILE");
  if (!pssm) return fail(BA_ERR_ARG, "pssm is null");
  return upload_common(a, cfg, n, q_bytes, q_off, nullptr, nullptr, nullptr, pssm, out);
}

static Params make_params(const BaBatch* b) {
  Params P;
  memset(&P, 0, sizeof(P));
  const bool prof = b->cfg.scoring == BA_SCORING_PROFILE;
  P.n_pairs = (uint32_t)b->n; P.order = b->d_order; P.seq = b->d_seq;
  if (b->use_active) { P.n_pairs = b->n_active; P.order = b->d_active; }
  P.q_off = b->d_qoff; P.q_len = b->d_qlen; P.r_off = b->d_roff; P.r_len = b->d_rlen;
  P.profiles = b->d_profiles; P.matrix = b->d_matrix;
  P.gap_open = b->cfg.gaps.open; P.gap_extend = b->cfg.gaps.extend;
  if (prof) { P.gap_open = 0; P.gap_extend = b->prof_ge; }   // per-position gap opens; the (uniform) extend cost of the fast phase
  P.min_size = b->min_size; P.max_size = b->max_size; P.x_drop = b->cfg.x_drop;
  P.flags = b->kflags; P.scoring = prof ? (int)kProfile : b->cfg.scoring;
  P.pk_smax = b->pk_smax; P.pk_enable = b->pk_enable;
  {
    auto pk2h = [](int v) { return ((uint32_t)v & 0xffffu) | ((uint32_t)v << 16); };
    const int go = prof ? 0 : b->cfg.gaps.open, ge = prof ? b->prof_ge : b->cfg.gaps.extend;
    P.kc.ge2 = pk2h(ge); P.kc.go1 = (uint32_t)go * 65537u; P.kc.or2 = pk2h(go - ge);
    for (int k = 0; k < 4; k++) P.kc.kge[k] = pk2h((k + 1) * ge);
    for (int s2 = 0; s2 < 5; s2++) P.kc.dec[s2] = pk2h((4 << s2) * ge);
    P.kc.lane1 = (uint32_t)(4 * ge) * 65537u;
    P.kc.ge1 = (uint32_t)ge * 65537u;
  }
  P.ext_flags = (uint32_t)(b->cfg.flags & (BA_LOCAL_START | BA_FREE_QUERY_START_GAPS | BA_FREE_QUERY_END_GAPS)); P.trace_zwords = b->d_zwords;
  P.out = b->d_out; P.ticket = b->d_ticket;
  P.ckpt = b->d_ckpt; P.slots_per_warp = b->slots_per_warp; P.gborders = b->d_gborders;
  P.fast_block = b->fast_rows >= 16 ? (8u << (b->fast_rows & 15)) : (uint32_t)(8 * b->fast_rows);
  P.pk_fast = b->fast_rows >= 16 ? 1u : 0u;
  P.trace_words = b->d_trace; P.trace_words_per_warp = b->cur.words; P.arena_of = b->d_arena_of;
  P.trace_pool = b->d_trace_pool; P.trace_pool_cursor = b->d_trace_pool_cursor; P.trace_pool_units = b->trace_pool_units;
  P.rects = b->d_rects; P.rects_per_warp = b->cur.rects;
  P.run_scratch = b->d_runs; P.runs_per_warp = b->cur.runs;
  P.cigar_stream = b->d_cigar; P.cigar_cap = b->cigar_cap; P.cigar_used = b->d_cigar_used;
  P.cigar_eq = b->cfg.cigar_eq ? 1u : 0u;
  P.overflow_list = b->d_overflow_list; P.overflow_n = b->d_overflow_n;
  P.step_log = b->d_steplog; P.step_log_cap = b->d_steplog ? (1u << 20) : 0u; P.step_log_n = b->d_steplog_n;
  return P;
}

// One launch of the alignment kernel over a work list, followed -- for TRACE batches -- by the traceback kernel over
// the same list (the list's traces are resident in their arenas then). Stream-ordered, no host synchronisation.
static int launch_wave(BaBatch* b, Params& P, const uint32_t* list, uint32_t n_list, const BaBatch::Wave& wv, int blocks) {
  dev_stream_t st = b->ss.stream;
  const bool trace = (b->cfg.flags & BA_TRACE) != 0;
  P.order = list; P.n_pairs = n_list;
  if (trace) { P.trace_words_per_warp = wv.words; P.rects_per_warp = wv.rects; P.runs_per_warp = wv.runs; b->cur = wv; }
  if (!n_list) return BA_OK;
  if (dzero(b->d_ticket, 4, st)) return BA_ERR_CUDA;
  if (b->d_trace_pool_cursor && dzero(b->d_trace_pool_cursor, 4, st)) return BA_ERR_CUDA;
  int rc = ba_launch_dispatch(P.scoring, P.flags, b->fast_rows, P, blocks, b->wpb, b->smem_bytes, st);
  if (rc) return rc == 1 ? BA_ERR_CUDA : rc;
  b->launches++;
  if (trace && P.cigar_stream) {
    if (launch_traceback_batch(P, list, n_list, st)) return BA_ERR_CUDA;
    b->launches++;
  }
  return BA_OK;
}

// Launch a resident batch on its own stream (no host synchronisation): one wave, or -- TRACE batches whose arenas do not
// fit the memory budget at once -- several in a row.
static int batch_launch(BaBatch* b) {
  if (!b) return fail(BA_ERR_ARG, "batch is null");
  BaAligner* al = b->al;
  AlLock lk(al->mu);
  dev_stream_t st = b->ss.stream;
#ifndef BA_EMU
  if (cudaSetDevice(al->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
#endif
  Params P = make_params(b);
  if (b->d_cigar_used && dzero(b->d_cigar_used, 8, st)) return BA_ERR_CUDA;
  if (b->d_overflow_n && dzero(b->d_overflow_n, 4, st)) return BA_ERR_CUDA;
  if (b->d_steplog_n && dzero(b->d_steplog_n, 4, st)) return BA_ERR_CUDA;
  b->downloaded = false;
  b->launches = 0;
  const bool trace = (b->cfg.flags & BA_TRACE) != 0;
  const uint32_t n_run = b->use_active ? b->n_active : (uint32_t)b->n;
#ifndef BA_EMU
  if (n_run) cudaEventRecord(b->ss.ev0, st);
#endif
  if (!trace || b->use_active) {
    BaBatch::Wave none;
    int rc = launch_wave(b, P, b->use_active ? b->d_active : b->d_order, n_run, none, b->blocks);
    if (rc) return rc;
  } else {
    if (b->h_arena_of_dirty) {     // a retry pass of the previous run re-assigned arenas
      for (const BaBatch::Wave& wv : b->waves)
        for (uint32_t t = 0; t < wv.n; t++) b->h_arena_of[b->h_order[wv.t0 + t]] = t;
      if (h2d(b->d_arena_of, b->h_arena_of.data(), b->n * 4, st) || dsync(st)) return BA_ERR_CUDA;
      b->h_arena_of_dirty = false;
    }
    std::fill(b->resident.begin(), b->resident.end(), 0);
    for (const BaBatch::Wave& wv : b->waves) {
      int rc = launch_wave(b, P, b->d_order + wv.t0, wv.n, wv, b->blocks);
      if (rc) return rc;
    }
    if (!b->waves.empty()) {
      const BaBatch::Wave& wv = b->waves.back();
      for (uint32_t t = 0; t < wv.n; t++) b->resident[b->h_order[wv.t0 + t]] = 1;
    }
  }
  b->launched = true;
  return BA_OK;
}

// Wait for a launched batch; re-run alignments whose (deliberately small) trace arena overflowed with worst-case arenas
// -- in the same buffers, as many at a time as fit --; report the device time between the first launch and the last
// completion.
static int batch_wait(BaBatch* b, BaStats* stats) {
  if (!b || !b->launched) return fail(BA_ERR_ARG, "batch was not launched");
  BaAligner* al = b->al;
  AlLock lk(al->mu);
#ifndef BA_EMU
  if (cudaSetDevice(al->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
#endif
  dev_stream_t st = b->ss.stream;
  float ms = 0;
  if (b->use_active ? b->n_active : b->n) {
    if (b->d_overflow_n && !b->use_active) {
      uint32_t n_over = 0;
      if (d2h(&n_over, b->d_overflow_n, 4, st) || dsync(st)) return BA_ERR_CUDA;
      if (getenv("BA_TIMING")) fprintf(stderr, "batch_wait: %zu pairs, %d blocks x %d warps, %u slots/warp, %zu wave(s), trace words/arena %llu, %u overflowed\n",
                                       b->n, b->blocks, b->wpb, b->slots_per_warp, b->waves.size(), (unsigned long long)b->cur.words, n_over);
      if (n_over) {
        std::vector<uint32_t> over(n_over);
        if (d2h(over.data(), b->d_overflow_list, (size_t)n_over * 4, st) || dsync(st)) return BA_ERR_CUDA;
        const uint64_t zmul = b->d_zwords ? 2 : 1;
        uint64_t len = 0;
        for (uint32_t k : over) len = std::max<uint64_t>(len, (uint64_t)b->h_qlen[k] + b->h_rlen[k] + 2);
        BaBatch::Wave wv;
        wv.words = trace_bound_words(b->cfg.flags, b->min_size, b->max_size, len);
        wv.rects = (uint32_t)std::min<uint64_t>(len * 2 + 8, 0x7fffffffu);
        wv.runs = (uint32_t)std::min<uint64_t>(len + 8, 0x7fffffffu);
        // the worst-case arenas live in the first pass's buffers (grown if not even one fits)
        auto fit = [&]() { return std::min<uint64_t>(std::min<uint64_t>(b->cap_trace / (wv.words * 4), b->cap_rects / ((uint64_t)wv.rects * sizeof(Rect))),
                                                     b->cap_runs / ((uint64_t)wv.runs * 4)); };
        if (fit() == 0) {
          void** bufs[] = {(void**)&b->d_trace, (void**)&b->d_zwords, (void**)&b->d_rects, (void**)&b->d_runs};
          const bool had_z = b->d_zwords != nullptr;
          for (void** q : bufs) { pool_release(al, *q); *q = nullptr; }
          b->cap_trace = std::max<uint64_t>(b->cap_trace, wv.words * 4);
          b->cap_rects = std::max<uint64_t>(b->cap_rects, (uint64_t)wv.rects * sizeof(Rect));
          b->cap_runs = std::max<uint64_t>(b->cap_runs, (uint64_t)wv.runs * 4);
          if (pool_alloc(al, (void**)&b->d_trace, b->cap_trace)) return BA_ERR_NOMEM;
          if (had_z && pool_alloc(al, (void**)&b->d_zwords, b->cap_trace)) return BA_ERR_NOMEM;
          if (pool_alloc(al, (void**)&b->d_rects, b->cap_rects)) return BA_ERR_NOMEM;
          if (pool_alloc(al, (void**)&b->d_runs, b->cap_runs)) return BA_ERR_NOMEM;
        }
        (void)zmul;
        const uint64_t per_wave = std::max<uint64_t>(1, fit());
        Params P2 = make_params(b);
        P2.overflow_list = nullptr; P2.overflow_n = nullptr;     // cannot overflow with worst-case arenas
        std::fill(b->resident.begin(), b->resident.end(), 0);
        for (uint32_t t0 = 0; t0 < n_over; t0 += (uint32_t)per_wave) {
          wv.t0 = t0; wv.n = (uint32_t)std::min<uint64_t>(per_wave, n_over - t0);
          for (uint32_t t = 0; t < wv.n; t++) b->h_arena_of[over[t0 + t]] = t;
          if (h2d(b->d_arena_of, b->h_arena_of.data(), b->n * 4, st) || dsync(st)) return BA_ERR_CUDA;
          b->h_arena_of_dirty = true;
          const uint64_t spw = b->slots_per_warp;
          const int blocks2 = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)b->blocks, (wv.n + b->wpb * spw - 1) / (b->wpb * spw)));
          int rc = launch_wave(b, P2, b->d_overflow_list + t0, wv.n, wv, blocks2);
          if (rc) return rc;
          if (t0 + wv.n >= n_over) for (uint32_t t = 0; t < wv.n; t++) b->resident[over[t0 + t]] = 1;
        }
      }
    }
#ifndef BA_EMU
    cudaEventRecord(b->ss.ev1, st);
    cudaError_t e = cudaEventSynchronize(b->ss.ev1);
    if (e != cudaSuccess) { cuda_fail(e, "alignment kernel"); return BA_ERR_CUDA; }
    cudaEventElapsedTime(&ms, b->ss.ev0, b->ss.ev1);
#endif
  }
  b->launched = false;
  if (stats) {
    memset(stats, 0, sizeof(*stats));
    stats->kernel_ms = ms; stats->pack_ms = b->pack_ms; stats->kernel_launches = b->launches;
  }
  return BA_OK;
}

extern "C" int ba_batch_run(BaBatch* b, BaStats* stats) {
  int rc = batch_launch(b);
  if (rc) return rc;
  return batch_wait(b, stats);
}

extern "C" int ba_batch_download(BaBatch* b, AlignResult* out) {
  if (!b) return fail(BA_ERR_ARG, "batch is null");
  BaAligner* al = b->al;
  AlLock lk(al->mu);
  dev_stream_t st = b->ss.stream;
#ifndef BA_EMU
  if (cudaSetDevice(al->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
#endif
  b->h_out.resize(b->n);
  if (d2h(b->h_out.data(), b->d_out, b->n * sizeof(DevResult), st)) return BA_ERR_CUDA;
  unsigned long long used = 0;
  if (b->d_cigar_used && d2h(&used, b->d_cigar_used, 8, st)) return BA_ERR_CUDA;
  if (dsync(st)) return BA_ERR_CUDA;
  if (b->d_cigar && !b->cigar_skip) {
    used = std::min<unsigned long long>(used, b->cigar_cap);
    if (b->cigar_dst) {     // straight into the caller's buffer (no host copy of the stream is kept)
      if (used > b->cigar_dst_cap) return fail(BA_ERR_OVERFLOW, "caller's CIGAR buffer is too small");
      if (d2h(b->cigar_dst, b->d_cigar, used * 4, st)) return BA_ERR_CUDA;
      b->cigar_dst_used = used;
    } else {
      b->h_cigar.resize(used);
      if (d2h(b->h_cigar.data(), b->d_cigar, used * 4, st)) return BA_ERR_CUDA;
    }
    if (dsync(st)) return BA_ERR_CUDA;
  }
  int bad = 0;
  for (size_t k = 0; k < b->n; k++) {
    const DevResult& r = b->h_out[k];
    if (out) { out[k].score = r.score; out[k].query_idx = r.query_idx; out[k].reference_idx = r.reference_idx; }
    if (r.status != kOk) bad++;
  }
  b->downloaded = true;
  if (bad) return fail(BA_ERR_OVERFLOW, "a per-warp trace or CIGAR buffer overflowed for " + std::to_string(bad) + " pair(s)");
  return BA_OK;
}

extern "C" int ba_batch_cigar(const BaBatch* b, size_t k, const uint32_t** runs, size_t* n_runs) {
  if (!b || !b->downloaded || k >= b->n || !(b->cfg.flags & BA_TRACE) || b->cigar_dst) return fail(BA_ERR_ARG, "no CIGAR available");
  const DevResult& r = b->h_out[k];
  if (r.cigar_off + r.cigar_n > b->h_cigar.size()) return fail(BA_ERR_OVERFLOW, "cigar stream overflow");
  *runs = b->h_cigar.data() + r.cigar_off; *n_runs = r.cigar_n;
  return BA_OK;
}

extern "C" int ba_batch_pair_stats(const BaBatch* b, size_t k, uint64_t* cells, uint32_t* steps, uint32_t* status) {
  if (!b || !b->downloaded || k >= b->n) return fail(BA_ERR_ARG, "no stats available");
  if (cells) *cells = b->h_out[k].cells;
  if (steps) *steps = b->h_out[k].steps;
  if (status) *status = b->h_out[k].status;
  return BA_OK;
}

#ifdef BA_EMU
extern "C" void ba_emu_stats(uint64_t* out4, int reset) {
  out4[0] = emu_stats::pk_cells; out4[1] = emu_stats::exact_cells; out4[2] = emu_stats::fast_steps; out4[3] = emu_stats::big_cells;
  if (getenv("BA_EMU_TB_STATS")) fprintf(stderr, "traceback cell steps: %llu staged, %llu from global memory\n", (unsigned long long)emu_stats::tb_staged, (unsigned long long)emu_stats::tb_global);
  if (reset) { emu_stats::tb_staged = 0; emu_stats::tb_global = 0; emu_stats::pk_cells = 0; emu_stats::exact_cells = 0; emu_stats::fast_steps = 0; emu_stats::big_cells = 0; }
}
#endif

// debug: fetch the per-step log of a single-pair batch run with BA_STEP_LOG=1
extern "C" size_t ba_debug_step_log(BaBatch* b, StepLog* out, size_t cap) {
  if (!b || !b->d_steplog) return 0;
  AlLock lk(b->al->mu);
  uint32_t n = 0;
  d2h(&n, b->d_steplog_n, 4, b->ss.stream); dsync(b->ss.stream);
  const size_t m = std::min<size_t>(std::min<size_t>(n, cap), 1u << 20);
  if (m) { d2h(out, b->d_steplog, m * sizeof(StepLog), b->ss.stream); dsync(b->ss.stream); }
  return n;
}

// upload + run + download. Large batches are cut into chunks that are pipelined: every chunk has its own
// stream, so the H2D copy (and convert/pad) of chunk k+1 overlaps the alignment kernel of chunk k, and the
// tail of kernel k overlaps the head of kernel k+1.
struct CigarOut { uint32_t* runs; size_t cap; uint64_t* off; uint32_t* len; size_t used; };
static int align_batch_impl(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                            const uint8_t* r_bytes, const uint64_t* r_off, AlignResult* out, BaStats* stats, CigarOut* cg);
extern "C" int ba_align_batch(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                              const uint8_t* r_bytes, const uint64_t* r_off, AlignResult* out, BaStats* stats) {
  return align_batch_impl(a, cfg, n, q_bytes, q_off, r_bytes, r_off, out, stats, nullptr);
}
// ba_align_batch for BA_TRACE batches with the CIGARs delivered into the caller's buffer: pair k's runs are
// runs[run_off[k] .. run_off[k] + run_len[k]), packed (len << 4) | Operation, forward order.
extern "C" int ba_align_batch_cigar(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                    const uint8_t* r_bytes, const uint64_t* r_off, AlignResult* out,
                                    uint32_t* runs, size_t runs_cap, uint64_t* run_off, uint32_t* run_len, size_t* runs_used,
                                    BaStats* stats) {
  if (!cfg || !(cfg->flags & BA_TRACE)) return fail(BA_ERR_ARG, "ba_align_batch_cigar needs BA_TRACE");
  if (n && (!runs || !run_off || !run_len)) return fail(BA_ERR_ARG, "null CIGAR output");
  CigarOut cg{runs, runs_cap, run_off, run_len, 0};
  const int rc = align_batch_impl(a, cfg, n, q_bytes, q_off, r_bytes, r_off, out, stats, &cg);
  if (runs_used) *runs_used = cg.used;
  return rc;
}
static int align_batch_impl(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                            const uint8_t* r_bytes, const uint64_t* r_off, AlignResult* out, BaStats* stats, CigarOut* cg) {
  if (!a) return fail(BA_ERR_ARG, "aligner is null");
  if (n && (!q_off || !r_off)) return fail(BA_ERR_ARG, "null offsets");
  AlLock lk(a->mu);
  const bool timing = getenv("BA_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  size_t K = 1;
  const uint64_t bytes = n ? (q_off[n] - q_off[0]) + (r_off[n] - r_off[0]) : 0;
  if (n >= 8192 && bytes >= ((uint64_t)32 << 20)) K = 4;
  // Large batches: geometric chunk sizes 1/16, 1/16, 1/8, 1/4, 1/2 -- a small first chunk starts the GPU early, large late
  // chunks keep the per-kernel load balance. Measured on C2 (2.1 GB, kernel alone 82.7 ms): 4 equal chunks 105.9 ms,
  // 8 equal 100.2 ms, 5 geometric 94.6 ms.
  bool geom = false;
  if (n >= 65536 && bytes >= ((uint64_t)512 << 20)) { K = 5; geom = true; }
  // mid-size batches (the shards of a multi-GPU call): 1/8, 1/8, 1/4, 1/2 -- 25 k pairs of C2: 28.8 -> 27.6 ms
  else if (n >= 16384 && bytes >= ((uint64_t)128 << 20)) { K = 4; geom = true; }
  if (const char* e = getenv("BA_PIPELINE_CHUNKS")) K = std::max<size_t>(1, std::min<size_t>((size_t)atoi(e), 64));
  // TRACE: every chunk in flight owns trace arenas and the memory budget is split between them. Their kernels cannot
  // overlap anyway (one chunk fills the GPU), so pipelining only hides the H2D copy; it is given up when halving the
  // budget would shrink the first-pass arenas (an overflow retry costs far more than the copy).
  if (cfg && (cfg->flags & BA_TRACE) && K > 2) K = 2;
  if (cfg && (cfg->flags & BA_TRACE) && K == 2 && n) {
    uint64_t max_len = 0;
    for (size_t k = 0; k < n; k++) max_len = std::max<uint64_t>(max_len, (q_off[k + 1] - q_off[k]) + (r_off[k + 1] - r_off[k]));
    const uint64_t mnb = std::max<uint64_t>(cfg->size.min, 32), mxb = cfg->size.max;
    const uint64_t want_words = 4 * (max_len + 2) * mnb / 8 + mxb * mxb / 4 + 4096;
    const uint64_t slots = (uint64_t)a->sm_count * 16 * 8;       // upper bound of the alignments in flight
    if (want_words * 4 * slots > (uint64_t)(a->mem_total * a->budget_frac / 2)) K = 1;
  }
  if (K > n) K = n ? n : 1;
  const double budget_saved = a->budget_frac;
  a->budget_frac = budget_saved / (double)K;
  // chunk boundaries with roughly equal input bytes
  std::vector<size_t> cut(K + 1, 0);
  cut[K] = n;
  if (const char* e = getenv("BA_PIPELINE_GEOM")) geom = atoi(e) != 0;
  if (K <= 1) geom = false;
  for (size_t c = 1, k = 0; c < K; c++) {
    const uint64_t target = geom ? (bytes >> (K - c)) : bytes * c / K;
    while (k < n && (q_off[k] - q_off[0]) + (r_off[k] - r_off[0]) < target) k++;
    cut[c] = k;
  }
  std::vector<BaBatch*> bs(K, nullptr);
  int rc = BA_OK;
  BaStats tot;
  memset(&tot, 0, sizeof(tot));
  for (size_t c = 0; c < K && !rc; c++) {
    const size_t lo = cut[c], hi = cut[c + 1];
    rc = ba_batch_upload(a, cfg, hi - lo, q_bytes, q_off + lo, r_bytes, r_off + lo, &bs[c]);
    if (!rc) rc = batch_launch(bs[c]);
  }
  for (size_t c = 0; c < K; c++) {
    if (!bs[c]) continue;
    BaStats st1;
    if (!rc) rc = batch_wait(bs[c], &st1);
    if (!rc && cg) { bs[c]->cigar_dst = cg->runs + cg->used; bs[c]->cigar_dst_cap = cg->cap - cg->used; }
    if (!rc && !cg) bs[c]->cigar_skip = true;
    if (!rc) rc = ba_batch_download(bs[c], out ? out + cut[c] : nullptr);
    if (!rc && cg) {
      for (size_t k = 0; k < bs[c]->n; k++) {
        cg->off[cut[c] + k] = cg->used + bs[c]->h_out[k].cigar_off;
        cg->len[cut[c] + k] = bs[c]->h_out[k].cigar_n;
      }
      cg->used += bs[c]->cigar_dst_used;
    }
    if (!rc) {
      tot.kernel_ms += st1.kernel_ms; tot.pack_ms += st1.pack_ms; tot.kernel_launches += st1.kernel_launches + (bs[c]->n ? 1 : 0);
      for (size_t k = 0; k < bs[c]->n; k++) {
        tot.cells += bs[c]->h_out[k].cells; tot.steps += bs[c]->h_out[k].steps;
        if (bs[c]->h_out[k].status) tot.n_failed++;
      }
    }
    ba_batch_free(bs[c]);
  }
  a->budget_frac = budget_saved;
  if (stats) *stats = tot;
  if (timing) fprintf(stderr, "ba_align_batch: %zu chunk(s), %.1f ms wall\n", K,
                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  return rc;
}

// Block::align_exp / align_profile_exp (scan_block.rs:884-902, 974-992) for a batch: every pair is aligned with min
// block size s = max(min, 16), 2s, ... <= max until its score reaches target_score[k]. min_size_used[k] = the size that
// reached it, 0 = the reference's None; out[k] = the result of the pair's last attempt (what Block::res() returns
// afterwards). The batch is uploaded ONCE and stays resident: after every round a device-side filter compacts the
// ids of the pairs still below their target into the next round's work list (order preserved, so the longest-first
// schedule survives), and only the length of that list crosses the bus.
struct ExpArgs {
  const DevResult* out; const int32_t* target; const uint32_t* in_list; uint32_t n_in;
  uint32_t* out_list; uint32_t* out_n; uint32_t* used; uint32_t sz;
  unsigned long long* totals;    // [0] cells, [1] steps, [2] failed pairs of all rounds
};
#ifdef BA_EMU
static void exp_filter(const ExpArgs& a, dev_stream_t) {
  uint32_t m = 0;
  for (uint32_t i = 0; i < a.n_in; i++) {
    const uint32_t pair = a.in_list[i];
    const DevResult& r = a.out[pair];
    a.totals[0] += r.cells; a.totals[1] += r.steps; a.totals[2] += r.status ? 1 : 0;
    if (r.score >= a.target[pair]) a.used[pair] = a.sz; else a.out_list[m++] = pair;
  }
  *a.out_n = m;
}
#else
__global__ void __launch_bounds__(1024) ba_exp_filter_kernel(ExpArgs a) {
  __shared__ uint32_t wsum[32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t base = 0, failed = 0;
  unsigned long long cells = 0, steps = 0;
  for (uint32_t i0 = 0; i0 < a.n_in; i0 += 1024) {
    const uint32_t i = i0 + threadIdx.x;
    bool keep = false; uint32_t pair = 0;
    if (i < a.n_in) {
      pair = a.in_list[i];
      const DevResult r = a.out[pair];
      cells += r.cells; steps += r.steps; failed += r.status ? 1u : 0u;
      if (r.score >= a.target[pair]) a.used[pair] = a.sz; else keep = true;
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) wsum[warp] = (uint32_t)__popc(m);
    __syncthreads();
    uint32_t woff = 0, tot = 0;
    for (uint32_t w = 0; w < 32; w++) { const uint32_t c = wsum[w]; if (w < warp) woff += c; tot += c; }
    if (keep) a.out_list[base + woff + (uint32_t)__popc(m & ((1u << lane) - 1u))] = pair;
    base += tot;
    __syncthreads();
  }
  atomicAdd(a.totals + 0, cells); atomicAdd(a.totals + 1, steps); atomicAdd(a.totals + 2, (unsigned long long)failed);
  if (threadIdx.x == 0) *a.out_n = base;
}
static void exp_filter(const ExpArgs& a, dev_stream_t st) { ba_exp_filter_kernel<<<1, 1024, 0, st>>>(a); }
#endif

static int align_exp_resident(BaBatch* b, size_t n, const int32_t* target_score, AlignResult* out, uintptr_t* min_size_used, BaStats* stats) {
  BaAligner* al = b->al;
  dev_stream_t st = b->ss.stream;
  const uint32_t mn = b->min_size, mx = b->max_size;
  int32_t* d_target = nullptr; uint32_t* d_used = nullptr; uint32_t* d_list[2] = {nullptr, nullptr}; uint32_t* d_n = nullptr;
  unsigned long long* d_tot = nullptr;
  auto cleanup = [&]() { pool_release(al, d_target); pool_release(al, d_used); pool_release(al, d_list[0]); pool_release(al, d_list[1]);
                         pool_release(al, d_n); pool_release(al, d_tot); };
#define EXP_TRY(x) do { int _r = (x); if (_r) { cleanup(); ba_batch_free(b); return _r == 1 ? BA_ERR_CUDA : _r; } } while (0)
  EXP_TRY(pool_alloc(al, (void**)&d_target, n * 4)); EXP_TRY(pool_alloc(al, (void**)&d_used, n * 4));
  EXP_TRY(pool_alloc(al, (void**)&d_list[0], n * 4)); EXP_TRY(pool_alloc(al, (void**)&d_list[1], n * 4));
  EXP_TRY(pool_alloc(al, (void**)&d_n, 4)); EXP_TRY(pool_alloc(al, (void**)&d_tot, 24));
  EXP_TRY(h2d(d_target, target_score, n * 4, st)); EXP_TRY(dzero(d_used, n * 4, st)); EXP_TRY(dzero(d_tot, 24, st));
  BaStats tot; memset(&tot, 0, sizeof(tot));
  const uint32_t* cur = b->d_order; uint32_t n_cur = (uint32_t)n;
  int flip = 0;
  for (uint64_t sz = mn; sz <= mx && n_cur; sz *= 2) {
    if (sz != mn) EXP_TRY(batch_configure(b, (uint32_t)sz));
    b->use_active = true; b->d_active = cur; b->n_active = n_cur;
    BaStats st1;
    EXP_TRY(batch_launch(b));
    EXP_TRY(batch_wait(b, &st1));
    tot.kernel_ms += st1.kernel_ms; tot.kernel_launches += st1.kernel_launches + 1;
    ExpArgs ea{b->d_out, d_target, cur, n_cur, d_list[flip], d_n, d_used, (uint32_t)sz, d_tot};
    exp_filter(ea, st);
    uint32_t n_next = 0;
    EXP_TRY(d2h(&n_next, d_n, 4, st)); EXP_TRY(dsync(st));
    cur = d_list[flip]; n_cur = n_next; flip ^= 1;
  }
  b->use_active = false;
  std::vector<uint32_t> used(n);
  unsigned long long totals[3] = {0, 0, 0};
  EXP_TRY(d2h(used.data(), d_used, n * 4, st)); EXP_TRY(d2h(totals, d_tot, 24, st)); EXP_TRY(dsync(st));
  for (size_t k = 0; k < n; k++) min_size_used[k] = used[k];
  cleanup();
  b->cigar_skip = true;
  int rc = ba_batch_download(b, out);
  tot.cells = totals[0]; tot.steps = totals[1]; tot.n_failed = (uint32_t)totals[2]; tot.pack_ms = b->pack_ms;
  tot.kernel_launches += n ? 1 : 0;   // convert/pad
  if (stats) *stats = tot;
  ba_batch_free(b);
  return rc;
#undef EXP_TRY
}
static int exp_check(const BaConfig* cfg, const int32_t* target_score, AlignResult* out, uintptr_t* min_size_used) {
  if (!cfg || !target_score || !out || !min_size_used) return fail(BA_ERR_ARG, "null argument");
  uint32_t mn = 0, mx = 0;
  int rc = check_config(cfg, &mn, &mx);
  if (rc) return rc;
  if (mn > mx) return fail(BA_ERR_SIZE, "min block size is larger than max block size");
  return BA_OK;
}
extern "C" int ba_align_batch_exp(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                  const uint8_t* r_bytes, const uint64_t* r_off, const int32_t* target_score,
                                  AlignResult* out, uintptr_t* min_size_used, BaStats* stats) {
  int rc = exp_check(cfg, target_score, out, min_size_used);
  if (rc) return rc;
  if (!a) return fail(BA_ERR_ARG, "aligner is null");
  AlLock lk(a->mu);
  BaBatch* b = nullptr;
  BaConfig c2 = *cfg; c2.flags &= ~BA_TRACE;     // only results are returned: the trace is not needed (it does not change them)
  rc = ba_batch_upload(a, &c2, n, q_bytes, q_off, r_bytes, r_off, &b);
  if (rc) return rc;
  return align_exp_resident(b, n, target_score, out, min_size_used, stats);
}
// Block::align_profile_exp (scan_block.rs:974-992) for a batch, profiles as host AAProfile objects ...
extern "C" int ba_align_batch_exp_profiles(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                           const AAProfile* const* profiles, const int32_t* target_score,
                                           AlignResult* out, uintptr_t* min_size_used, BaStats* stats) {
  int rc = exp_check(cfg, target_score, out, min_size_used);
  if (rc) return rc;
  if (!a) return fail(BA_ERR_ARG, "aligner is null");
  AlLock lk(a->mu);
  BaBatch* b = nullptr;
  BaConfig c2 = *cfg; c2.flags &= ~BA_TRACE;
  rc = ba_batch_upload_profiles(a, &c2, n, q_bytes, q_off, profiles, &b);
  if (rc) return rc;
  return align_exp_resident(b, n, target_score, out, min_size_used, stats);
}
// ... and as raw PSSM rows the library turns into profiles on the device
extern "C" int ba_align_batch_exp_pssm(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                       const BaPssmBatch* pssm, const int32_t* target_score,
                                       AlignResult* out, uintptr_t* min_size_used, BaStats* stats) {
  int rc = exp_check(cfg, target_score, out, min_size_used);
  if (rc) return rc;
  if (!a) return fail(BA_ERR_ARG, "aligner is null");
  AlLock lk(a->mu);
  BaBatch* b = nullptr;
  BaConfig c2 = *cfg; c2.flags &= ~BA_TRACE;
  rc = ba_batch_upload_pssm(a, &c2, n, q_bytes, q_off, pssm, &b);
  if (rc) return rc;
  return align_exp_resident(b, n, target_score, out, min_size_used, stats);
}

// sequence-to-profile counterpart of ba_align_batch (no chunking: profiles are uploaded as one arena)
static int align_uploaded_profiles(BaBatch* b, size_t n, AlignResult* out, BaStats* stats);
extern "C" int ba_align_batch_profiles(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                       const AAProfile* const* profiles, AlignResult* out, BaStats* stats) {
  BaBatch* b = nullptr;
  int rc = ba_batch_upload_profiles(a, cfg, n, q_bytes, q_off, profiles, &b);
  if (rc) return rc;
  return align_uploaded_profiles(b, n, out, stats);
}
// the same for profiles built on the device from raw PSSM rows (BaPssmBatch). Large batches are cut into chunks on separate
// streams like ba_align_batch: the upload + profile construction of chunk k + 1 overlaps the alignment kernel of chunk k.
extern "C" int ba_align_batch_pssm(BaAligner* a, const BaConfig* cfg, size_t n, const uint8_t* q_bytes, const uint64_t* q_off,
                                   const BaPssmBatch* pssm, AlignResult* out, BaStats* stats) {
  if (!a || !pssm) return fail(BA_ERR_ARG, "null aligner / pssm");
  AlLock lk(a->mu);
  size_t K = 1;
  const uint64_t bytes = (n && pssm->score_off) ? pssm->score_off[n] - pssm->score_off[0] : 0;
  if (n >= 8192 && bytes >= ((uint64_t)32 << 20)) K = 4;
  if (const char* e = getenv("BA_PIPELINE_CHUNKS")) K = std::max<size_t>(1, std::min<size_t>((size_t)atoi(e), 64));
  if (K > n) K = n ? n : 1;
  if (K == 1) {
    BaBatch* b = nullptr;
    int rc = ba_batch_upload_pssm(a, cfg, n, q_bytes, q_off, pssm, &b);
    if (rc) return rc;
    return align_uploaded_profiles(b, n, out, stats);
  }
  std::vector<size_t> cut(K + 1, n);
  cut[0] = 0;
  for (size_t c = 1, k = 0; c < K; c++) {
    const uint64_t target = bytes * c / K;
    while (k < n && pssm->score_off[k] - pssm->score_off[0] < target) k++;
    cut[c] = k;
  }
  std::vector<BaBatch*> bs(K, nullptr);
  int rc = BA_OK;
  BaStats tot; memset(&tot, 0, sizeof(tot));
  for (size_t c = 0; c < K && !rc; c++) {
    const size_t lo = cut[c], hi = cut[c + 1];
    BaPssmBatch p = *pssm;
    p.score_off = pssm->score_off + lo;
    if (pssm->gap_off) p.gap_off = pssm->gap_off + lo;
    rc = ba_batch_upload_pssm(a, cfg, hi - lo, q_bytes, q_off + lo, &p, &bs[c]);
    if (!rc) rc = batch_launch(bs[c]);
  }
  for (size_t c = 0; c < K; c++) {
    if (!bs[c]) continue;
    BaStats st1;
    if (!rc) rc = batch_wait(bs[c], &st1);
    if (!rc) rc = ba_batch_download(bs[c], out ? out + cut[c] : nullptr);
    if (!rc) {
      tot.kernel_ms += st1.kernel_ms; tot.pack_ms += st1.pack_ms; tot.kernel_launches += st1.kernel_launches + (bs[c]->n ? 2 : 0);
      for (size_t k = 0; k < bs[c]->n; k++) {
        tot.cells += bs[c]->h_out[k].cells; tot.steps += bs[c]->h_out[k].steps;
        if (bs[c]->h_out[k].status) tot.n_failed++;
      }
    }
    ba_batch_free(bs[c]);
  }
  if (stats) *stats = tot;
  return rc;
}
static int align_uploaded_profiles(BaBatch* b, size_t n, AlignResult* out, BaStats* stats) {
  int rc = ba_batch_run(b, stats);
  if (!rc && stats && n) stats->kernel_launches += 2;   // convert/pad + profile build kernels of the upload
  if (!rc) rc = ba_batch_download(b, out);
  if (!rc && stats) {
    for (size_t k = 0; k < n; k++) { stats->cells += b->h_out[k].cells; stats->steps += b->h_out[k].steps; if (b->h_out[k].status) stats->n_failed++; }
  }
  ba_batch_free(b);
  return rc;
}

// -------------------------------------------------------------------------------------------------
// In-library multi-GPU entry points (SURVEY.md 8b "Needed extension" / 8e): one call, one batch, several GPUs of the
// box. Pairs are independent (Block::align touches only its own scratch, scan_block.rs:847-878), so the batch is cut
// into one contiguous shard per device with equal sum(|q| + |r|) -- the block aligner's work per pair is proportional
// to its path length times the block sizes it visits --, every shard runs on its own host thread through the same
// pipelined path as ba_align_batch (own aligner, own streams), and each thread writes its results straight into the
// caller's arrays at the shard's offset: original order, no gather step, no collective.
// -------------------------------------------------------------------------------------------------
static std::mutex g_multi_mu;
static std::vector<BaAligner*> g_multi;      // process-wide aligners of the multi-GPU entry points, index = device

static BaAligner* multi_aligner(int dev, int* rc) {
  std::lock_guard<std::mutex> lk(g_multi_mu);
  if (dev < 0) { *rc = fail(BA_ERR_ARG, "negative device index"); return nullptr; }
  if ((size_t)dev >= g_multi.size()) g_multi.resize((size_t)dev + 1, nullptr);
  if (!g_multi[dev]) { *rc = ba_create(dev, &g_multi[dev]); if (*rc) return nullptr; }
  *rc = BA_OK;
  return g_multi[dev];
}
extern "C" int ba_device_count(void) {
#ifdef BA_EMU
  const char* e = getenv("BA_EMU_DEVICES");
  return e ? std::max(1, atoi(e)) : 1;
#else
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
#endif
}
// destroys the aligners the multi-GPU entry points created (they are otherwise kept for the life of the process)
extern "C" void ba_multi_release(void) {
  std::lock_guard<std::mutex> lk(g_multi_mu);
  for (BaAligner* a : g_multi) if (a) ba_destroy(a);
  g_multi.clear();
}

struct MultiShard { size_t lo = 0, hi = 0; int dev = 0; int rc = 0; BaStats st; std::string err; };

// cut [0, n) into `parts` contiguous ranges of equal sum(|q| + |r| + 64)
static std::vector<size_t> multi_cuts(size_t n, const uint64_t* q_off, const uint64_t* r_off, size_t parts) {
  std::vector<size_t> cut(parts + 1, n);
  cut[0] = 0;
  auto work = [&](size_t k) { return (q_off[k] - q_off[0]) + (r_off ? r_off[k] - r_off[0] : 0) + 64 * (uint64_t)k; };
  const uint64_t total = n ? work(n) : 0;
  for (size_t c = 1, k = 0; c < parts; c++) {
    const uint64_t target = total / parts * c;
    while (k < n && work(k) < target) k++;
    cut[c] = k;
  }
  return cut;
}

template <class F, class G>
static int multi_run(const int* devices, int n_dev, size_t n, const uint64_t* q_off, const uint64_t* r_off, BaStats* stats,
                     std::vector<MultiShard>& sh, F body, G prepare) {
  const int avail = ba_device_count();
  if (n_dev <= 0) n_dev = avail;
  if (n_dev <= 0) return fail(BA_ERR_CUDA, "no CUDA device available; this library has no CPU fallback");
  if (n && !q_off) return fail(BA_ERR_ARG, "null offsets");
  sh.assign((size_t)n_dev, MultiShard());
  std::vector<BaAligner*> als((size_t)n_dev, nullptr);
  for (int d = 0; d < n_dev; d++) {
    sh[d].dev = devices ? devices[d] : d;
    for (int e = 0; e < d; e++) if (sh[e].dev == sh[d].dev) return fail(BA_ERR_ARG, "device listed twice");
    int rc = 0;
    als[d] = multi_aligner(sh[d].dev, &rc);
    if (rc) return rc;
  }
  const std::vector<size_t> cut = multi_cuts(n, q_off, r_off, (size_t)n_dev);
  for (int d = 0; d < n_dev; d++) { sh[d].lo = cut[d]; sh[d].hi = cut[d + 1]; }
  prepare(sh);
  auto worker = [&](int d) {
    memset(&sh[d].st, 0, sizeof(BaStats));
    if (sh[d].hi == sh[d].lo) { sh[d].rc = BA_OK; return; }
    sh[d].rc = body(als[d], sh[d], d);
    if (sh[d].rc) sh[d].err = g_last_error;
  };
  std::vector<std::thread> th;
  for (int d = 1; d < n_dev; d++) th.emplace_back(worker, d);
  worker(0);
  for (auto& t : th) t.join();
  BaStats tot; memset(&tot, 0, sizeof(tot));
  int rc = BA_OK;
  for (int d = 0; d < n_dev; d++) {
    if (sh[d].rc && !rc) { rc = sh[d].rc; g_last_error = "device " + std::to_string(sh[d].dev) + ": " + sh[d].err; }
    tot.cells += sh[d].st.cells; tot.steps += sh[d].st.steps; tot.kernel_launches += sh[d].st.kernel_launches; tot.n_failed += sh[d].st.n_failed;
    tot.kernel_ms = std::max(tot.kernel_ms, sh[d].st.kernel_ms); tot.pack_ms = std::max(tot.pack_ms, sh[d].st.pack_ms);
  }
  if (stats) *stats = tot;
  return rc;
}

extern "C" int ba_align_batch_multi(const int* devices, int n_dev, const BaConfig* cfg, size_t n,
                                    const uint8_t* q_bytes, const uint64_t* q_off, const uint8_t* r_bytes, const uint64_t* r_off,
                                    AlignResult* out, BaStats* stats) {
  if (n && !r_off) return fail(BA_ERR_ARG, "null offsets");
  if (!n) { if (stats) memset(stats, 0, sizeof(*stats)); return BA_OK; }
  std::vector<MultiShard> sh;
  return multi_run(devices, n_dev, n, q_off, r_off, stats, sh, [&](BaAligner* a, MultiShard& s, int) {
    return align_batch_impl(a, cfg, s.hi - s.lo, q_bytes, q_off + s.lo, r_bytes, r_off + s.lo, out ? out + s.lo : nullptr, &s.st, nullptr);
  }, [](std::vector<MultiShard>&) {});
}

// CIGARs included (BA_TRACE): like ba_align_batch_cigar, but run_off[k] is an absolute index into `runs` and the shards'
// regions of `runs` are not contiguous with each other -- each device gets a slice of the buffer in proportion to its
// shard's worst case sum(|q| + |r| + 5). *runs_used = words written in total.
extern "C" int ba_align_batch_multi_cigar(const int* devices, int n_dev, const BaConfig* cfg, size_t n,
                                          const uint8_t* q_bytes, const uint64_t* q_off, const uint8_t* r_bytes, const uint64_t* r_off,
                                          AlignResult* out, uint32_t* runs, size_t runs_cap, uint64_t* run_off, uint32_t* run_len,
                                          size_t* runs_used, BaStats* stats) {
  if (!cfg || !(cfg->flags & BA_TRACE)) return fail(BA_ERR_ARG, "ba_align_batch_multi_cigar needs BA_TRACE");
  if (n && (!runs || !run_off || !run_len || !r_off)) return fail(BA_ERR_ARG, "null CIGAR output");
  if (!n) { if (stats) memset(stats, 0, sizeof(*stats)); if (runs_used) *runs_used = 0; return BA_OK; }
  std::vector<MultiShard> sh;
  std::vector<size_t> base, used;
  const int rc = multi_run(devices, n_dev, n, q_off, r_off, stats, sh, [&](BaAligner* a, MultiShard& s, int d) {
    if (base.empty()) return fail(BA_ERR_ARG, "internal: shard table not ready");
    CigarOut cg{runs + base[d], base[d + 1] - base[d], run_off + s.lo, run_len + s.lo, 0};
    const int r = align_batch_impl(a, cfg, s.hi - s.lo, q_bytes, q_off + s.lo, r_bytes, r_off + s.lo, out ? out + s.lo : nullptr, &s.st, &cg);
    used[d] = cg.used;
    if (!r) for (size_t k = s.lo; k < s.hi; k++) run_off[k] += base[d];
    return r;
  }, [&](std::vector<MultiShard>& shards) {
    // slice the caller's buffer: worst case per shard when it fits, proportional shares otherwise
    const size_t nd = shards.size();
    std::vector<uint64_t> worst(nd, 0);
    uint64_t wsum = 0;
    for (size_t d = 0; d < nd; d++) {
      const MultiShard& s = shards[d];
      worst[d] = (q_off[s.hi] - q_off[s.lo]) + (r_off[s.hi] - r_off[s.lo]) + 5 * (uint64_t)(s.hi - s.lo);
      wsum += worst[d];
    }
    base.assign(nd + 1, 0); used.assign(nd, 0);
    for (size_t d = 0; d < nd; d++) {
      const uint64_t share = wsum <= runs_cap ? worst[d] : (uint64_t)((long double)runs_cap * worst[d] / std::max<uint64_t>(wsum, 1));
      base[d + 1] = base[d] + (size_t)share;
    }
  });
  size_t tot = 0;
  for (size_t u : used) tot += u;
  if (runs_used) *runs_used = tot;
  return rc;
}

// sequence-to-profile batches (profiles built on every device from its shard of the raw PSSM rows)
extern "C" int ba_align_batch_multi_pssm(const int* devices, int n_dev, const BaConfig* cfg, size_t n,
                                         const uint8_t* q_bytes, const uint64_t* q_off, const BaPssmBatch* pssm,
                                         AlignResult* out, BaStats* stats) {
  if (!pssm || (n && !pssm->score_off)) return fail(BA_ERR_ARG, "pssm is null");
  if (!n) { if (stats) memset(stats, 0, sizeof(*stats)); return BA_OK; }
  std::vector<MultiShard> sh;
  // profile positions dominate the work of a pair: balance on |q| + 20 * len
  return multi_run(devices, n_dev, n, q_off, pssm->score_off, stats, sh, [&](BaAligner* a, MultiShard& s, int) {
    BaPssmBatch p = *pssm;
    p.score_off = pssm->score_off + s.lo;
    if (pssm->gap_off) p.gap_off = pssm->gap_off + s.lo;
    return ba_align_batch_pssm(a, cfg, s.hi - s.lo, q_bytes, q_off + s.lo, &p, out ? out + s.lo : nullptr, &s.st);
  }, [](std::vector<MultiShard>&) {});
}

// Walk the stored trace of pair k back from an arbitrary end position (Trace::cigar / cigar_eq,
// scan_block.rs:1469-1480). The trace of a pair stays resident until a later pair of the batch reuses its slot's
// arena: always for batches of one (the legacy block_cigar_* calls), otherwise for the last pair each slot ran --
// anything else is refused (BA_ERR_ARG) instead of walking somebody else's trace.
extern "C" int ba_batch_traceback(BaBatch* b, size_t k, size_t query_idx, size_t reference_idx, int eq,
                                  const uint32_t** runs, size_t* n_runs) {
  if (!b || !runs || !n_runs || !(b->cfg.flags & BA_TRACE) || k >= b->n || !b->downloaded) return fail(BA_ERR_ARG, "no trace available");
  if (eq && b->cfg.scoring == BA_SCORING_PROFILE) return fail(BA_ERR_ARG, "cigar_eq needs a reference sequence");
  // "Traceback cigar end position must be in bounds!" (scan_block.rs:1483)
  if (query_idx > b->h_qlen[k] || reference_idx > b->h_rlen[k]) return fail(BA_ERR_ARG, "Traceback cigar end position must be in bounds!");
  if (b->h_out[k].status != kOk) return fail(BA_ERR_OVERFLOW, "the alignment of this pair did not complete");
  BaAligner* al = b->al;
  AlLock lk(al->mu);
  dev_stream_t st = b->ss.stream;
#ifndef BA_EMU
  if (cudaSetDevice(al->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
#endif
  if (!b->resident[k]) return fail(BA_ERR_ARG, "the trace of this pair is no longer resident (a later launch of the batch reused its arena)");
  Params P = make_params(b);
  if (!b->d_tb_res && pool_alloc(al, (void**)&b->d_tb_res, sizeof(DevResult))) return BA_ERR_CUDA;
  if (dzero(b->d_cigar_used, 8, st)) return BA_ERR_CUDA;
  if (launch_traceback(P, (uint32_t)k, (uint32_t)query_idx, (uint32_t)reference_idx, eq, b->d_tb_res, st)) return BA_ERR_CUDA;
  DevResult res;
  if (d2h(&res, b->d_tb_res, sizeof(res), st) || dsync(st)) return BA_ERR_CUDA;
  if (res.status != kOk) return fail(BA_ERR_OVERFLOW, "traceback failed (buffer overflow or an end position the alignment never reached)");
  b->h_tb.resize(res.cigar_n);
  if (d2h(b->h_tb.data(), b->d_cigar + res.cigar_off, (size_t)res.cigar_n * 4, st) || dsync(st)) return BA_ERR_CUDA;
  *runs = b->h_tb.data(); *n_runs = res.cigar_n;
  return BA_OK;
}

extern "C" int ba_batch_total_stats(const BaBatch* b, BaStats* stats) {
  if (!b || !b->downloaded || !stats) return fail(BA_ERR_ARG, "no stats available");
  stats->cells = 0; stats->steps = 0; stats->n_failed = 0;
  for (size_t k = 0; k < b->n; k++) { stats->cells += b->h_out[k].cells; stats->steps += b->h_out[k].steps; if (b->h_out[k].status) stats->n_failed++; }
  return BA_OK;
}

// -------------------------------------------------------------------------------------------------
// Integer-ALU peak micro-benchmark: the roofline denominator for this kernel family (integer max-plus
// DP is bound by the issue rate of add/max on the INT pipes, not by HBM or tensor cores).
// Each thread runs 8 independent chains of the two DPX forms the aligner uses (max(a+b,c), max3).
// Reported: giga "i16-cell ops" per second, counting max(a+b,c) as 2 ops and max3 as 2 ops, which is
// how the 13-ops-per-cell figure of the recurrence is counted.
// -------------------------------------------------------------------------------------------------
#ifndef BA_EMU
template <bool PACKED>
__global__ void ba_int_peak_kernel(int* out, int iters, int seed) {
  unsigned a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = (unsigned)(seed + k * 7 + threadIdx.x);
  const unsigned g = PACKED ? 0xfffefffeu : (unsigned)(-(seed & 3) - 1);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      if (PACKED) {
        a[k] = __viaddmax_s16x2(a[k], g, a[(k + 1) & 7]);
        a[k] = __vimax3_s16x2(a[k], a[(k + 3) & 7], g);
      } else {
        a[k] = (unsigned)__viaddmax_s32((int)a[k], (int)g, (int)a[(k + 1) & 7]);
        a[k] = (unsigned)__vimax3_s32((int)a[k], (int)a[(k + 3) & 7], (int)g);
      }
    }
  }
  unsigned r = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) r ^= a[k];
  if (r == 0x7fffffffu) out[0] = (int)r;
}
#endif
static int measure_int_peak(BaAligner* al, bool packed, double* giga_ops_per_s) {
#ifdef BA_EMU
  (void)al; (void)packed; *giga_ops_per_s = 0; return BA_OK;
#else
  if (!al || !giga_ops_per_s) return fail(BA_ERR_ARG, "null");
  AlLock lk(al->mu);
  if (cudaSetDevice(al->device) != cudaSuccess) return fail(BA_ERR_CUDA, "cudaSetDevice failed");
  int* d = nullptr;
  if (dmalloc((void**)&d, 4)) return BA_ERR_CUDA;
  const int blocks = al->sm_count * 8, threads = 256, iters = 4096;
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(al->ev0, al->stream);
    if (packed) ba_int_peak_kernel<true><<<blocks, threads, 0, al->stream>>>(d, iters, 12345 + rep);
    else ba_int_peak_kernel<false><<<blocks, threads, 0, al->stream>>>(d, iters, 12345 + rep);
    cudaEventRecord(al->ev1, al->stream);
    if (cudaEventSynchronize(al->ev1) != cudaSuccess) { dfree(d); return fail(BA_ERR_CUDA, "int peak kernel failed"); }
    float ms = 0; cudaEventElapsedTime(&ms, al->ev0, al->ev1);
    if (rep > 0 && ms < best) best = ms;
  }
  dfree(d);
  // per chain step: 2 instructions x 2 ops (add + max / two max) x 1 (s32) or 2 (s16x2) values each
  const double ops = (double)blocks * threads * (double)iters * 8 * 4 * (packed ? 2 : 1);
  *giga_ops_per_s = ops / (best * 1e-3) / 1e9;
  return BA_OK;
#endif
}
// issue rate of the DPX add-max / max3 instructions the DP is made of: one 32-bit value per lane and instruction ...
extern "C" int ba_measure_int_peak(BaAligner* al, double* giga_ops_per_s) { return measure_int_peak(al, false, giga_ops_per_s); }
// ... and two 16-bit values per lane and instruction (VIADDMNMX.S16x2 / VIMNMX3.S16x2), the packed path's peak
extern "C" int ba_measure_int_peak_packed(BaAligner* al, double* giga_ops_per_s) { return measure_int_peak(al, true, giga_ops_per_s); }
