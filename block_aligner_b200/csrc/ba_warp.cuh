// ba_warp.cuh -- the handful of warp-level and DPX primitives the aligner is written against.
//
// Product build (nvcc, sm_100a): thin wrappers over CUDA intrinsics.
// Test build (g++ -DBA_EMU, tests/emu/): the same names are backed by a cooperative-fiber SIMT
// emulator so the *same device source* can be checked against the CPU oracle without a GPU.
// The emulator is test infrastructure; the shipped library contains no CPU execution path.
#pragma once
#include <stdint.h>

#ifdef BA_EMU
#include "emu.h"
#define BA_DEV inline
#define BA_HD inline
#define BA_DEV_NOINLINE static
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
namespace ba { namespace wp {
inline int lane_id() { return emu::lane(); }
inline int shfl_up(int v, int d) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return l >= d ? (int)a[l - d] : v; }
inline int shfl_down(int v, int d) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return l + d < 32 ? (int)a[l + d] : v; }
inline int shfl_idx(int v, int src) { const uint32_t* a = emu::exchange((uint32_t)v); return (int)a[src & 31]; }
// width-8 sub-warp variants (CUDA semantics of the `width` argument of __shfl_*_sync)
inline int shfl_up8(int v, int d) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (l & 7) >= d ? (int)a[l - d] : v; }
inline int shfl_down8(int v, int d) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (l & 7) + d < 8 ? (int)a[l + d] : v; }
inline int shfl_idx8(int v, int src) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (int)a[(l & ~7) | (src & 7)]; }
inline int shfl_xor8(int v, int m) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (int)a[(l & ~7) | ((l ^ m) & 7)]; }
inline int shfl_up_w(int v, int d, int w) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (l & (w - 1)) >= d ? (int)a[l - d] : v; }
inline int shfl_idx_w(int v, int src, int w) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (int)a[(l & ~(w - 1)) | (src & (w - 1))]; }
inline int shfl_xor_w(int v, int m, int w) { const uint32_t* a = emu::exchange((uint32_t)v); int l = emu::lane(); return (int)a[(l & ~(w - 1)) | ((l ^ m) & (w - 1))]; }
// packed 2 x i16 (DPX / SIMD-in-a-word) operations: lane-wise on the two halfwords, wrapping adds
inline int h_lo(uint32_t p) { return (int)(int16_t)(p & 0xffffu); }
inline int h_hi(uint32_t p) { return (int)(int16_t)(p >> 16); }
inline uint32_t h_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
inline uint32_t vadd2(uint32_t a, uint32_t b) { return h_pack(h_lo(a) + h_lo(b), h_hi(a) + h_hi(b)); }
inline uint32_t vmax2(uint32_t a, uint32_t b) { return h_pack(h_lo(a) > h_lo(b) ? h_lo(a) : h_lo(b), h_hi(a) > h_hi(b) ? h_hi(a) : h_hi(b)); }
inline uint32_t vmin2(uint32_t a, uint32_t b) { return h_pack(h_lo(a) < h_lo(b) ? h_lo(a) : h_lo(b), h_hi(a) < h_hi(b) ? h_hi(a) : h_hi(b)); }
inline uint32_t vmax3_2(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
inline uint32_t vmin3_2(uint32_t a, uint32_t b, uint32_t c) { return vmin2(vmin2(a, b), c); }
inline uint32_t viaddmax2(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vadd2(a, b), c); }
inline uint32_t viaddmin2(uint32_t a, uint32_t b, uint32_t c) { return vmin2(vadd2(a, b), c); }
inline uint32_t vibmax2(uint32_t a, uint32_t b, bool& ph, bool& pl) { pl = h_lo(a) >= h_lo(b); ph = h_hi(a) >= h_hi(b); return vmax2(a, b); }
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t src = ((uint64_t)b << 32) | a; uint32_t r = 0;
  for (int i = 0; i < 4; i++) { const uint32_t n = (sel >> (4 * i)) & 0xfu; uint32_t byte = (uint32_t)(src >> (8 * (n & 7))) & 0xffu; if (n & 8) byte = (byte & 0x80u) ? 0xffu : 0u; r |= byte << (8 * i); }
  return r;
}
inline uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) { return prmt(a, b, sel); }
inline uint32_t opaque_zero() { return 0u; }
inline void pmov_fma(bool p, uint32_t& d, uint32_t s, uint32_t) { if (p) d = s; }
inline int red_max(int v) { const uint32_t* a = emu::exchange((uint32_t)v); int m = (int)a[0]; for (int i = 1; i < 32; i++) m = (int)a[i] > m ? (int)a[i] : m; return m; }
inline unsigned red_max_u(unsigned v) { const uint32_t* a = emu::exchange(v); unsigned m = a[0]; for (int i = 1; i < 32; i++) m = a[i] > m ? a[i] : m; return m; }
inline unsigned ballot(bool p) { const uint32_t* a = emu::exchange(p ? 1u : 0u); unsigned m = 0; for (int i = 0; i < 32; i++) m |= (a[i] & 1u) << i; return m; }
inline void syncwarp() { emu::exchange(0); }
inline unsigned atomic_add(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
inline unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline int viaddmax(int a, int b, int c) { int s = a + b; return s > c ? s : c; }
inline int vimax3(int a, int b, int c) { int m = a > b ? a : b; return m > c ? m : c; }
inline int imax(int a, int b) { return a > b ? a : b; }
inline int imin(int a, int b) { return a < b ? a : b; }
template <class T> inline T ldg(const T* p) { return *p; }
inline void touch(const void*) {}
inline void touch_l2(const void*) {}
inline void cp_async16(void* dst, const void* src) { __builtin_memcpy(dst, src, 16); }
inline void cp_async_commit() {}
inline void cp_async_wait_all() {}
}}  // namespace ba::wp
#else
#define BA_DEV __device__ __forceinline__
#define BA_HD __host__ __device__ __forceinline__
#define BA_DEV_NOINLINE __device__ __noinline__
namespace ba { namespace wp {
constexpr unsigned kFull = 0xffffffffu;
BA_DEV int lane_id() { return (int)(threadIdx.x & 31); }
BA_DEV int shfl_up(int v, int d) { return __shfl_up_sync(kFull, v, d); }
BA_DEV int shfl_down(int v, int d) { return __shfl_down_sync(kFull, v, d); }
BA_DEV int shfl_idx(int v, int src) { return __shfl_sync(kFull, v, src); }
BA_DEV int shfl_up8(int v, int d) { return __shfl_up_sync(kFull, v, d, 8); }
BA_DEV int shfl_down8(int v, int d) { return __shfl_down_sync(kFull, v, d, 8); }
BA_DEV int shfl_idx8(int v, int src) { return __shfl_sync(kFull, v, src, 8); }
BA_DEV int shfl_xor8(int v, int m) { return __shfl_xor_sync(kFull, v, m, 8); }
BA_DEV int shfl_up_w(int v, int d, int w) { return __shfl_up_sync(kFull, v, d, w); }
BA_DEV int shfl_idx_w(int v, int src, int w) { return __shfl_sync(kFull, v, src, w); }
BA_DEV int shfl_xor_w(int v, int m, int w) { return __shfl_xor_sync(kFull, v, m, w); }
// packed 2 x i16 DPX operations (VIADD.16x2, VIMNMX.S16x2, VIMNMX3.S16x2, VIADDMNMX.S16x2 on sm_100a)
BA_DEV int h_lo(uint32_t p) { return (int)(int16_t)(p & 0xffffu); }
BA_DEV int h_hi(uint32_t p) { return (int)p >> 16; }
BA_DEV uint32_t h_pack(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
BA_DEV uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
BA_DEV uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
BA_DEV uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
BA_DEV uint32_t vmax3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
BA_DEV uint32_t vmin3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
BA_DEV uint32_t viaddmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
BA_DEV uint32_t viaddmin2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmin_s16x2(a, b, c); }
BA_DEV uint32_t vibmax2(uint32_t a, uint32_t b, bool& ph, bool& pl) { return __vibmax_s16x2(a, b, &ph, &pl); }
BA_DEV uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
// PRMT with sign replication (selector nibbles with bit 3 set copy the sign of the selected byte into all eight bits):
// spelled in PTX, the documented form of that mode
BA_DEV uint32_t prmt_sx(uint32_t a, uint32_t b, uint32_t sel) { uint32_t d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel)); return d; }
// always 0 (blocks are one-dimensional) but thread-varying for the compiler: keeps a value out of the uniform datapath
BA_DEV uint32_t opaque_zero() { return threadIdx.y; }
// if (p) d = s, issued as a predicated IMAD (s * one + 0, `one` = 1 but opaque to the compiler): the FMA pipe instead of
// the ALU pipe, which PRMT / SEL / MOV would use and which bounds the packed column loop
BA_DEV void pmov_fma(bool p, uint32_t& d, uint32_t s, uint32_t one) {
  asm("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %1, 0;\n\t@q mad.lo.u32 %0, %2, %3, 0;\n\t}" : "+r"(d) : "r"((uint32_t)p), "r"(s), "r"(one));
}
BA_DEV int red_max(int v) { return __reduce_max_sync(kFull, v); }
BA_DEV unsigned red_max_u(unsigned v) { return __reduce_max_sync(kFull, v); }
BA_DEV unsigned ballot(bool p) { return __ballot_sync(kFull, p); }
BA_DEV void syncwarp() { __syncwarp(); }
BA_DEV unsigned atomic_add(unsigned* p, unsigned v) { return atomicAdd(p, v); }
BA_DEV unsigned long long atomic_add64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
// DPX (sm_90+): max(a + b, c) and 3-input max in one instruction each
BA_DEV int viaddmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }
BA_DEV int vimax3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
BA_DEV int imax(int a, int b) { return max(a, b); }
BA_DEV int imin(int a, int b) { return min(a, b); }
template <class T> BA_DEV T ldg(const T* p) { return __ldg(p); }
// bring the 128-byte line of p into L1 / L2 without waiting for it (a load whose result is never used)
// (a prefetch instruction, not a load into a dummy register: the register of an outstanding load is scoreboarded, and
// the next instruction the compiler gives that register to waits for the data -- measured in the batch traceback)
BA_DEV void touch(const void* p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
BA_DEV void touch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
// asynchronous 16-byte copy global -> shared (LDGSTS, bypassing L1): the issuing thread does not wait for the data
BA_DEV void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
BA_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
BA_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
}}  // namespace ba::wp
#endif
