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
}}  // namespace ba::wp
#endif
