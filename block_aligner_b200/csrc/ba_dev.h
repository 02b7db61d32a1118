// ba_dev.h -- device abstraction shared by the host runtime (ba_runtime.cu) and the kernel translation units
// (ba_launch.cu): CUDA in the product build, plain host memory + the fiber emulator in the BA_EMU test build.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <string>

namespace ba {
std::string& last_error_ref();     // thread-local message behind ba_last_error_message()
inline int fail(int code, const std::string& msg) { last_error_ref() = msg; return code; }
}
using ba::fail;

#ifdef BA_EMU
typedef int dev_stream_t;
static inline int dmalloc(void** p, size_t n) { *p = malloc(n ? n : 1); if (!*p) return 1; memset(*p, 0xCD, n); return 0; }
static inline void dfree(void* p) { free(p); }
static inline int h2d(void* d, const void* s, size_t n, dev_stream_t) { if (n) memcpy(d, s, n); return 0; }
static inline int d2h(void* d, const void* s, size_t n, dev_stream_t) { if (n) memcpy(d, s, n); return 0; }
static inline int d2d(void* d, const void* s, size_t n, dev_stream_t) { if (n) memcpy(d, s, n); return 0; }
static inline int dzero(void* d, size_t n, dev_stream_t) { if (n) memset(d, 0, n); return 0; }
static inline int dsync(dev_stream_t) { return 0; }
#define CUDA_OK(x) (x)
#else
#include <cuda_runtime.h>
typedef cudaStream_t dev_stream_t;
static inline int cuda_fail(cudaError_t e, const char* what) {
  ba::last_error_ref() = std::string(what) + ": " + cudaGetErrorString(e);
  return 1;
}
#define CK(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return cuda_fail(_e, #x); } while (0)
static inline int dmalloc(void** p, size_t n) { CK(cudaMalloc(p, n ? n : 1)); return 0; }
static inline void dfree(void* p) { if (p) cudaFree(p); }
static inline int h2d(void* d, const void* s, size_t n, dev_stream_t st) { if (n) CK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st)); return 0; }
static inline int d2h(void* d, const void* s, size_t n, dev_stream_t st) { if (n) CK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st)); return 0; }
static inline int d2d(void* d, const void* s, size_t n, dev_stream_t st) { if (n) CK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st)); return 0; }
static inline int dzero(void* d, size_t n, dev_stream_t st) { if (n) CK(cudaMemsetAsync(d, 0, n, st)); return 0; }
static inline int dsync(dev_stream_t st) { CK(cudaStreamSynchronize(st)); return 0; }
#define CUDA_OK(x) (x)
#endif

// alignment-kernel dispatch (ba_launch.cu; one translation unit per scoring kind x fast-phase group in the product build)
namespace ba { struct Params; }
int ba_launch_dispatch(int scoring, int flags, int fr, const ba::Params& P, int blocks, int wpb, size_t smem, dev_stream_t st);
int ba_occupancy_dispatch(int scoring, int flags, int fr, int wpb, size_t smem, int* blocks_per_sm);
