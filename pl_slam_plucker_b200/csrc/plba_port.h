// Kernel-source portability layer.
//
// Product build (nvcc, sm_100a): the macros below expand to plain CUDA (threadIdx, __syncthreads, atomics).
//
// -DPLBA_HOST_EMU (g++ only, used by tests/emu to debug KERNEL LOGIC on a box without a GPU): every kernel is a
// sequence of PHASE blocks separated by block barriers; the host build turns a phase into a loop over the CTA's
// threads and a launch into a loop over CTAs.  The emulation library is test tooling: it is never loaded by the
// package (pl_slam_plucker_b200/_lib.py loads libplba.so only and fails loudly when CUDA is unavailable).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstring>
#include <cmath>

#ifndef PLBA_HOST_EMU
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define PLBA_HD __host__ __device__ __forceinline__
#define PLBA_D __device__ __forceinline__
#define PLBA_KERNEL __global__
#define PLBA_SMEM(ptr) extern __shared__ __align__(16) unsigned char plba_smem_raw[]; unsigned char *ptr = plba_smem_raw
#define PHASE_BEGIN { const int tid = (int)threadIdx.x; (void)tid;
#define PHASE_END } __syncthreads();
#define PLBA_NT ((int)blockDim.x)
#define PLBA_BID ((int)blockIdx.x)
#define PLBA_BIDY ((int)blockIdx.y)
#define PLBA_NB ((int)gridDim.x)
#define PLBA_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define plba_isfinite(x) isfinite(x)
PLBA_D void plba_atomic_add(double *p, double v) { atomicAdd(p, v); }
PLBA_D void plba_atomic_add_i(int *p, int v) { atomicAdd(p, v); }
// max over non-negative doubles (bit pattern order == value order)
PLBA_D void plba_atomic_max_pos(double *p, double v) { atomicMax((unsigned long long *)p, (unsigned long long)__double_as_longlong(v)); }
// CTA-wide accumulation into one shared-memory word: warp shuffle tree, then one shared atomic per warp.
// Must be reached by every thread of the CTA (blockDim is a multiple of 32).
PLBA_D void plba_block_add(double *smem_dst, double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(smem_dst, v);
}
#else
// ------------------------------------------------------------------ host emulation
#include <cstdlib>
#include <vector>
#define PLBA_HD inline
#define PLBA_D inline
#define PLBA_KERNEL static
struct plba_dim3 { unsigned x, y, z; plba_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef plba_dim3 dim3;
struct PlbaEmuCtx { int nt = 1, bid = 0, bidy = 0, nb = 1; unsigned char *smem = nullptr; };
inline PlbaEmuCtx &plba_emu() { static thread_local PlbaEmuCtx c; return c; }
#define PLBA_SMEM(ptr) unsigned char *ptr = plba_emu().smem
#define PHASE_BEGIN for (int tid = 0; tid < plba_emu().nt; ++tid) {
#define PHASE_END }
#define PLBA_NT (plba_emu().nt)
#define PLBA_BID (plba_emu().bid)
#define PLBA_BIDY (plba_emu().bidy)
#define PLBA_NB (plba_emu().nb)
#define PLBA_LAUNCH(kernel, grid, block, smem_bytes, stream, ...)                                   \
    do {                                                                                            \
        dim3 g_ = dim3(grid), b_ = dim3(block);                                                                 \
        std::vector<unsigned char> sm_((size_t)(smem_bytes) + 64);                                  \
        PlbaEmuCtx &c_ = plba_emu();                                                                \
        c_.nt = (int)b_.x; c_.nb = (int)g_.x; c_.smem = sm_.data();                                 \
        for (unsigned by_ = 0; by_ < g_.y; by_++) for (unsigned bx_ = 0; bx_ < g_.x; bx_++) {       \
            c_.bid = (int)bx_; c_.bidy = (int)by_; kernel(__VA_ARGS__);                             \
        }                                                                                           \
    } while (0)
#define plba_isfinite(x) std::isfinite(x)
inline void plba_atomic_add(double *p, double v) { *p += v; }
inline void plba_atomic_add_i(int *p, int v) { *p += v; }
inline void plba_atomic_max_pos(double *p, double v) { if (v > *p) *p = v; }
inline void plba_block_add(double *d, double v) { *d += v; }
// minimal CUDA runtime stand-ins
typedef int cudaError_t; typedef void *cudaStream_t; typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return 0; }
inline cudaError_t cudaFree(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = std::calloc(n ? n : 1, 1); return 0; }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { if (n) std::memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char *cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return 0; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
#endif
