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

struct alignas(16) plba_d2 { double x, y; };
#if defined(__CUDA_ARCH__)
#define plba_rsqrt_hd(x) rsqrt(x)
#else
#define plba_rsqrt_hd(x) (1.0 / std::sqrt(x))
#endif   // 128-bit load unit of the SoA observation arrays

#ifndef PLBA_HOST_EMU
// ------------------------------------------------------------------ CUDA
#include <cuda_runtime.h>
#define PLBA_HD __host__ __device__ __forceinline__
#define PLBA_D __device__ __forceinline__
#define PLBA_KERNEL __global__
#define PLBA_COLD __host__ __device__ __noinline__
#define PLBA_BOUNDS(t, b) __launch_bounds__(t, b)
#define PLBA_SMEM(ptr) extern __shared__ __align__(16) unsigned char plba_smem_raw[]; unsigned char *ptr = plba_smem_raw
#define PLBA_SHARED __shared__
#define PHASE_BEGIN { const int tid = (int)threadIdx.x; (void)tid;
#define PHASE_END } __syncthreads();
#define PLBA_NT ((int)blockDim.x)
#define PLBA_BID ((int)blockIdx.x)
#define PLBA_BIDY ((int)blockIdx.y)
#define PLBA_NB ((int)gridDim.x)
#define PLBA_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define plba_isfinite(x) isfinite(x)
// accumulation into HBM/L2-resident sums: red.global.add.f64 (fire and forget).  Written as PTX so that the compiler can
// not fall back to a generic-address atomic (QSPC + branch + CAS loop for the shared window) when it loses the address space.
PLBA_D void plba_atomic_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
PLBA_D void plba_atomic_add_i(int *p, int v) { atomicAdd(p, v); }
// max over non-negative doubles (bit pattern order == value order)
PLBA_D void plba_atomic_max_pos(double *p, double v) { atomicMax((unsigned long long *)p, (unsigned long long)__double_as_longlong(v)); }
// CTA-wide accumulation into one shared-memory word: warp shuffle tree, then one shared atomic per warp.
// Must be reached by every thread of the CTA (blockDim is a multiple of 32).
PLBA_D void plba_block_add(double *smem_dst, double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(smem_dst, v);     // <= 8 adds per call: the shared-memory CAS loop is fine here
}
// ---- warp-autonomous kernels (plba_warp.h): a warp owns its work item and its slice of shared memory; phases are separated
// by __syncwarp only, so the warps of an SM drift apart and hide each other's latencies.  Lane-private values that live
// across phases are plain registers here and per-lane arrays in the host emulation (LANE_VAR / LANE_ARR / LANE_BIND).
#define WPHASE_BEGIN { const int lane = (int)(threadIdx.x & 31u); (void)lane;
#define WPHASE_END } __syncwarp();
#define PLBA_WARP_ID ((int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)))
#define PLBA_NWARPS ((int)(gridDim.x * (blockDim.x >> 5)))
#define PLBA_WARP_IN_CTA ((int)(threadIdx.x >> 5))
#define LANE_VAR(T, name) T name
#define LANE_ARR(T, name, N) T name[N]
#define LANE_BIND(name)
// thread-private values that live across the PHASE blocks of a CTA kernel (registers here, per-thread arrays in the host emulation)
#define THR_VAR(T, name) T name
#define THR_ARR(T, name, N) T name[N]
#define THR_BIND(name)
PLBA_D double plba_warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }
PLBA_D double plba_warp_max(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); return v; }
// every lane of the warp calls these (inside a WPHASE): the warp's lanes are summed and leave the SM as one red / atomicMax
#define PLBA_WARP_FLUSH_ADD(ptr, v) do { const double s_ = plba_warp_sum(v); if (lane == 0 && s_ != 0.0) plba_atomic_add((ptr), s_); } while (0)
#define PLBA_WARP_FLUSH_MAX(ptr, v) do { const double s_ = plba_warp_max(v); if (lane == 0 && s_ > 0.0) plba_atomic_max_pos((ptr), s_); } while (0)
PLBA_HD void plba_sincos(double x, double *s, double *c) { sincos(x, s, c); }
// pull a line into L1 ahead of its use (no register is tied up: the next pass of a warp kernel reads it after ~1 000 instructions)
PLBA_D void plba_prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// optional per-phase cycle accounting (development builds only: -DPLBA_PROF): thread 0 of every CTA adds the cycles between
// two marks to a global table that tools read back through plba_debug_prof()
#ifdef PLBA_PROF
__device__ unsigned long long plba_prof_table[64];
#define PROF_DECL long long prof_t0_ = clock64()
#define PROF_MARK(id) do { if (threadIdx.x == 0) { const long long t_ = clock64(); atomicAdd(&plba_prof_table[id], (unsigned long long)(t_ - prof_t0_)); prof_t0_ = t_; } } while (0)
#else
#define PROF_DECL
#define PROF_MARK(id)
#endif
// kernel parameters live in device memory (one fixed address per handle, so that the CUDA graph of the LM loop never
// has to be rebuilt); every CTA stages them into shared memory once.
#define PLBA_COUNT_LAUNCH(Pp) do { if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) atomicAdd(&(Pp)->counters[11 /* CNT_KLAUNCH */], 1); } while (0)
#define PLBA_PARAMS(P, Pp)                                                                                   \
    PLBA_COUNT_LAUNCH(Pp);                                                                                   \
    for (int i_ = (int)threadIdx.x; i_ < (int)(sizeof(DevP) / 4); i_ += (int)blockDim.x)                       \
        ((int *)&plba_params_smem)[i_] = ((const int *)(Pp))[i_];                                            \
    __syncthreads();                                                                                         \
    const DevP &P = plba_params_smem
// inside device functions: bind to the shared copy directly so that parameter reads stay shared-space loads
#define PLBA_PARAMS_REF(P, Pin) const DevP &P = plba_params_smem; (void)Pin
// reads of values that other CTAs of the SAME launch produced with atomics (L2 is the point of coherence)
PLBA_D double plba_ld_l2(const double *p) { return __ldcg(p); }
PLBA_D int plba_ld_l2(const int *p) { return __ldcg(p); }
PLBA_D void plba_fence() { __threadfence(); }
#define PLBA_TID0 (threadIdx.x)
// Block-wide wait for a flag that another kernel running at the same time sets to non-zero (release / acquire at GPU scope).  Thread 0
// polls, the block barrier hands the acquire on to the other threads.
PLBA_D void plba_wait_flag(const int *flag, int at_least = 1) {
    if (threadIdx.x == 0) {
        int v;
        do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory"); if (v < at_least) __nanosleep(100); } while (v < at_least);
    }
    __syncthreads();
}
PLBA_D void plba_set_flag(int *flag, int value = 1) { __threadfence(); asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory"); }
// programmatic dependent launch: kernels connected to this one by a programmatic edge may start now
PLBA_D void plba_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
PLBA_D int plba_atomic_fetch_add_i(int *p, int v) { return atomicAdd(p, v); }
PLBA_D double plba_rsqrt(double x) { return rsqrt(x); }   // (an FP32-seeded Newton variant measured slower: profiles/README.md)
// Reciprocal / reciprocal square root for LATENCY-bound dependency chains (the 6x6 pivots of k_solve_small): hardware seed (2^-23) plus
// two Newton steps, no special-case branches: 5 / 7 dependent operations instead of the ~10 of the IEEE-rounded library routines.
// Faithful to about one ulp for normal, positive arguments (callers have already rejected non-positive or non-finite pivots).
PLBA_HD double plba_rcp_fast(double d) {
#if defined(__CUDA_ARCH__)
    double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0); r = fma(r, e, r);
    e = fma(-d, r, 1.0); r = fma(r, e, r);
    return r;
#else
    return 1.0 / d;
#endif
}
// the same for arguments that are not known to be normal and non-zero (depths, line norms): the IEEE routine takes over outside
// [1e-290, 1e290] (never in practice; keeps zero / infinite / NaN inputs on the library's behaviour)
PLBA_HD double plba_rcp_safe(double d) { const double a = fabs(d); return (a > 1e-290 && a < 1e290) ? plba_rcp_fast(d) : 1.0 / d; }
PLBA_HD double plba_rsqrt_fast(double x);
PLBA_HD double plba_rsqrt_safe(double x) { return (x > 1e-290 && x < 1e290) ? plba_rsqrt_fast(x) : 1.0 / sqrt(x); }
PLBA_HD double plba_rsqrt_fast(double x) {
#if defined(__CUDA_ARCH__)
    double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = y * y, e = fma(-x, t, 1.0); y = fma(0.5 * y, e, y);
    t = y * y; e = fma(-x, t, 1.0); y = fma(0.5 * y, e, y);
    return y;
#else
    return 1.0 / std::sqrt(x);
#endif
}
// WHILE / IF nodes of the LM-loop graph are steered from the controller (cudaGraphSetConditional is a device runtime builtin)
PLBA_D void plba_graph_set(unsigned long long handle, unsigned int v) { cudaGraphSetConditional((cudaGraphConditionalHandle)handle, v); }
#else
// ------------------------------------------------------------------ host emulation
#include <cstdlib>
#include <vector>
#define PLBA_HD inline
#define PLBA_D inline
#define PLBA_KERNEL static
#define PLBA_COLD inline
#define PLBA_BOUNDS(t, b)
struct plba_dim3 { unsigned x, y, z; plba_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
typedef plba_dim3 dim3;
struct PlbaEmuCtx { int nt = 1, bid = 0, bidy = 0, nb = 1; unsigned char *smem = nullptr; };
inline PlbaEmuCtx &plba_emu() { static thread_local PlbaEmuCtx c; return c; }
#define PLBA_SMEM(ptr) unsigned char *ptr = plba_emu().smem
#define PLBA_SHARED static
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
#define WPHASE_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define WPHASE_END }
#define PLBA_WARP_ID (plba_emu().bid)          // the emulation launches warp kernels with one warp per "CTA"
#define PLBA_NWARPS (plba_emu().nb)
#define PLBA_WARP_IN_CTA 0
#define LANE_VAR(T, name) T name##_L[32]
#define LANE_ARR(T, name, N) T name##_L[32][N]
#define LANE_BIND(name) auto &name = name##_L[lane]
#define THR_VAR(T, name) T name##_T[1024]
#define THR_ARR(T, name, N) T name##_T[1024][N]
#define THR_BIND(name) auto &name = name##_T[tid]
#define PLBA_WARP_FLUSH_ADD(ptr, v) do { *(ptr) += (v); } while (0)
#define PLBA_WARP_FLUSH_MAX(ptr, v) do { if ((v) > *(ptr)) *(ptr) = (v); } while (0)
inline void plba_sincos(double x, double *s, double *c) { *s = std::sin(x); *c = std::cos(x); }
inline void plba_prefetch_l1(const void *) {}
#define PLBA_COUNT_LAUNCH(Pp) do { if (plba_emu().bid == 0 && plba_emu().bidy == 0) (Pp)->counters[11]++; } while (0)
#define PLBA_PARAMS(P, Pp) PLBA_COUNT_LAUNCH(Pp); const DevP &P = *(Pp)
#define PLBA_PARAMS_REF(P, Pin) const DevP &P = Pin
#define PROF_DECL
#define PROF_MARK(id)
inline double plba_ld_l2(const double *p) { return *p; }
inline int plba_ld_l2(const int *p) { return *p; }
inline void plba_fence() {}
#define PLBA_TID0 0
inline void plba_wait_flag(const int *flag, int at_least = 1) { if (*flag < at_least) std::abort(); }      // the emulation runs kernels one after the other: the flag must be set already
inline void plba_set_flag(int *flag, int value = 1) { *flag = value; }
inline void plba_launch_dependents() {}
inline int plba_atomic_fetch_add_i(int *p, int v) { int o = *p; *p += v; return o; }
inline double plba_rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline double plba_rcp_fast(double d) { return 1.0 / d; }
inline double plba_rsqrt_fast(double x) { return 1.0 / std::sqrt(x); }
inline double plba_rcp_safe(double d) { return 1.0 / d; }
inline double plba_rsqrt_safe(double x) { return 1.0 / std::sqrt(x); }
inline void plba_graph_set(unsigned long long, unsigned int) {}
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
