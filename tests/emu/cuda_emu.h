// cuda_emu.h -- TEST-ONLY CUDA thread emulator (CPU).
//
// Lets the *unmodified* kernel sources under composable-sdr_b200/csrc/*.cuh be compiled with g++ and run one
// thread block at a time, one OS thread per CUDA thread, with a real barrier behind __syncthreads().  Purpose:
// check tile geometry / index arithmetic / state carry of the kernels against the oracle in the CPU-only test
// suite (there is no GPU in the build container).  It is NOT a product path: nothing under composable-sdr_b200/
// includes or loads it, libcsdr_b200.so is built by nvcc only and fails loudly without a CUDA device.
#pragma once
#include <pthread.h>
#include <mutex>
#include <condition_variable>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <cmath>
#include <algorithm>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
static inline double2 make_double2(double a, double b) { double2 r; r.x = a; r.y = b; return r; }

namespace csdr_emu {
inline thread_local uint3 t_threadIdx, t_blockIdx;
inline uint3 g_blockDim, g_gridDim;
inline pthread_barrier_t g_barrier;
// warp-level exchange (shuffles, ballots): every lane of a warp must take part, as on the GPU with a full mask
constexpr int kMaxWarps = 64;
inline pthread_barrier_t g_warp_barrier[kMaxWarps];
inline unsigned long long g_warp_slot[kMaxWarps][32];
inline thread_local unsigned t_warp = 0, t_lane = 0, t_warp_lanes = 32;
inline unsigned char *g_dyn_smem = nullptr;
inline unsigned char *dyn_smem() { return g_dyn_smem; }
inline unsigned long long g_launches = 0;

// named barriers (bar.sync / bar.arrive id, count)
struct NamedBarrier { std::mutex m; std::condition_variable cv; int count = 0; unsigned gen = 0; };
inline NamedBarrier g_named[16];
inline void named_barrier(int id, int expected, bool wait)
{
    NamedBarrier &b = g_named[id & 15];
    std::unique_lock<std::mutex> lk(b.m);
    const unsigned g = b.gen;
    if (++b.count == expected) { b.count = 0; b.gen++; b.cv.notify_all(); return; }
    if (wait) b.cv.wait(lk, [&] { return b.gen != g; });
}

template <class K, class... Args>
void launch(dim3 grid, dim3 block, size_t smem_bytes, K kernel, Args... args)
{
    g_launches++;
    g_blockDim = {block.x, block.y, block.z};
    g_gridDim = {grid.x, grid.y, grid.z};
    unsigned nthreads = block.x * block.y * block.z;
    std::vector<unsigned char> smem(smem_bytes + 1024 + 64);          // 1024-byte aligned like the swizzled TMA tiles need
    g_dyn_smem = (unsigned char *)(((uintptr_t)smem.data() + 1023) & ~(uintptr_t)1023);
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        for (auto &nb : g_named) { nb.count = 0; }
        pthread_barrier_init(&g_barrier, nullptr, nthreads);
        const unsigned nwarps = (nthreads + 31) / 32;
        if (nwarps > (unsigned)kMaxWarps) abort();
        for (unsigned w = 0; w < nwarps; w++) pthread_barrier_init(&g_warp_barrier[w], nullptr, std::min(32u, nthreads - 32 * w));
        std::vector<std::thread> th;
        th.reserve(nthreads);
        for (unsigned t = 0; t < nthreads; t++) {
            th.emplace_back([=]() {
                t_blockIdx = {bx, by, bz};
                t_threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                t_warp = t / 32; t_lane = t % 32; t_warp_lanes = std::min(32u, nthreads - 32 * (t / 32));
                kernel(args...);
            });
        }
        for (auto &x : th) x.join();
        pthread_barrier_destroy(&g_barrier);
        for (unsigned w = 0; w < nwarps; w++) pthread_barrier_destroy(&g_warp_barrier[w]);
    }
    g_dyn_smem = nullptr;
}
}  // namespace csdr_emu

#define threadIdx (::csdr_emu::t_threadIdx)
#define blockIdx (::csdr_emu::t_blockIdx)
#define blockDim (::csdr_emu::g_blockDim)
#define gridDim (::csdr_emu::g_gridDim)

static inline void __syncthreads() { pthread_barrier_wait(&::csdr_emu::g_barrier); }
static inline void __syncwarp() { pthread_barrier_wait(&::csdr_emu::g_warp_barrier[::csdr_emu::t_warp]); }    // lanes are OS threads here: a real barrier
namespace csdr_emu {
template <class T> inline T warp_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "emulated shuffles move at most 8 bytes");
    unsigned long long raw = 0;
    memcpy(&raw, &v, sizeof(T));
    g_warp_slot[t_warp][t_lane] = raw;
    pthread_barrier_wait(&g_warp_barrier[t_warp]);
    T out = v;
    if (src_lane >= 0 && src_lane < (int)t_warp_lanes) memcpy(&out, &g_warp_slot[t_warp][src_lane], sizeof(T));
    pthread_barrier_wait(&g_warp_barrier[t_warp]);
    return out;
}
}  // namespace csdr_emu
template <class T> static inline T __shfl_sync(unsigned, T v, int src) { return ::csdr_emu::warp_exchange(v, src & 31); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d)
{
    const int src = (int)::csdr_emu::t_lane - (int)d;
    return ::csdr_emu::warp_exchange(v, src < 0 ? (int)::csdr_emu::t_lane : src);      // lanes below d keep their own value
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return ::csdr_emu::warp_exchange(v, (int)(::csdr_emu::t_lane ^ (unsigned)m)); }
static inline unsigned __ballot_sync(unsigned, int pred)
{
    using namespace ::csdr_emu;
    g_warp_slot[t_warp][t_lane] = pred ? 1ull : 0ull;
    pthread_barrier_wait(&g_warp_barrier[t_warp]);
    unsigned r = 0;
    for (unsigned l = 0; l < t_warp_lanes; l++) r |= (unsigned)g_warp_slot[t_warp][l] << l;
    pthread_barrier_wait(&g_warp_barrier[t_warp]);
    return r;
}
static inline void __threadfence() { __sync_synchronize(); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }

// math intrinsics (host libm stands in for the SFU approximations; tests use tolerances)
// (glibc declares __sinf & co. itself, so the CUDA intrinsic names are mapped by macro)
static inline float emu_sinf(float x) { return sinf(x); }
static inline float emu_cosf(float x) { return cosf(x); }
static inline void emu_sincosf(float x, float *s, float *c) { *s = sinf(x); *c = cosf(x); }
static inline float emu_log2f(float x) { return log2f(x); }
static inline float emu_expf(float x) { return expf(x); }
static inline float emu_logf(float x) { return logf(x); }
#define __sinf emu_sinf
#define __cosf emu_cosf
#define __sincosf emu_sincosf
#define __log2f emu_log2f
#define __expf emu_expf
#define __logf emu_logf
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __int2float_rn(int x) { return (float)x; }
static inline float __uint2float_rn(unsigned x) { return (float)x; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline long long min(long long a, long long b) { return a < b ? a : b; }
static inline long long max(long long a, long long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }

// atomics
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicMin(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
