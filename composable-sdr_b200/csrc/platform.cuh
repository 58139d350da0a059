// platform.cuh -- the one place that knows whether kernel sources are being compiled by nvcc (product,
// sm_100a) or by g++ inside the TEST-ONLY thread emulator (tests/emu/cuda_emu.h, -DCSDR_EMU).  The emulator
// exists so that tile/index arithmetic of the real kernel sources can be checked in the CPU-only test suite;
// it is never built into libcsdr_b200.so and the package never loads it.
#pragma once

#ifdef CSDR_EMU
#include "cuda_emu.h"
#define CSDR_DYN_SMEM(name) unsigned char *name = ::csdr_emu::dyn_smem()
#define CSDR_DYN_SMEM_1K(name) unsigned char *name = ::csdr_emu::dyn_smem()
#define CSDR_GRID_CONSTANT
#else
#include <cuda_runtime.h>
#define CSDR_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define CSDR_DYN_SMEM_1K(name) extern __shared__ __align__(1024) unsigned char name[]    /* 128-byte-swizzled TMA tiles */
#define CSDR_GRID_CONSTANT __grid_constant__
#endif

#include <stdint.h>

namespace csdr {

// ---- TMA bulk copy (cp.async.bulk, global -> shared, completion on an mbarrier) ------------------------------
// Used to stage the next tile of raw input while the current tile is being filtered.  Under the test-only CPU
// emulation the copy is a synchronous memcpy by the issuing thread and the barrier wait is a no-op (the kernels'
// own __syncthreads order the accesses).
#ifdef CSDR_EMU
__device__ inline void bulk_init(unsigned long long *) {}
__device__ inline void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *) { memcpy(dst, src, bytes); }
__device__ inline void bulk_wait(unsigned long long *, unsigned) {}
#else
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_init(unsigned long long *bar)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// one elected thread: arm the barrier with the byte count, then launch the copy (16-byte aligned, size % 16 == 0)
__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses to dst are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif

// ---- TMA tensor copy with the 128-byte swizzle: the raw input viewed as rows of 16 samples (128 bytes) ------------
// A tile of whole rows lands densely in shared memory (1024-byte aligned) with the 16-byte chunk c of row r stored at
// chunk c ^ (r & 7), so that threads which read the same chunk position of consecutive rows hit different banks.
// The descriptor is a CUtensorMap built on the host (cuTensorMapEncodeTiled) and passed as a __grid_constant__
// parameter.  Under the test-only CPU emulation it is a plain (base, rows, stream stride) triple and the copy is a
// loop that applies the same permutation.
#ifdef CSDR_EMU
struct FeTmap { const float2 *base; long long nrows; long long stream_stride; };
__device__ inline void tma_load_rows(void *dst, const FeTmap *tm, int row0, int stream, int rows, unsigned long long *, int = -1)
{
    const float2 *src = tm->base + (long long)stream * tm->stream_stride + (long long)row0 * 16;
    float2 *d = reinterpret_cast<float2 *>(dst);
    // like the hardware, the swizzle is a function of the shared-memory ADDRESS (bits 7-9 of the row's address), so a
    // second box that starts at a row which is not a multiple of 8 continues the pattern of the first
    const int phase = (int)((reinterpret_cast<uintptr_t>(dst) >> 7) & 7);
    for (int r = 0; r < rows; r++)
        for (int c = 0; c < 8; c++)
            for (int k = 0; k < 2; k++) d[(r * 8 + (c ^ ((r + phase) & 7))) * 2 + k] = src[r * 16 + c * 2 + k];
}
#else
struct alignas(64) FeTmap { unsigned long long opaque[16]; };
// announce: rows to announce on the barrier with this copy (-1: this box; 0: announced already by an earlier box)
__device__ __forceinline__ void tma_load_rows(void *dst, const FeTmap *tm, int row0, int stream, int rows, unsigned long long *bar,
                                              int announce = -1)
{
    (void)rows;
    if (announce != 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                     "r"((unsigned)(announce < 0 ? rows : announce) * 128u) : "memory");
    }
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(tm)), "r"(smem_u32(bar)), "r"(0), "r"(row0), "r"(stream)
                 : "memory");
}
#endif

// ---- asynchronous global -> shared copies (cp.async, SASS LDGSTS): issued early, waited for just before use --------
#ifdef CSDR_EMU
__device__ inline void cp_async16(void *dst, const void *src) { memcpy(dst, src, 16); }
__device__ inline void cp_async_commit() {}
__device__ inline void cp_async_wait_all() {}
template <int N> __device__ inline void cp_async_wait_group() {}
#else
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }   // all but the N newest groups
#endif

// ---- 16-byte values that validate themselves (decoupled look-back between CTAs): one vector store publishes, one
// vector load polls -- no flag word, hence no memory fence (a fence behind a tile's worth of output stores waits for all
// of them to drain).  Unpublished slots hold all-ones, which no arithmetic result is (the hardware's NaN is 0x7ff8...).
struct alignas(16) SelfValid16 { unsigned long long a, b; };
#ifdef CSDR_EMU
__device__ inline void sv16_store(SelfValid16 *p, double x, double y)
{
    unsigned long long a, b; memcpy(&a, &x, 8); memcpy(&b, &y, 8);
    __atomic_store_n(&p->b, b, __ATOMIC_SEQ_CST); __atomic_store_n(&p->a, a, __ATOMIC_SEQ_CST);
}
__device__ inline bool sv16_load(const SelfValid16 *p, double &x, double &y)
{
    const unsigned long long a = __atomic_load_n(&p->a, __ATOMIC_SEQ_CST), b = __atomic_load_n(&p->b, __ATOMIC_SEQ_CST);
    if (a == ~0ull || b == ~0ull) return false;
    memcpy(&x, &a, 8); memcpy(&y, &b, 8);
    return true;
}
#else
__device__ __forceinline__ void sv16_store(SelfValid16 *p, double x, double y)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(__double_as_longlong(x)), "l"(__double_as_longlong(y)) : "memory");
}
__device__ __forceinline__ bool sv16_load(const SelfValid16 *p, double &x, double &y)
{
    unsigned long long a, b;
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    if (a == ~0ull || b == ~0ull) return false;
    x = __longlong_as_double((long long)a); y = __longlong_as_double((long long)b);
    return true;
}
#endif

// ---- named barriers (bar.sync / bar.arrive id, count): producer / consumer hand-over between two groups of warps ---
// sync: wait until `count` threads have arrived (this one included); arrive: count this thread and go on.
#ifdef CSDR_EMU
__device__ inline void named_bar_sync(int id, int count) { ::csdr_emu::named_barrier(id, count, true); }
__device__ inline void named_bar_arrive(int id, int count) { ::csdr_emu::named_barrier(id, count, false); }
#else
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
#endif

constexpr int kMaxStages = 12;   // half-band stages (2^12 decimation) supported by the fused front end
constexpr int kMaxHbM    = 16;   // max half-band semi-length m (2m taps)
constexpr int kHsub      = 14;   // taps per polyphase branch of the arbitrary resampler (2*7, msresamp.c)
constexpr int kFePlanePad = 1;   // elements between the even and the odd plane of a level buffer: a thread pair that writes
                                 // (even, odd) samples of the same pair index then hits two different banks
constexpr int kFeTopR  = 8;           // outputs per thread slot of the first stage of k_frontend_direct = pairs per 128-byte row of the raw tile
constexpr int kFeRawRow = 16;    // samples per row of the swizzled raw tile
constexpr int kHcPad     = 16;   // c-rate history carried into every tile (>= kHsub-1, multiple of 8)

__host__ __device__ inline float2 cf(float re, float im) { float2 z; z.x = re; z.y = im; return z; }

// acc += g * v on both halves of a complex sample with ONE instruction: Blackwell's packed FP32 FMA (PTX
// fma.rn.f32x2, SASS FFMA2; a uniform real tap is broadcast as a scalar operand).  Two IEEE fused multiply-adds,
// bit-identical to two fmaf().
__device__ __forceinline__ void ffma2(float2 &acc, float g, float2 v)
{
#ifdef CSDR_EMU
    acc.x = fmaf(g, v.x, acc.x); acc.y = fmaf(g, v.y, acc.y);
#else
    unsigned long long a = *reinterpret_cast<unsigned long long *>(&acc);
    const unsigned long long b = *reinterpret_cast<const unsigned long long *>(&v);
    asm("{\n\t.reg .b64 gg;\n\tmov.b64 gg, {%2, %2};\n\tfma.rn.f32x2 %0, gg, %1, %0;\n\t}" : "+l"(a) : "l"(b), "f"(g));
    acc = *reinterpret_cast<float2 *>(&a);
#endif
}

}  // namespace csdr
