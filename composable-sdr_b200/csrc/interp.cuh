// interp.cuh -- msresamp_crcf for rates above 1 (MSRESAMP(_interp_execute), liquid msresamp.c): the mirror image of
// the decimating front end -- the arbitrary resampler (rate_arb in (1, 2], 1-2 outputs per input) runs FIRST, then S
// half-band interpolators double the rate, lowest-rate stage first.  Reference: resampler r, Liquid.chs:56-117 (the
// Haskell side passes any r = bandwidth / samplerate).
//
// Every object starts from zeroed delay lines, so each stage is a causal FIR of the ABSOLUTE sample index and the
// whole cascade is time-parallel:
//   arbitrary stage   A[O] = sum_j bank[(O step mod 2^24) >> (24 - bits)][j] * x[(O step >> 24) - j]       (resamp.fixed.c)
//   half-band stage   V[2i] = U[i - m],   V[2i+1] = sum_t h1[t] * U[i - (2m-1) + t]                        (resamp2.c)
// A call recomputes the few samples of halo each stage needs from `hcap` raw input samples kept from the previous
// call, so no per-stage state is carried.  One kernel per stage, intermediates in global memory (they stay in L2);
// the path is bound by writing 2^S r output samples per input sample and is not the hot path of the chain.
#pragma once
#include "platform.cuh"
#include "frontend.cuh"
#include <algorithm>

namespace csdr {

constexpr int kInterpHcap = 64;     // raw input samples carried between calls (halo of the whole cascade <= ~45)

struct InterpArbParams {
    const float2 *x, *hist; long long x_stride; int hcap;
    long long n0, nx;               // absolute index of x[0], samples in the chunk
    unsigned step; int bits; const float *bank;     // [npfb][kHsub]
    long long O_begin; int count;   // absolute index of A[0] (>= 0), samples to compute
    float2 *A; long long A_stride;
};

__global__ void k_interp_arb(const InterpArbParams p)
{
    const float2 *xs = p.x + (long long)blockIdx.y * p.x_stride;
    const float2 *hs = p.hist + (long long)blockIdx.y * p.hcap;
    float2 *A = p.A + (long long)blockIdx.y * p.A_stride;
    const unsigned mask = (1u << p.bits) - 1u;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < p.count; idx += gridDim.x * blockDim.x) {
        const unsigned long long ph = (unsigned long long)(p.O_begin + idx) * p.step;
        const long long k = (long long)(ph >> 24);                       // absolute index of the newest input sample
        const float *h = p.bank + (((unsigned)ph >> (24 - p.bits)) & mask) * kHsub;
        float ar = 0.f, ai = 0.f;
#pragma unroll
        for (int j = 0; j < kHsub; j++) {
            const long long i = k - j, rel = i - p.n0;
            float2 v = cf(0.f, 0.f);
            if (i >= 0) v = (rel >= 0) ? xs[rel] : hs[p.hcap + rel];     // rel >= -hcap by construction
            ar = fmaf(h[j], v.x, ar);
            ai = fmaf(h[j], v.y, ai);
        }
        A[idx] = cf(ar, ai);
    }
}

struct InterpHbParams {
    const float2 *U; long long U_begin, U_stride;    // absolute index of U[0] (>= 0)
    float2 *V; long long V_begin, V_stride; long long count;   // absolute index of V[0] (>= 0), samples to compute
    int m; float h1[2 * kMaxHbM];
};

__global__ void k_interp_hb(const InterpHbParams p)
{
    const float2 *U = p.U + (long long)blockIdx.y * p.U_stride;
    float2 *V = p.V + (long long)blockIdx.y * p.V_stride;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < p.count; idx += stride) {
        const long long J = p.V_begin + idx, I = J >> 1;
        float2 out = cf(0.f, 0.f);
        if ((J & 1) == 0) {
            const long long i = I - p.m;                                  // delay branch
            if (i >= 0) out = U[i - p.U_begin];
        } else {
            float ar = 0.f, ai = 0.f;
            for (int t = 0; t < 2 * p.m; t++) {                           // oldest sample first, as liquid's dotprod
                const long long i = I - (2 * p.m - 1) + t;
                if (i >= 0) { const float2 v = U[i - p.U_begin]; ar = fmaf(p.h1[t], v.x, ar); ai = fmaf(p.h1[t], v.y, ai); }
            }
            out = cf(ar, ai);
        }
        V[idx] = out;
    }
}

// ---- launch sequence (shared by csdr_b200.cu and the CPU-only emulation test) --------------------------------
struct InterpPlan {
    int S = 0; int m[kMaxStages] = {}; float h1[kMaxStages][2 * kMaxHbM] = {};
    unsigned step = 0; int bits = 0;
};

inline unsigned long long interp_outputs_before(unsigned long long n_abs, unsigned step)
{
    // arbitrary-stage outputs emitted by the first n_abs pushes: O with O step < n_abs 2^24
    const unsigned __int128 span = (unsigned __int128)n_abs << 24;
    return (unsigned long long)((span + step - 1) / step);
}
inline long long interp_max_out(const InterpPlan &ip, long long nx)
{
    return (long long)(((unsigned __int128)(nx + 1) << 24) / ip.step + 2) << ip.S;
}

// x: chunk (already mixed if the chain mixes), hist: kInterpHcap samples before it.  buf(slot, bytes) returns scratch
// memory (slots 0, 1: ping-pong, at least nstreams * stride * 8 bytes).  Returns outputs per stream written to y.
template <class Launch, class Buf>
inline long long interp_launch(Launch &launch, Buf &&buf, const InterpPlan &ip, const float *bank_dev, int nstreams,
                               const float2 *x, long long x_stride, const float2 *hist, unsigned long long n_abs,
                               long long nx, float2 *y, long long y_stride)
{
    const long long O0 = (long long)interp_outputs_before(n_abs, ip.step);
    const long long O1 = (long long)interp_outputs_before(n_abs + (unsigned long long)nx, ip.step);
    if (O1 <= O0) return 0;
    // ranges [b[L], e[L]) of absolute indices needed at every level: L = S is the output, L = 0 the arbitrary stage
    long long b[kMaxStages + 1], e[kMaxStages + 1];
    b[ip.S] = O0 << ip.S; e[ip.S] = O1 << ip.S;
    for (int L = ip.S; L >= 1; L--) {
        const int m = ip.m[L - 1];
        b[L - 1] = std::max<long long>(0, (b[L] >> 1) - (2 * m - 1));
        e[L - 1] = ((e[L] - 1) >> 1) + 1;
    }
    long long stride = 0;
    for (int L = 0; L <= ip.S; L++) stride = std::max(stride, e[L] - b[L]);
    stride = (stride + 1) / 2 * 2;
    float2 *cur = (ip.S == 0) ? y : (float2 *)buf(0, (size_t)nstreams * stride * sizeof(float2));
    long long cur_stride = (ip.S == 0) ? y_stride : stride;
    {
        InterpArbParams a{};
        a.x = x; a.hist = hist; a.x_stride = x_stride; a.hcap = kInterpHcap; a.n0 = (long long)n_abs; a.nx = nx;
        a.step = ip.step; a.bits = ip.bits; a.bank = bank_dev;
        a.O_begin = b[0]; a.count = (int)(e[0] - b[0]); a.A = cur; a.A_stride = cur_stride;
        launch(k_interp_arb, dim3((unsigned)std::max(1, std::min(4096, (a.count + 255) / 256)), nstreams), dim3(256), 0, a);
    }
    for (int L = 1; L <= ip.S; L++) {
        InterpHbParams h{};
        h.U = cur; h.U_begin = b[L - 1]; h.U_stride = cur_stride;
        const bool last = (L == ip.S);
        h.V = last ? y : (float2 *)buf(L & 1, (size_t)nstreams * stride * sizeof(float2));
        h.V_stride = last ? y_stride : stride;
        h.V_begin = b[L]; h.count = e[L] - b[L];
        h.m = ip.m[L - 1];
        for (int t = 0; t < 2 * h.m; t++) h.h1[t] = ip.h1[L - 1][t];
        launch(k_interp_hb, dim3((unsigned)std::max<long long>(1, std::min<long long>(8192, (h.count + 255) / 256)), nstreams),
               dim3(256), 0, h);
        cur = h.V; cur_stride = h.V_stride;
    }
    return (O1 - O0) << ip.S;
}

}  // namespace csdr
