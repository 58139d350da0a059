// frontend.cuh -- fused front end of the receive chain:
//     NCO offset mix  ->  msresamp_crcf  (half-band decimator cascade + arbitrary polyphase resampler)
// replacing, for one chunk and in ONE kernel, what the reference does with
//     nco_crcf_mix_block_down/up       (Liquid.chs:793-809, liquid nco.c)
//     msresamp_crcf_execute            (Liquid.chs:76-98,   liquid msresamp.c / msresamp2.c / resamp2.c /
//                                       resamp.fixed.c / firpfb.c)
//
// Time-parallel formulation.  All liquid objects on this path start from zeroed delay lines, so every stage is
// a causal FIR of the absolute sample index and the sequential object state collapses to closed forms:
//   * NCO phase            theta(n) = theta0 + n*dtheta  (mod 2^32)                       [nco.c: uint32 phase]
//   * half-band stage      y[k]     = sum_i h[i] x[2k+1-i],  h = 4m+1 taps, odd taps + centre tap 1
//   * msresamp2            stages run from index S-1 (input rate) down to 0, output scaled by 2^-S
//   * arbitrary resampler  output o is emitted at push k = (o*step)>>24 with branch ((o*step)>>16)&255,
//                          step = round(2^24/rate)                                        [resamp.fixed.c]
// A tile is Tc consecutive outputs of the half-band cascade ("c" samples); the CTA loads the 2^S*Tc new input
// samples plus a halo, keeps every intermediate in shared memory and writes only final outputs.  The halo of
// the first tile of a chunk comes from `hist` (the last hcap raw input samples of the previous chunks), which
// is the only sample state carried between calls.
//
// Shared-memory layout of a level-L buffer (input of half-band stage L-1), D = R of the consuming stage:
//   sample i -> plane (i&1), pair p = i>>1 -> sub-array p % D at index p / D
//   addr = ((i&1)*D + (p % D)) * stride + p / D
// so that a thread producing R consecutive outputs reads every tap with a compile-time sub-array and offset
// and consecutive lanes hit consecutive addresses (conflict-free LDS.64, each sample read ~ (2R+2m-1)/R times
// per output instead of 2m+1).
#pragma once
#include "platform.cuh"
#include "frontend_geom.hpp"

namespace csdr {

struct FrontendParams {
    // chunk; stream s (blockIdx.y) lives at x + s*x_stride, hist + s*hcap, y + s*y_stride
    const float2 *x;        // chunk (device)
    const float2 *hist;     // hcap samples preceding x[0]; hist[hcap-1] is sample n0-1
    float2 *y;              // outputs of this call, y[o'] for o' = 0 .. ny-1
    long long x_stride, y_stride;
    long long n0;           // absolute index of x[0]
    long long nx;           // chunk length
    int hcap;
    // mixer (Liquid.chs:200-205: f>0 mixDown, f<0 mixUp)
    int mix_mode;           // 0 none, 1 down, 2 up
    unsigned theta0, dtheta;
    int quantize;           // 1: 1024-level phase (liquid 1.3.x table NCO), 0: full 32-bit phase
    // half-band cascade; stage s=0 is the LOWEST-rate stage (liquid's stage index)
    int S;
    int m[kMaxStages];          // semi-length of stage s
    int R[kMaxStages];          // outputs per thread slot of stage s (= layout factor of its input buffer)
    int d[kMaxStages + 1];      // lo_L(tile) = (c_lo << L) + d[L]
    int n[kMaxStages + 1];      // samples held at level L
    int stride[kMaxStages + 1]; // sub-array stride of level L (L >= 1)
    int off[kMaxStages + 1];    // float2 offset of level L buffer in dynamic smem
    float taps[kMaxStages][2 * kMaxHbM];   // h1 of stage s: multiplies O[q+u], u = 0..2m-1
    float zeta;                 // 2^-S
    // tiles of Tc c-samples; this call produces c indices [K0, K1)
    long long K0, K1;
    int Tc;
    int ntiles;
    // arbitrary resampler
    unsigned step; int bits;    // npfb = 1<<bits
    const float *bank;          // [npfb][kHsub], bank[i][j] multiplies c[k-j]
    unsigned long long ph0;     // resampler phase (liquid's q->phase) before push K0: output o' of this call has
                                // phase ph0 + o'*step relative to push K0
    int off_bank;               // float offset (in floats) of the bank copy in dynamic smem
    int tma_ok, tma_r;          // k_frontend_direct: a tensor map covers x from sample tma_r on (rows of 16 samples)
    long long tma_rows;         // whole rows it covers
    int smem_bytes;
};

// address of sample i in a level buffer with layout factor D (power of two) and sub-array stride: plane i & 1
// (planes are D * stride + kFePlanePad apart), pair p = i >> 1 -> sub-array p & (D-1), index p / D
template <int D>
__device__ __forceinline__ int fe_addr(int i, int stride)
{
    int p = i >> 1;
    return (i & 1) * (D * stride + kFePlanePad) + (p & (D - 1)) * stride + (p / D);
}
__device__ __forceinline__ int fe_addr_rt(int i, int D, int stride)
{
    const int p = i >> 1;
    const int lg = (D == 8) ? 3 : (D == 4) ? 2 : (D == 2) ? 1 : 0;      // D is 1, 2, 4 or 8
    return (i & 1) * (D * stride + kFePlanePad) + (p & (D - 1)) * stride + (p >> lg);
}

// Phasor with the quantisation mode known at compile time.  The phase index is turned into a float by bit insertion
// (2^23 + k has k in its mantissa), not by an int->float conversion: conversions share the XU pipe with the sin/cos
// evaluations.  `th` is the phase word PLUS HALF A TURN (callers fold the 2^31 into their base phase), so that
// k - K/2 is the signed index and the angle lies in [-pi, pi), where the SFU approximations are most accurate
// (abs err ~2^-21.4).  Q = 1: NCO sine table, K = 1024 levels, th also carries the +2^21 rounding offset;
// Q = 0: K = 2^23, the top 23 phase bits.
constexpr unsigned kFePhaseBias = 0x80000000u;
template <int Q>
__device__ __forceinline__ float2 fe_phasor_q(unsigned th)
{
#ifdef CSDR_EMU
    const unsigned kbits = 0x4B000000u | (Q ? (th >> 22) : (th >> 9));
#else
    const unsigned kbits = Q ? __funnelshift_r(th, 0x4B000000u >> 10, 22) : __funnelshift_r(th, 0x4B000000u >> 23, 9);
#endif
    // revolutions in [-0.5, 0.5): (2^23 + k) / K - (2^23 / K + 0.5), every step exact in fp32
    const float rev = Q ? fmaf(__uint_as_float(kbits), 9.765625e-4f, -8192.5f) : (__uint_as_float(kbits) - 8388608.0f) * 1.1920928955078125e-7f - 0.5f;
    float s, c;
    __sincosf(rev * 6.283185307179586f, &s, &c);
    return cf(c, s);
}

// phasor of the NCO at phase word `th`:  (cos, sin)
__device__ __forceinline__ float2 fe_phasor(unsigned th, int quantize)
{
    return quantize ? fe_phasor_q<1>(th + (1u << 21) + kFePhaseBias) : fe_phasor_q<0>(th + kFePhaseBias);   // NCO(_index): round to 1024 levels
}

// One half-band decimation stage over a tile: n_out outputs, R per thread slot.
//   out[q] = E[q+M] + sum_{u<2M} g[u] * O[q+u]       (E/O = even/odd samples of the input level)
// LAST: write zeta*out to the linear c buffer, else into the next level's (D2, stride2) layout.
template <int M, int R, bool LAST>
__device__ __forceinline__ void fe_stage(const float2 *__restrict__ in, int stride, float2 *__restrict__ out,
                                         int D2, int stride2, int n_out, const float *__restrict__ g_taps,
                                         float zeta)
{
    float g[2 * M];
#pragma unroll
    for (int u = 0; u < 2 * M; u++) g[u] = g_taps[u];
    const float2 *E = in;
    const float2 *O = in + R * stride + kFePlanePad;
    const int nslots = n_out / R;
    for (int t = threadIdx.x; t < nslots; t += blockDim.x) {
        float ar[R], ai[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            float2 e = E[((M + r) % R) * stride + t + (M + r) / R];
            ar[r] = e.x; ai[r] = e.y;
        }
#pragma unroll
        for (int c = 0; c < R + 2 * M - 1; c++) {
            float2 v = O[(c % R) * stride + t + c / R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                int u = c - r;
                if (u >= 0 && u < 2 * M) {
                    ar[r] = fmaf(g[u], v.x, ar[r]);
                    ai[r] = fmaf(g[u], v.y, ai[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            int q = t * R + r;
            if (LAST) out[q] = cf(ar[r] * zeta, ai[r] * zeta);
            else      out[fe_addr_rt(q, D2, stride2)] = cf(ar[r], ai[r]);
        }
    }
}

// generic (any m) stage: one output per thread iteration, taps from the parameter block
template <bool LAST>
__device__ void fe_stage_generic(const float2 *__restrict__ in, int D, int stride, float2 *__restrict__ out,
                                 int D2, int stride2, int n_out, int M, const float *__restrict__ g, float zeta)
{
    for (int q = threadIdx.x; q < n_out; q += blockDim.x) {
        float2 e = in[fe_addr_rt(2 * (q + M), D, stride)];
        float ar = e.x, ai = e.y;
        for (int u = 0; u < 2 * M; u++) {
            float2 v = in[fe_addr_rt(2 * (q + u) + 1, D, stride)];
            ar = fmaf(g[u], v.x, ar);
            ai = fmaf(g[u], v.y, ai);
        }
        if (LAST) out[q] = cf(ar * zeta, ai * zeta);
        else      out[fe_addr_rt(q, D2, stride2)] = cf(ar, ai);
    }
}

template <bool LAST>
__device__ __forceinline__ void fe_stage_dispatch(const FrontendParams &p, int s, float2 *smem)
{
    const float2 *in = smem + p.off[s + 1];
    float2 *out = smem + p.off[s];
    const int stride = p.stride[s + 1];
    const int D2 = LAST ? 1 : p.R[s - (LAST ? 0 : 1)];
    const int stride2 = p.stride[s];
    const int n_out = p.n[s];
    const float *g = p.taps[s];
    const int M = p.m[s], R = p.R[s];
    if (M == 3 && R == 8)        fe_stage<3, 8, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else if (M == 5 && R == 8)   fe_stage<5, 8, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else if (M == 5 && R == 4)   fe_stage<5, 4, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else if (M == 10 && R == 8)  fe_stage<10, 8, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else if (M == 10 && R == 4)  fe_stage<10, 4, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else if (M == 10 && R == 2)  fe_stage<10, 2, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else if (M == 3 && R == 4)   fe_stage<3, 4, LAST>(in, stride, out, D2, stride2, n_out, g, p.zeta);
    else                         fe_stage_generic<LAST>(in, R, stride, out, D2, stride2, n_out, M, g, p.zeta);
}

__global__ void __launch_bounds__(256, 3) k_frontend(const CSDR_GRID_CONSTANT FrontendParams p)
{
    CSDR_DYN_SMEM(smem_raw);
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    float *bank_s = reinterpret_cast<float *>(smem_raw) + p.off_bank;
    __shared__ long long s_orange[2];

    // polyphase bank -> smem once per CTA, rows padded to kHsub+1 floats (odd stride: conflict-light)
    const int npfb = 1 << p.bits;
    for (int i = threadIdx.x; i < npfb * kHsub; i += blockDim.x) {
        int row = i / kHsub, col = i - row * kHsub;
        bank_s[row * (kHsub + 1) + col] = p.bank[i];
    }

    const int S = p.S;
    const float2 *xs = p.x + (long long)blockIdx.y * p.x_stride;
    const float2 *hs = p.hist + (long long)blockIdx.y * p.hcap;
    float2 *ys = p.y + (long long)blockIdx.y * p.y_stride;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const long long kArel = (long long)tile * p.Tc;                 // pushes relative to K0
        const long long kBrel = min(kArel + (long long)p.Tc, p.K1 - p.K0);
        const long long kA = p.K0 + kArel;
        const long long c_lo = kA - kHcPad;
        if (threadIdx.x == 0) {
            // outputs emitted by pushes [kA, kB): o' with kArel*2^24 <= ph0 + o'*step < kBrel*2^24
            const unsigned long long st = p.step;
            const unsigned long long a = (unsigned long long)kArel << 24, b = (unsigned long long)kBrel << 24;
            s_orange[0] = (a > p.ph0) ? (long long)((a - p.ph0 + st - 1) / st) : 0;
            s_orange[1] = (b > p.ph0) ? (long long)((b - p.ph0 + st - 1) / st) : 0;
        }

        // ---- load + mix the top level (raw input samples) ----
        {
            const long long lo = c_lo * (1LL << S) + p.d[S];    // absolute index of local sample 0
            float2 *dst = smem + p.off[S];
            const int nS = p.n[S];
            const int D = S ? p.R[S - 1] : 1, stride = p.stride[S];
            for (int i = threadIdx.x; i < nS; i += blockDim.x) {
                long long g = lo + i;
                long long rel = g - p.n0;
                float2 v = cf(0.f, 0.f);
                if (rel >= 0) { if (rel < p.nx) v = xs[rel]; }
                else if (rel >= -(long long)p.hcap) v = hs[p.hcap + rel];
                if (p.mix_mode) {
                    float2 w = fe_phasor(p.theta0 + (unsigned)g * p.dtheta, p.quantize);
                    float s = (p.mix_mode == 1) ? -w.y : w.y;     // down: multiply by conj
                    v = cf(v.x * w.x - v.y * s, v.y * w.x + v.x * s);
                }
                if (S) dst[fe_addr_rt(i, D, stride)] = v;
                else   dst[i] = v;
            }
        }
        __syncthreads();

        // ---- half-band cascade, input-rate stage (S-1) first ----
        for (int s = S - 1; s >= 1; s--) {
            fe_stage_dispatch<false>(p, s, smem);
            __syncthreads();
        }
        if (S >= 1) {
            fe_stage_dispatch<true>(p, 0, smem);
            __syncthreads();
        }

        // ---- arbitrary resampler: one output per thread iteration ----
        {
            const float2 *cbuf = smem + p.off[0];
            const long long oA = s_orange[0], oB = s_orange[1];
            const unsigned mask = (unsigned)npfb - 1u;
            for (long long o = oA + threadIdx.x; o < oB; o += blockDim.x) {
                unsigned long long ph = p.ph0 + (unsigned long long)o * p.step;
                int k = (int)((long long)(ph >> 24) - kArel) + kHcPad;
                const float *h = bank_s + ((unsigned)(ph >> (24 - p.bits)) & mask) * (kHsub + 1);
                float ar = 0.f, ai = 0.f;
#pragma unroll
                for (int j = 0; j < kHsub; j++) {
                    float2 v = cbuf[k - j];
                    ar = fmaf(h[j], v.x, ar);
                    ai = fmaf(h[j], v.y, ai);
                }
                ys[o] = cf(ar, ai);
            }
        }
        __syncthreads();   // smem is reused by the next tile
    }
}

// hist_out <- last hcap samples of concat(hist_in, x[0..nx))
__global__ void k_hist_update(const float2 *__restrict__ hist_in, float2 *__restrict__ hist_out,
                              const float2 *__restrict__ x, long long x_stride, long long nx, int hcap)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= hcap) return;
    const long long s = blockIdx.y;
    long long rel = (long long)j - hcap + nx;      // index into x of the sample that lands at hist_out[j]
    hist_out[s * hcap + j] = (rel >= 0) ? x[s * x_stride + rel] : hist_in[s * hcap + hcap + rel];
}

// stand-alone NCO mixer (nco_crcf_mix_block_down/up, Liquid.chs:793-809)
__global__ void k_nco_mix(const float2 *__restrict__ x, float2 *__restrict__ y, long long n, unsigned theta0,
                          unsigned dtheta, int quantize, int up)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float2 v = x[i];
        float2 w = fe_phasor(theta0 + (unsigned)i * dtheta, quantize);
        float s = up ? w.y : -w.y;
        y[i] = cf(v.x * w.x - v.y * s, v.y * w.x + v.x * s);
    }
}

}  // namespace csdr
