// pfb.cuh -- firpfbch_crcf analysis channelizer (critically sampled, M in -> M out per frame)
// replacing firpfbchChan (Liquid.chs:827-862): nco pre-rotation of the chunk, then per frame
// firpfbch_crcf_analyzer_execute (liquid src/multichannel/src/firpfbch.c) and the Haskell per-element pokes
// that transpose the result to channel-major [M][nframes].
//
// Closed form of the sequential object (all windows start at zero):
//   X_t[n] = sum_{k<P} h[(M-1-n) + k*M] * xr[(t-k)*M + n]        P = 2m taps per branch, xr = pre-rotated input
//   y_t[c] = sum_n X_t[n] exp(-j 2 pi c n / M)                   (unnormalised forward DFT)
// One CTA computes F consecutive frames: the polyphase sums go straight into shared memory, the M-point DFT runs
// there (radix-2 for powers of two, direct otherwise) and the result is stored transposed so that each channel
// receives a contiguous run of F samples.
#pragma once
#ifndef CSDR_EMU
#include <cooperative_groups.h>
#endif
#include "platform.cuh"

namespace csdr {

// power of a channel sample exactly as k_be_prep computes it
__device__ __forceinline__ float pfb_power(float2 v) { return __fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)); }

struct PfbParams {
    const float2 *xr;       // pre-rotated samples: xr[0 .. (P-1)*M) = history frames, then nf*M new samples
    float2 *y;              // [M][y_stride] channel-major output, frame t at column t
    long long y_stride;
    int M, P, nf, F;        // channels, taps per branch, frames in this call, frames per CTA
    int log2M;              // >= 0 when M is a power of two, else -1
    const float *h;         // prototype, P*M taps
    const float2 *tw;       // M twiddles exp(-j 2 pi t / M)
    float *pw; long long pw_stride;   // optional: |y|^2, same layout (input of the per-channel AGC gain loop)
    int hop;                // input samples per frame: M (firpfbch), M/2 (firpfbch2, history (P-1) M + M/2)
    int over2, parity0;     // firpfbch2 analyzer: y_t[c] *= (-1)^(c t) exp(-j 2 pi c / M) / M, t counted from parity0
    float scale;
};

__device__ __forceinline__ unsigned pfb_bitrev(unsigned v, int bits) { return bits ? (__brev(v) >> (32 - bits)) : 0u; }

__global__ void __launch_bounds__(256) k_pfb(const PfbParams p)
{
    CSDR_DYN_SMEM(smem_raw);
    float2 *buf = reinterpret_cast<float2 *>(smem_raw);          // [F][M] (+ second half for the direct DFT)
    const int M = p.M, P = p.P;
    const int t0 = blockIdx.x * p.F;
    const int nfr = min(p.F, p.nf - t0);
    const int total = nfr * M;
    const bool pow2 = p.log2M >= 0;
    float2 *xbuf = pow2 ? buf : buf + p.F * M;                   // direct DFT reads X from the second half

    // polyphase filter: X[f][n]; consecutive threads -> consecutive n (coalesced reads of xr, h)
    for (int e = threadIdx.x; e < total; e += blockDim.x) {
        const int f = e / M, n = e - f * M;
        const float2 *src = p.xr + (long long)(P - 1) * M + (long long)(t0 + f) * p.hop + n;     // newest sample of this branch
        const float *hh = p.h + (M - 1 - n);
        float ar = 0.f, ai = 0.f;
        for (int k = 0; k < P; k++) {
            float2 v = src[-(long long)k * M];
            float c = hh[k * M];
            ar = fmaf(c, v.x, ar);
            ai = fmaf(c, v.y, ai);
        }
        const int pos = pow2 ? (int)pfb_bitrev((unsigned)n, p.log2M) : n;
        xbuf[f * M + pos] = cf(ar, ai);
    }
    __syncthreads();

    if (pow2) {
        // in-place radix-2 DIT over every frame of the tile
        for (int len = 2; len <= M; len <<= 1) {
            const int half = len >> 1, tws = M / len;
            const int nb = nfr * (M >> 1);
            for (int b = threadIdx.x; b < nb; b += blockDim.x) {
                const int f = b / (M >> 1), r = b - f * (M >> 1);
                const int grp = r / half, j = r - grp * half;
                float2 *a = buf + f * M + grp * len + j;
                const float2 w = p.tw[j * tws];
                const float2 u = a[0], v = a[half];
                const float tr = v.x * w.x - v.y * w.y, ti = v.x * w.y + v.y * w.x;
                a[0] = cf(u.x + tr, u.y + ti);
                a[half] = cf(u.x - tr, u.y - ti);
            }
            __syncthreads();
        }
    } else {
        for (int e = threadIdx.x; e < total; e += blockDim.x) {
            const int f = e / M, c = e - f * M;
            const float2 *X = xbuf + f * M;
            float ar = 0.f, ai = 0.f;
            int t = 0;
            for (int n = 0; n < M; n++) {
                const float2 w = p.tw[t], v = X[n];
                ar += v.x * w.x - v.y * w.y;
                ai += v.x * w.y + v.y * w.x;
                t += c; if (t >= M) t -= M;
            }
            buf[f * M + c] = cf(ar, ai);
        }
        __syncthreads();
    }

    // transposed store: consecutive threads -> consecutive frames of one channel
    for (int e = threadIdx.x; e < M * nfr; e += blockDim.x) {
        const int c = e / nfr, f = e - c * nfr;
        float2 v = buf[f * M + c];
        if (p.over2) {
            // firpfbch2: the analyzer's backward DFT over the window slots and its alternating commutator, folded
            // into a per-channel factor of the forward DFT (closed form at the top of the firpfbch2 section below)
            const float2 w = p.tw[c];
            const float sg = ((p.parity0 + t0 + f) & c & 1) ? -p.scale : p.scale;
            v = cf((v.x * w.x - v.y * w.y) * sg, (v.x * w.y + v.y * w.x) * sg);
        }
        p.y[(long long)c * p.y_stride + t0 + f] = v;
        if (p.pw) p.pw[(long long)c * p.pw_stride + t0 + f] = pfb_power(v);
    }
}

// ---- firpfbch2_crcf analyzer (2x oversampled: M/2 samples in, M channels out per frame) -----------
// liquid's sequential object (src/multichannel/src/firpfbch2.c: two half-frames of windows filled alternately, branch
// i reading window (offset + i) mod M, backward DFT, 1/M) has the closed form, with n_t = (t+1) M/2 samples received:
//   u_t[s] = sum_{a<P} h[M a + s] x[n_t - 1 - s - a M]              (the polyphase sums of the last P M samples)
//   y_t[c] = (1/M) (-1)^(c t) sum_s u_t[s] exp(+j 2 pi s c / M)
// With n = M-1-s the sums are exactly k_pfb's X_t[n] on a window that advances by M/2, and
//   y_t[c] = (1/M) (-1)^(c t) exp(-j 2 pi c / M) * sum_n X_t[n] exp(-j 2 pi c n / M)
// so k_pfb runs it with hop = M/2 and the factor applied in its store (over2 = 1).

// ---- small power-of-two M (2..32): one thread per frame -------------------------------------------------------
// The CTA stages F + P - 1 consecutive frames in shared memory once (coalesced 16-byte loads; rows padded to M + 2
// samples so that threads one frame apart hit different 16-byte banks).  Thread f then evaluates all M polyphase
// branches of frame f from P rows (taps straight from the kernel-parameter constant bank, packed FP32 FMAs), runs the
// M-point DFT in registers (radix-2 DIT, every index a compile-time constant) and stores channel-major: for each
// channel the warp writes 32 consecutive frames = 256 contiguous bytes.
constexpr int kPfbTileP = 14;              // 2 m taps per branch (m = 7, Liquid.chs:865)
constexpr int kPfbTileF = 256;             // frames per CTA = threads per CTA
constexpr int kPfbTileMaxM = 32;

struct PfbTileParams {
    const float2 *xr; float2 *y; long long y_stride; int nf;
    float *pw; long long pw_stride;        // optional: |y|^2, same layout
    int ocs, oco;                          // frame f is output column f * ocs + oco (firpfbch2: the even / odd frames of the
                                           // half-frame hop are two critically sampled passes, ocs = 2)
    int over2; float sc_even, sc_odd;      // firpfbch2: y[c] *= exp(-j 2 pi c / M) * (c even ? sc_even : sc_odd)
    float2 tw[kPfbTileMaxM / 2];           // exp(-j 2 pi t / M)
    float h[kPfbTileP * kPfbTileMaxM];     // h[k * M + n] = prototype[(M - 1 - n) + k * M]
};

template <int LM> __host__ __device__ constexpr int pfb_rev(int i)
{
    int r = 0;
    for (int b = 0; b < LM; b++) r |= ((i >> b) & 1) << (LM - 1 - b);
    return r;
}

// Radix-2 DIT butterflies of an M-point DFT held in registers, every index a template constant: butterfly i of stage s,
// then the next one (the nested-loop form was left partly rolled by the compiler, which then indexed the register array
// through chains of predicated moves: 28 % of the kernel's instructions).  Twiddles 1 and -j cost no multiplication.
template <int LM, int s, int i>
__device__ __forceinline__ void pfb_dit_bf(float2 (&b)[1 << LM], const float2 *__restrict__ tw)
{
    constexpr int M = 1 << LM, half = 1 << (s - 1), j = i & (half - 1), lo = ((i >> (s - 1)) << s) + j, hi = lo + half;
    constexpr int t = j * (M >> s);                        // twiddle exp(-j 2 pi t / M)
    const float2 u = b[lo], v = b[hi];
    float2 r;
    if constexpr (t == 0) r = v;
    else if constexpr (4 * t == M) r = cf(v.y, -v.x);
    else { const float2 w = tw[t]; r = cf(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x); }
    b[lo] = cf(u.x + r.x, u.y + r.y);
    b[hi] = cf(u.x - r.x, u.y - r.y);
    if constexpr (i + 1 < M / 2) pfb_dit_bf<LM, s, i + 1>(b, tw);
    else if constexpr (s < LM) pfb_dit_bf<LM, s + 1, 0>(b, tw);
}

template <int LM>
__global__ void __launch_bounds__(kPfbTileF) k_pfb_tile(const CSDR_GRID_CONSTANT PfbTileParams p)
{
    constexpr int M = 1 << LM, P = kPfbTileP, F = kPfbTileF, RS = M + 2;
    CSDR_DYN_SMEM(smem_raw);
    float2 *in = reinterpret_cast<float2 *>(smem_raw);           // [(F + P - 1)][RS]
    const int t0 = blockIdx.x * F;
    const int nfr = min(F, p.nf - t0);
    // rows t0 .. t0 + nfr + P - 2 of xr (row r of xr = history or new frame r - (P-1))
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.xr + (long long)t0 * M);
        const int pairs = (nfr + P - 1) * (M / 2);
        for (int e = threadIdx.x; e < pairs; e += F) {
            const int r = e >> (LM - 1), c2 = e & (M / 2 - 1);
            *reinterpret_cast<float4 *>(in + r * RS + 2 * c2) = src[e];
        }
    }
    __syncthreads();
    const int f = threadIdx.x;
    if (f >= nfr) return;
    float2 a[M];
#pragma unroll
    for (int n = 0; n < M; n++) a[n] = cf(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < P; k++) {
        const float4 *row = reinterpret_cast<const float4 *>(in + (f + P - 1 - k) * RS);
#pragma unroll
        for (int j = 0; j < M / 2; j++) {
            const float4 q = row[j];
            // (scalar FMAs take the tap straight from the constant bank; the packed form needs it copied into a register
            // pair first: three instructions instead of two)
            const float h0 = p.h[k * M + 2 * j], h1 = p.h[k * M + 2 * j + 1];
            a[2 * j].x = fmaf(h0, q.x, a[2 * j].x); a[2 * j].y = fmaf(h0, q.y, a[2 * j].y);
            a[2 * j + 1].x = fmaf(h1, q.z, a[2 * j + 1].x); a[2 * j + 1].y = fmaf(h1, q.w, a[2 * j + 1].y);
        }
    }
    // M-point forward DFT, radix-2 decimation in time
    float2 b[M];
#pragma unroll
    for (int i = 0; i < M; i++) b[i] = a[pfb_rev<LM>(i)];
#ifndef CSDR_PFB_SKIP_DFT       // ablation (profiles/r02_tensorcore_question.txt): what a DFT that costs nothing would gain
    pfb_dit_bf<LM, 1, 0>(b, p.tw);
#endif
    if (p.over2) {
        // firpfbch2: the analyzer's backward DFT over the window slots and its alternating commutator are a per-channel
        // factor of the forward DFT (closed form at the top of the firpfbch2 section); W^(c + M/2) = -W^c
#pragma unroll
        for (int c = 0; c < M; c++) {
            const float2 w = p.tw[c & (M / 2 - 1)];
            const float sc = ((c & 1) ? p.sc_odd : p.sc_even) * ((M > 1 && c >= M / 2) ? -1.f : 1.f);
            b[c] = cf((b[c].x * w.x - b[c].y * w.y) * sc, (b[c].x * w.y + b[c].y * w.x) * sc);
        }
    }
    const long long col = (long long)(t0 + f) * p.ocs + p.oco;
    float2 *yo = p.y + col;
#pragma unroll
    for (int c = 0; c < M; c++) yo[(long long)c * p.y_stride] = b[c];
    if (p.pw) {
        float *po = p.pw + col;
#pragma unroll
        for (int c = 0; c < M; c++) po[(long long)c * p.pw_stride] = pfb_power(b[c]);
    }
}

// ---- even M <= 24 that is not a power of two (the reference's published run has 20 channels, README.md:182-193) -----
// k_pfb_tile with the register DFT as one radix-2 split and two direct M/2-point DFTs: y[c] = E[c] + W^c O[c],
// y[c + M/2] = E[c] - W^c O[c]; every twiddle exponent is a template constant (table in the constant bank, W^(t + M/2) = -W^t;
// 1, -1, -j, +j cost no multiplication).  M^2/2 + M complex MACs per frame (220 at M = 20), all in registers.
template <int M, int t>
__device__ __forceinline__ void pfb_mac_tw(float2 &acc, float2 a, const float2 *__restrict__ tw)
{
    constexpr int tt = ((t % M) + M) % M;
    if constexpr (tt == 0) { acc.x += a.x; acc.y += a.y; }
    else if constexpr (2 * tt == M) { acc.x -= a.x; acc.y -= a.y; }
    else if constexpr (4 * tt == M) { acc.x += a.y; acc.y -= a.x; }
    else if constexpr (4 * tt == 3 * M) { acc.x -= a.y; acc.y += a.x; }
    else if constexpr (tt < M / 2) { const float2 w = tw[tt]; acc.x += a.x * w.x - a.y * w.y; acc.y += a.x * w.y + a.y * w.x; }
    else { const float2 w = tw[tt - M / 2]; acc.x -= a.x * w.x - a.y * w.y; acc.y -= a.x * w.y + a.y * w.x; }
}
template <int M, int c, int n>
__device__ __forceinline__ void pfb_dft_half(const float2 (&a)[M], float2 &e, float2 &o, const float2 *__restrict__ tw)
{
    pfb_mac_tw<M, 2 * c * n>(e, a[2 * n], tw);
    pfb_mac_tw<M, 2 * c * n>(o, a[2 * n + 1], tw);
    if constexpr (n + 1 < M / 2) pfb_dft_half<M, c, n + 1>(a, e, o, tw);
}
template <int M, int c>
__device__ __forceinline__ void pfb_dft_any(const float2 (&a)[M], float2 (&y)[M], const float2 *__restrict__ tw)
{
    float2 e = cf(0.f, 0.f), o = cf(0.f, 0.f), t = cf(0.f, 0.f);
    pfb_dft_half<M, c, 0>(a, e, o, tw);
    pfb_mac_tw<M, c>(t, o, tw);
    y[c] = cf(e.x + t.x, e.y + t.y);
    y[c + M / 2] = cf(e.x - t.x, e.y - t.y);
    if constexpr (c + 1 < M / 2) pfb_dft_any<M, c + 1>(a, y, tw);
}

template <int M>
__global__ void __launch_bounds__(kPfbTileF) k_pfb_tile_any(const CSDR_GRID_CONSTANT PfbTileParams p)
{
    static_assert(M % 2 == 0 && M <= kPfbTileMaxM, "even channel counts up to 32");
    constexpr int P = kPfbTileP, F = kPfbTileF, RS = M + 2, H = M / 2;
    CSDR_DYN_SMEM(smem_raw);
    float2 *in = reinterpret_cast<float2 *>(smem_raw);           // [(F + P - 1)][RS]
    const int t0 = blockIdx.x * F;
    const int nfr = min(F, p.nf - t0);
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.xr + (long long)t0 * M);
        const int pairs = (nfr + P - 1) * H;
        for (int e = threadIdx.x; e < pairs; e += F) {
            const int r = e / H, c2 = e - r * H;
            *reinterpret_cast<float4 *>(in + r * RS + 2 * c2) = src[e];
        }
    }
    __syncthreads();
    const int f = threadIdx.x;
    if (f >= nfr) return;
    float2 a[M];
#pragma unroll
    for (int n = 0; n < M; n++) a[n] = cf(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < P; k++) {
        const float4 *row = reinterpret_cast<const float4 *>(in + (f + P - 1 - k) * RS);
#pragma unroll
        for (int j = 0; j < H; j++) {
            const float4 q = row[j];
            const float h0 = p.h[k * M + 2 * j], h1 = p.h[k * M + 2 * j + 1];
            a[2 * j].x = fmaf(h0, q.x, a[2 * j].x); a[2 * j].y = fmaf(h0, q.y, a[2 * j].y);
            a[2 * j + 1].x = fmaf(h1, q.z, a[2 * j + 1].x); a[2 * j + 1].y = fmaf(h1, q.w, a[2 * j + 1].y);
        }
    }
    float2 b[M];
    pfb_dft_any<M, 0>(a, b, p.tw);
    if (p.over2) {
#pragma unroll
        for (int c = 0; c < M; c++) {
            const float2 w = p.tw[c % H];
            const float sc = ((c & 1) ? p.sc_odd : p.sc_even) * ((c >= H) ? -1.f : 1.f);
            b[c] = cf((b[c].x * w.x - b[c].y * w.y) * sc, (b[c].x * w.y + b[c].y * w.x) * sc);
        }
    }
    const long long col = (long long)(t0 + f) * p.ocs + p.oco;
    float2 *yo = p.y + col;
#pragma unroll
    for (int c = 0; c < M; c++) yo[(long long)c * p.y_stride] = b[c];
    if (p.pw) {
        float *po = p.pw + col;
#pragma unroll
        for (int c = 0; c < M; c++) po[(long long)c * p.pw_stride] = pfb_power(b[c]);
    }
}

// ---- M = 8, 16: two frames per thread ----------------------------------------------------------------------------
// k_pfb_tile is bound by the shared-memory data pipe: every frame reads its 14 rows (81 % of the pipe's peak at M = 16,
// profiles/r02_tensorcore_question.txt).  Here thread t evaluates frames 2t and 2t + 1 together: row 2t + j is tap 13 - j
// of the first and tap 14 - j of the second, so 15 rows are read for two frames instead of 28.  Rows are stored unpadded
// with an XOR swizzle of their 16-byte chunks (chunk q of the tile at (q & ~7) | ((q & 7) ^ ((q >> 3 >> (log2 M - 3)) & 7)):
// the eight threads of a quarter-warp, two rows apart, hit eight different 16-byte bank groups, and so do the staging
// stores.  The two frames of a thread are neighbours in the channel-major output: one 16-byte store per channel.
template <int LM> __device__ __forceinline__ int pfb_tile2_phys(int q) { return (q & ~7) | ((q & 7) ^ (((q >> 3) >> (LM - 3)) & 7)); }

template <int LM>
__global__ void __launch_bounds__(kPfbTileF / 2) k_pfb_tile2(const CSDR_GRID_CONSTANT PfbTileParams p)
{
    static_assert(LM == 3 || LM == 4, "two frames per thread: 4 M accumulator registers");
    constexpr int M = 1 << LM, P = kPfbTileP, F = kPfbTileF, CH = M / 2;      // CH 16-byte chunks (sample pairs) per row
    CSDR_DYN_SMEM(smem_raw);
    float4 *in4 = reinterpret_cast<float4 *>(smem_raw);           // [(F + P - 1) * CH] chunks, swizzled
    const int t0 = blockIdx.x * F;
    const int nfr = min(F, p.nf - t0);
    {
        const float4 *src = reinterpret_cast<const float4 *>(p.xr + (long long)t0 * M);
        const int chunks = (nfr + P - 1) * CH;
        for (int e = threadIdx.x; e < chunks; e += F / 2) in4[pfb_tile2_phys<LM>(e)] = src[e];
    }
    __syncthreads();
    const int t = threadIdx.x, f = 2 * t;
    if (f >= nfr) return;
    float2 a[M], b[M];                                              // polyphase sums of frames f and f + 1
#pragma unroll
    for (int n = 0; n < M; n++) { a[n] = cf(0.f, 0.f); b[n] = cf(0.f, 0.f); }
#pragma unroll
    for (int j = 0; j < P + 1; j++) {
        // row f + j: tap P - 1 - j of frame f (j < P), tap P - j of frame f + 1 (j >= 1)
        const int q0 = (f + j) * CH;
#pragma unroll
        for (int c = 0; c < CH; c++) {
            const float4 q = in4[pfb_tile2_phys<LM>(q0 + c)];
            if (j < P) {
                const float h0 = p.h[(P - 1 - j) * M + 2 * c], h1 = p.h[(P - 1 - j) * M + 2 * c + 1];
                a[2 * c].x = fmaf(h0, q.x, a[2 * c].x); a[2 * c].y = fmaf(h0, q.y, a[2 * c].y);
                a[2 * c + 1].x = fmaf(h1, q.z, a[2 * c + 1].x); a[2 * c + 1].y = fmaf(h1, q.w, a[2 * c + 1].y);
            }
            if (j >= 1) {
                const float h0 = p.h[(P - j) * M + 2 * c], h1 = p.h[(P - j) * M + 2 * c + 1];
                b[2 * c].x = fmaf(h0, q.x, b[2 * c].x); b[2 * c].y = fmaf(h0, q.y, b[2 * c].y);
                b[2 * c + 1].x = fmaf(h1, q.z, b[2 * c + 1].x); b[2 * c + 1].y = fmaf(h1, q.w, b[2 * c + 1].y);
            }
        }
    }
    // the two M-point DFTs (bit-reversed input order, then the radix-2 butterflies of k_pfb_tile)
    float2 ya[M], yb[M];
#pragma unroll
    for (int i = 0; i < M; i++) { ya[i] = a[pfb_rev<LM>(i)]; yb[i] = b[pfb_rev<LM>(i)]; }
    pfb_dit_bf<LM, 1, 0>(ya, p.tw);
    pfb_dit_bf<LM, 1, 0>(yb, p.tw);
    if (p.over2) {
#pragma unroll
        for (int c = 0; c < M; c++) {
            const float2 w = p.tw[c & (M / 2 - 1)];
            const float sc = ((c & 1) ? p.sc_odd : p.sc_even) * ((c >= M / 2) ? -1.f : 1.f);
            ya[c] = cf((ya[c].x * w.x - ya[c].y * w.y) * sc, (ya[c].x * w.y + ya[c].y * w.x) * sc);
            yb[c] = cf((yb[c].x * w.x - yb[c].y * w.y) * sc, (yb[c].x * w.y + yb[c].y * w.x) * sc);
        }
    }
    const bool two = f + 1 < nfr;
    const long long col = (long long)(t0 + f) * p.ocs + p.oco;
    float2 *yo = p.y + col;
    // frames f and f + 1 are neighbouring columns: one 16-byte store per channel where the layout allows it
    const bool vec = two && p.ocs == 1 && ((reinterpret_cast<uintptr_t>(yo) | (uintptr_t)(p.y_stride * sizeof(float2))) & 15) == 0;
    if (vec) {
#pragma unroll
        for (int c = 0; c < M; c++) *reinterpret_cast<float4 *>(yo + (long long)c * p.y_stride) = make_float4(ya[c].x, ya[c].y, yb[c].x, yb[c].y);
    } else {
#pragma unroll
        for (int c = 0; c < M; c++) {
            yo[(long long)c * p.y_stride] = ya[c];
            if (two) yo[(long long)c * p.y_stride + p.ocs] = yb[c];
        }
    }
    if (p.pw) {
        float *po = p.pw + col;
        const bool vec2 = two && p.ocs == 1 && ((reinterpret_cast<uintptr_t>(po) | (uintptr_t)(p.pw_stride * sizeof(float))) & 7) == 0;
        if (vec2) {
#pragma unroll
            for (int c = 0; c < M; c++) *reinterpret_cast<float2 *>(po + (long long)c * p.pw_stride) = cf(pfb_power(ya[c]), pfb_power(yb[c]));
        } else {
#pragma unroll
            for (int c = 0; c < M; c++) {
                po[(long long)c * p.pw_stride] = pfb_power(ya[c]);
                if (two) po[(long long)c * p.pw_stride + p.ocs] = pfb_power(yb[c]);
            }
        }
    }
}

// ---- large power-of-two M (128..1024): a CTA slides over its frames with a ring of rows --------------------------
// Shared memory holds the last P + 1 = 15 rows of M samples (ring), two M-point DFT buffers and a [TF][M] output
// tile.  Two frames are processed per iteration by the two halves of the CTA (M/4 threads each): the two new rows are
// loaded once (coalesced, prefetched through registers one iteration ahead) and thread i of a half evaluates the four
// polyphase branches n = i + j M/4 of its frame (4 x 14 taps in registers for the whole kernel).  Those four values
// are exactly the inputs of one radix-4 decimation-in-frequency butterfly, so the first DFT level runs in registers;
// further radix-4 levels run in shared memory (consecutive threads = consecutive addresses) until the blocks are 16
// or 32 long, and those are finished as register FFTs, one thread per block (blocks padded by two samples so that the
// 16-byte loads of neighbouring threads fall into different banks).  DIF leaves the frequencies digit-reversed; the
// permutation is folded into the (padded) output tile, which is written out every TF frames, TF consecutive frames
// (64 bytes) per channel.  Each input sample is read from HBM once (plus 13 rows of halo per CTA), each output once.
constexpr int kPfbRingP = 14, kPfbRingTF = 8, kPfbRingCPT = 4, kPfbRingRows = kPfbRingP + 1;

struct PfbRingParams {
    const float2 *xr; float2 *y; long long y_stride;
    float *pw; long long pw_stride;        // optional: |y|^2, same layout
    int nf, T;               // frames in this call, frames per CTA (multiple of TF)
    int M, log2M;
    const float *h;          // prototype, P*M taps
    const float2 *tw;        // M twiddles exp(-j 2 pi t / M)
};
// log2 of the register-FFT size: 4 when log2 M is even, 5 when it is odd
inline int pfb_ring_lfz(int log2M) { return (log2M & 1) ? 5 : 4; }
inline size_t pfb_ring_work(int M, int lfz) { return (size_t)M + 2 * ((size_t)M >> lfz); }          // padded DFT buffer
inline size_t pfb_ring_orow(int M) { return (size_t)M + ((size_t)M >> 4) + 2; }                     // padded tile row
inline size_t pfb_ring_smem(int M, int log2M)
{
    return sizeof(float2) * ((size_t)kPfbRingRows * M + 2 * pfb_ring_work(M, pfb_ring_lfz(log2M)) + kPfbRingTF * pfb_ring_orow(M) + M);
}

__device__ __forceinline__ float2 pfb_cmul(float2 a, float2 w) { return cf(a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x); }

// radix-4 DIF butterfly on x[0..3] (inputs N/4 apart), outputs y_q for frequencies = q (mod 4), twiddled by w^q
__device__ __forceinline__ void pfb_dif4(float2 (&x)[4], float2 w1, float2 w2, float2 w3)
{
    const float2 t0 = cf(x[0].x + x[2].x, x[0].y + x[2].y), t1 = cf(x[0].x - x[2].x, x[0].y - x[2].y);
    const float2 t2 = cf(x[1].x + x[3].x, x[1].y + x[3].y), t3 = cf(x[1].y - x[3].y, x[3].x - x[1].x);   // -j (x1 - x3)
    x[0] = cf(t0.x + t2.x, t0.y + t2.y);
    x[1] = pfb_cmul(cf(t1.x + t3.x, t1.y + t3.y), w1);
    x[2] = pfb_cmul(cf(t0.x - t2.x, t0.y - t2.y), w2);
    x[3] = pfb_cmul(cf(t1.x - t3.x, t1.y - t3.y), w3);
}

template <int LFZ>
__global__ void __launch_bounds__(512, 1) k_pfb_ring(const PfbRingParams p)
{
    constexpr int P = kPfbRingP, TF = kPfbRingTF, CPT = kPfbRingCPT, RR = kPfbRingRows, FZ = 1 << LFZ;
    CSDR_DYN_SMEM(smem_raw);
    const int M = p.M, lm = p.log2M, NT = M / CPT;                // blockDim.x = 2 NT
    const int half_id = threadIdx.x / NT, tid = threadIdx.x - half_id * NT, gtid = threadIdx.x;
    const int WP = M + 2 * (M >> LFZ), OR = M + (M >> 4) + 2;
    float2 *ring = reinterpret_cast<float2 *>(smem_raw);          // [RR][M]
    float2 *work = ring + RR * M + half_id * WP;                   // [2][WP], element i at i + 2 (i >> LFZ)
    float2 *obuf = ring + RR * M + 2 * WP;                         // [TF][OR], channel c at c + (c >> 4)
    float2 *stw = obuf + TF * OR;                                  // [M] twiddles
    const int t0 = blockIdx.x * p.T, t1 = min(t0 + p.T, p.nf);
    if (t0 >= t1) return;
    for (int i = gtid; i < M; i += 2 * NT) stw[i] = p.tw[i];
    // taps of this thread's columns: hh[j][k] = h[(M-1-n) + k M], n = tid + j NT
    float hh[CPT][P];
#pragma unroll
    for (int j = 0; j < CPT; j++) {
        const int n = tid + j * NT;
#pragma unroll
        for (int k = 0; k < P; k++) hh[j][k] = p.h[(M - 1 - n) + k * M];
    }
    // history of the first frame: xr rows t0 .. t0 + P - 2 -> ring slots row % RR
    for (int r = t0; r < t0 + P - 1; r++) {
        const float4 *src = reinterpret_cast<const float4 *>(p.xr + (long long)r * M);
        float4 *dst = reinterpret_cast<float4 *>(ring + (r % RR) * M);
        for (int e = gtid; e < M / 2; e += 2 * NT) dst[e] = src[e];
    }
    // the two rows of the NEXT iteration travel through registers while the current pair of frames is computed:
    // 2 rows = M float4, 2 NT threads, two float4 each (half h fetches the row of its own frame)
    float4 nx0 = make_float4(0.f, 0.f, 0.f, 0.f), nx1 = nx0;
    const long long last_row = (long long)p.nf + P - 2;           // last row that exists in xr
    auto fetch_rows = [&](int t) {
        const long long row = (long long)t + P - 1 + half_id;
        if (row <= last_row) {
            const float4 *src = reinterpret_cast<const float4 *>(p.xr + row * M);
            nx0 = src[tid]; nx1 = src[tid + NT];
        }
    };
    fetch_rows(t0);
    const int nlev = (lm - LFZ) / 2;                                // radix-4 levels (the first one in registers)
    for (int t = t0; t < t1; t += 2) {
        const int my_t = t + half_id;                               // this half's frame
        const bool have = my_t < t1;
        {
            float4 *dst = reinterpret_cast<float4 *>(ring + ((t + P - 1 + half_id) % RR) * M);
            dst[tid] = nx0; dst[tid + NT] = nx1;
        }
        __syncthreads();
        if (t + 2 < t1) fetch_rows(t + 2);
        if (have) {
            int slot = (my_t + P - 1) % RR;
            float2 acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; j++) acc[j] = cf(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < P; k++) {
                const float2 *row = ring + slot * M + tid;
#pragma unroll
                for (int j = 0; j < CPT; j++) ffma2(acc[j], hh[j][k], row[j * NT]);
                slot = (slot == 0) ? RR - 1 : slot - 1;
            }
            // first DIF level: inputs n = tid + j M/4, outputs to the same positions, twiddle W_M^(q tid)
            pfb_dif4(acc, stw[tid], stw[2 * tid], stw[3 * tid]);
#pragma unroll
            for (int j = 0; j < CPT; j++) { const int i = tid + j * NT; work[i + 2 * (i >> LFZ)] = acc[j]; }
        }
        __syncthreads();
        for (int lev = 1; lev < nlev; lev++) {
            // blocks of 4 q: butterfly b = tid -> block b / q, offset jj = b % q, twiddle W_{4q}^(jj) = W_M^(jj M / 4q)
            if (have) {
                const int lq = lm - 2 * (lev + 1), q = 1 << lq, jj = tid & (q - 1), base = ((tid >> lq) << (lq + 2)) + jj;
                float2 x[4];
#pragma unroll
                for (int k = 0; k < 4; k++) { const int i = base + k * q; x[k] = work[i + 2 * (i >> LFZ)]; }
                const int tws = jj << (2 * lev);
                pfb_dif4(x, stw[tws], stw[2 * tws], stw[3 * tws]);
#pragma unroll
                for (int k = 0; k < 4; k++) { const int i = base + k * q; work[i + 2 * (i >> LFZ)] = x[k]; }
            }
            __syncthreads();
        }
        if (have && tid < (M >> LFZ)) {
            // FZ-point DFT of block u = tid in registers (radix-2 DIT, natural-order output k); its frequencies are
            // c = digit-reverse(u) + (M / FZ) k
            const float4 *src = reinterpret_cast<const float4 *>(work + tid * (FZ + 2));
            float2 a[FZ], b[FZ];
#pragma unroll
            for (int i = 0; i < FZ / 2; i++) { const float4 v = src[i]; a[2 * i] = cf(v.x, v.y); a[2 * i + 1] = cf(v.z, v.w); }
#pragma unroll
            for (int i = 0; i < FZ; i++) b[i] = a[pfb_rev<LFZ>(i)];
#pragma unroll
            for (int s = 1; s <= LFZ; s++) {
                const int len = 1 << s, hl = len >> 1;
#pragma unroll
                for (int grp = 0; grp < FZ; grp += len) {
#pragma unroll
                    for (int j = 0; j < hl; j++) {
                        const float2 w = stw[(j * (FZ / len)) << (lm - LFZ)];
                        const float2 u = b[grp + j], v = pfb_cmul(b[grp + j + hl], w);
                        b[grp + j] = cf(u.x + v.x, u.y + v.y);
                        b[grp + j + hl] = cf(u.x - v.x, u.y - v.y);
                    }
                }
            }
            int cr = 0;
            for (int d = 0, u = tid; d < nlev; d++, u >>= 2) cr = (cr << 2) | (u & 3);   // base-4 digits of u reversed
            float2 *ob = obuf + ((my_t - t0) & (TF - 1)) * OR;
#pragma unroll
            for (int k = 0; k < FZ; k++) { const int c = cr + (k << (lm - LFZ)); ob[c + (c >> 4)] = b[k]; }
        }
        const int last = min(t + 1, t1 - 1);                        // last frame parked so far
        const int tfl = (last - t0) & (TF - 1);
        if (tfl == TF - 1 || last == t1 - 1) {
            __syncthreads();
            const int cnt = tfl + 1, tb = last - tfl;               // frames parked in the tile, first of them
            for (int e = gtid; e < M * TF; e += 2 * NT) {
                const int c = e >> 3, f = e & (TF - 1);
                if (f < cnt) {
                    const float2 v = obuf[f * OR + c + (c >> 4)];
                    p.y[(long long)c * p.y_stride + tb + f] = v;
                    if (p.pw) p.pw[(long long)c * p.pw_stride + tb + f] = pfb_power(v);
                }
            }
        }
        // the barrier at the top of the next iteration separates these reads from the next writes of work / obuf
    }
}

// ---- large power-of-two M (128..1024), one thread per polyphase branch --------------------------------------------
// Thread n of the CTA (M threads) owns branch n: its 14 taps and a sliding window of the branch's last samples live in
// REGISTERS, so the polyphase filter reads every input sample once from HBM (coalesced rows of M samples) and nothing
// from shared memory.  Four frames are filtered per iteration (17 window samples, static indices, one shift by four);
// their M-point DFTs run in shared memory as in-place radix-4 decimation-in-frequency passes with all M threads busy
// (4 frames x M/4 butterflies; an odd log2 M starts with one radix-2 pass), on a layout padded by one element per 16 so
// that the short-stride passes at the end stay (nearly) conflict-free.  The last pass writes straight into the
// transposed output tile (frequency of a position: `perm`, computed on the host by following the passes), which is
// flushed every 8 frames: 64 contiguous bytes per channel.  The loads of the next four frames are issued before the DFT
// passes and land while they run.
constexpr int kPfbStP = 14, kPfbStFI = 4, kPfbStTF = 8;
struct PfbStreamParams {
    const float2 *xr; float2 *y; long long y_stride;
    float *pw; long long pw_stride;        // optional: |y|^2, same layout
    int nf, T;                             // frames in this call, frames per CTA (multiple of kPfbStTF)
    int M, log2M;
    const float *h;                        // prototype, P*M taps
    const float2 *tw;                      // M twiddles exp(-j 2 pi t / M)
    const unsigned short *perm;            // [M] frequency held by position p behind the DIF passes
    int ocs, oco;                          // frame f is output column f * ocs + oco (firpfbch2: two passes, ocs = 2)
    int over2; float sc_even, sc_odd;      // firpfbch2: y[c] *= exp(-j 2 pi c / M) * (c even ? sc_even : sc_odd)
    int nf_odd; float sc_odd1;             // PAIR (clusters of two CTAs): frames and odd-channel factor of the odd pass
};
inline size_t pfb_stream_wp(int M) { return (size_t)M + ((size_t)M >> 4); }
inline size_t pfb_stream_orow(int M) { return (size_t)M + ((size_t)M >> 4) + 2; }
// twiddles: one contiguous table per pass (3 x N/4 entries W_N^(q j), q = 1..3; M/2 entries for the radix-2 pass), so that
// consecutive butterflies read consecutive entries (a shared table of W_M^t is read with strides 4, 16, 64, ...: 16-way
// bank conflicts in the short passes)
inline size_t pfb_stream_ntw(int M) { return 2 * (size_t)M; }
inline size_t pfb_stream_smem(int M)
{
    return sizeof(float2) * (kPfbStFI * pfb_stream_wp(M) + kPfbStTF * pfb_stream_orow(M) + pfb_stream_ntw(M)) + sizeof(unsigned short) * (size_t)M;
}
// frequency held by every position after the in-place DIF passes (radix-2 first when log2 M is odd, then radix-4)
inline void pfb_stream_perm(int M, unsigned short *perm)
{
    struct Job { int base, N, f0, s; };
    Job stack[64]; int sp = 0;
    stack[sp++] = Job{0, M, 0, 1};
    while (sp) {
        const Job j = stack[--sp];
        if (j.N == 1) { perm[j.base] = (unsigned short)(j.f0 & (M - 1)); continue; }
        int lg = 0; while ((1 << lg) < j.N) lg++;
        const int r = (lg & 1) ? 2 : 4;
        for (int q = 0; q < r; q++) stack[sp++] = Job{j.base + q * (j.N / r), j.N / r, j.f0 + j.s * q, j.s * r};
    }
}

// The radix-4 DIF passes over blocks of N (compile time) = M, M/4, ..., 16 with the thread's element and twiddle offsets computed
// ONCE (they depend on the thread, not on the frame; in the 16-point pass 16 consecutive butterflies take the same j of 16
// consecutive blocks: on the padded layout their elements are 17 apart -- 16 different banks -- and they share their twiddles): po[I] = padded offset of the butterfly's first element in pass I, to[I] = offset of its first twiddle.  The other
// three elements are at compile-time distances -- k Q + k Q / 16 on the padded layout (Q a multiple of 16; for Q = 4 the four
// elements share one group of 16) -- so the loads and stores of a pass are one register plus immediates.
template <int N, int M, int I = 0>
__device__ __forceinline__ void pfb_stream_pre(int bq, int tw0, int (&po)[5], int (&to)[5])
{
    if constexpr (N > 4) {
        constexpr int Q = N >> 2;
        int j = bq & (Q - 1), base = (bq / Q) * N + j;
        if constexpr (N == 16 && M >= 256) { j = (bq >> 4) & 3; base = (((bq & 15) | ((bq >> 6) << 4)) << 4) + j; }
        po[I] = base + (base >> 4);
        to[I] = tw0 + j;
        pfb_stream_pre<(N >> 2), M, I + 1>(bq, tw0 + 3 * Q, po, to);
    }
}
template <int N, int M, int I = 0>
__device__ __forceinline__ void pfb_stream_passes_pre(float2 *wf, const float2 *stw, const int (&po)[5], const int (&to)[5])
{
    if constexpr (N > 4) {
        constexpr int Q = N >> 2, D = Q >= 16 ? Q + Q / 16 : Q;      // padded distance of the butterfly's elements
        float2 *e = wf + po[I];
        const float2 *tp = stw + to[I];
        float2 x[4];
#pragma unroll
        for (int k = 0; k < 4; k++) x[k] = e[k * D];
        pfb_dif4(x, tp[0], tp[Q], tp[2 * Q]);
#pragma unroll
        for (int k = 0; k < 4; k++) e[k * D] = x[k];
        __syncthreads();
        pfb_stream_passes_pre<(N >> 2), M, I + 1>(wf, stw, po, to);
    }
}

// PAIR (firpfbch2, launched as clusters of two CTAs): the CTA of cluster rank 0 runs the even frames of a stretch, rank 1 the
// odd frames (the same filterbank on xr + M/2); every eight frames the two exchange their output tiles through distributed
// shared memory and each writes half of the channels with BOTH parities -- 16 neighbouring columns, whole 128-byte lines --
// instead of every other column of all channels (half sectors: 3.9 ms per pass against 2.2 ms for the same number of
// firpfbch frames).
template <int LM, bool PAIR = false>
__global__ void __launch_bounds__(1 << LM, 1) k_pfb_stream(const PfbStreamParams p)
{
    constexpr int P = kPfbStP, FI = kPfbStFI, TF = kPfbStTF;
    CSDR_DYN_SMEM(smem_raw);
    constexpr int M = 1 << LM, lm = LM;
    const int n = threadIdx.x;                                      // blockDim.x = M
    constexpr int WP = M + (M >> 4), OR = M + (M >> 4) + 2;
    float2 *work = reinterpret_cast<float2 *>(smem_raw);            // [FI][WP], element i at i + (i >> 4)
    float2 *obuf = work + FI * WP;                                  // [TF][OR], channel c at c + (c >> 4)
    float2 *stw = obuf + TF * OR;                                   // per-pass twiddle tables, < 2 M entries in all
    unsigned short *sperm = reinterpret_cast<unsigned short *>(stw + 2 * M);
    int par = 0, nf_own = p.nf;                                     // this CTA's parity (PAIR) and the frames of its pass
    const float2 *xin = p.xr;
#ifndef CSDR_EMU
    if constexpr (PAIR) {
        par = (int)cooperative_groups::this_cluster().block_rank();
        if (par) { xin += M / 2; nf_own = p.nf_odd; }
    }
#endif
    // t1c: end of the stretch in the pass with the most frames (the even one) -- both CTAs of a pair run the same iterations
    const int t0 = (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) * p.T, t1c = min(t0 + p.T, p.nf), t1 = min(t1c, nf_own);
    if (t0 >= t1c) return;
    sperm[n] = p.perm[n];
    {
        // radix-2 pass (odd log2 M): W_M^b, b < M/2; then for N = M (or M/2), N/4, ..., 16: W_N^(q j), q = 1..3, j < N/4
        int off = 0, N = M;
        if (lm & 1) { if (n < (M >> 1)) stw[n] = p.tw[n]; off = M >> 1; N = M >> 1; }
        for (; N > 4; N >>= 2) {
            const int Q = N >> 2;
            for (int e = n; e < 3 * Q; e += M) { const int q = e / Q + 1, j = e - (q - 1) * Q; stw[off + e] = p.tw[(q * j * (M / N)) & (M - 1)]; }
            off += 3 * Q;
        }
    }
    // taps of this thread's branch: hh[k] = h[(M-1-n) + k M]
    float hh[P];
#pragma unroll
    for (int k = 0; k < P; k++) hh[k] = p.h[(M - 1 - n) + k * M];
    // window: w[j] = sample of row (t + j) of xr in column n, t = first frame of the iteration (row r of xr = history or new
    // frame r - (P-1)); frame t + f needs rows t + f .. t + f + P - 1
    float2 w[P - 1 + FI];
    const long long last_row = (long long)nf_own + P - 2;           // last row that exists in xr
    const float2 *col = xin + n;
#pragma unroll
    for (int j = 0; j < P - 1 + FI; j++) {
        const long long row = (long long)t0 + j;
        w[j] = (row <= last_row) ? __ldg(col + row * M) : cf(0.f, 0.f);
    }
    constexpr int QB = M >> 2;                                      // butterflies per frame and radix-4 pass
    const int fi = n / QB, bq = n - fi * QB;                        // this thread's frame and butterfly in the radix-4 passes
    float2 *wf = work + fi * WP;
    int po[5], to[5];                                               // per-pass element / twiddle offsets of this thread
    if constexpr (lm & 1) pfb_stream_pre<(M >> 1), M>(bq, M >> 1, po, to); else pfb_stream_pre<M, M>(bq, 0, po, to);
    __syncthreads();
    for (int t = t0; t < t1c; t += FI) {
        // ---- polyphase filter: four frames from registers
#pragma unroll
        for (int f = 0; f < FI; f++) {
            float2 acc = cf(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < P; k++) ffma2(acc, hh[k], w[P - 1 + f - k]);
            work[f * WP + n + (n >> 4)] = acc;
        }
        // shift the window by four frames and fetch the next four rows (they land while the DFT passes run)
#pragma unroll
        for (int j = 0; j < P - 1; j++) w[j] = w[j + FI];
#pragma unroll
        for (int f = 0; f < FI; f++) {
            const long long row = (long long)t + FI + P - 1 + f;
            w[P - 1 + f] = (row <= last_row && t + FI < t1) ? __ldg(col + row * M) : cf(0.f, 0.f);
        }
        __syncthreads();
        // ---- DFT passes, in place (every block size a compile-time constant)
        if constexpr (lm & 1) {
            // radix-2: 4 frames x M/2 butterflies, two per thread
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const int e = n + r * M, f2 = e / (M >> 1), b = e - f2 * (M >> 1);
                float2 *wr = work + f2 * WP;
                const int i0 = b, i1 = b + (M >> 1);
                const float2 a = wr[i0 + (i0 >> 4)], c = wr[i1 + (i1 >> 4)];
                wr[i0 + (i0 >> 4)] = cf(a.x + c.x, a.y + c.y);
                wr[i1 + (i1 >> 4)] = pfb_cmul(cf(a.x - c.x, a.y - c.y), stw[b]);
            }
            __syncthreads();
            pfb_stream_passes_pre<(M >> 1), M>(wf, stw, po, to);
        } else {
            pfb_stream_passes_pre<M, M>(wf, stw, po, to);
        }
        {
            // last pass (N = 4, no twiddles): results go into the transposed output tile
            const int base = bq * 4;
            float2 x[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { const int i = base + k; x[k] = wf[i + (i >> 4)]; }
            {
                const float2 a0 = cf(x[0].x + x[2].x, x[0].y + x[2].y), a1 = cf(x[0].x - x[2].x, x[0].y - x[2].y);
                const float2 a2 = cf(x[1].x + x[3].x, x[1].y + x[3].y), a3 = cf(x[1].y - x[3].y, x[3].x - x[1].x);   // -j (x1 - x3)
                x[0] = cf(a0.x + a2.x, a0.y + a2.y); x[1] = cf(a1.x + a3.x, a1.y + a3.y);
                x[2] = cf(a0.x - a2.x, a0.y - a2.y); x[3] = cf(a1.x - a3.x, a1.y - a3.y);
            }
            const int my_t = t + fi;
            if (my_t < t1) {
                float2 *ob = obuf + ((my_t - t0) & (TF - 1)) * OR;
#pragma unroll
                for (int k = 0; k < 4; k++) { const int c = sperm[base + k]; ob[c + (c >> 4)] = x[k]; }
            }
        }
        const int last = min(t + FI - 1, t1c - 1);                  // last frame parked so far
        const int tfl = (last - t0) & (TF - 1);
        if (tfl == TF - 1 || last == t1c - 1) {
            const int cnt = tfl + 1, tb = last - tfl;               // frames parked in the tile, first of them
#ifndef CSDR_EMU
            if constexpr (PAIR) {
                auto cluster = cooperative_groups::this_cluster();
                cluster.sync();                                      // both tiles are parked
                const float2 *peer = cluster.map_shared_rank(obuf, (unsigned)(par ^ 1));
                const float2 *tile_of[2] = {par ? peer : obuf, par ? obuf : peer};       // [parity of the frame]
                for (int e = n; e < (M / 2) * 2 * TF; e += M) {
                    const int c = par * (M / 2) + (e >> 4), q = e & 15, fp = q & 1, f = q >> 1;
                    if (f < cnt && tb + f < (fp ? p.nf_odd : p.nf)) {
                        float2 v = tile_of[fp][f * OR + c + (c >> 4)];
                        const float2 w = __ldg(p.tw + c);
                        const float sc = (c & 1) ? (fp ? p.sc_odd1 : p.sc_odd) : p.sc_even;
                        v = cf((v.x * w.x - v.y * w.y) * sc, (v.x * w.y + v.y * w.x) * sc);
                        const long long col = (long long)(tb + f) * 2 + fp;
                        p.y[(long long)c * p.y_stride + col] = v;
                        if (p.pw) p.pw[(long long)c * p.pw_stride + col] = pfb_power(v);
                    }
                }
                cluster.sync();                                      // the peer has read this tile
            } else
#endif
            {
            __syncthreads();
            for (int e = n; e < M * TF; e += M) {
                const int c = e >> 3, f = e & (TF - 1);
                if (f < cnt && tb + f < t1) {
                    float2 v = obuf[f * OR + c + (c >> 4)];
                    if (p.over2) {
                        const float2 w = __ldg(p.tw + c);
                        const float sc = (c & 1) ? p.sc_odd : p.sc_even;
                        v = cf((v.x * w.x - v.y * w.y) * sc, (v.x * w.y + v.y * w.x) * sc);
                    }
                    const long long col = (long long)(tb + f) * p.ocs + p.oco;
                    p.y[(long long)c * p.y_stride + col] = v;
                    if (p.pw) p.pw[(long long)c * p.pw_stride + col] = pfb_power(v);
                }
            }
            }
        }
        __syncthreads();          // work / obuf are rewritten by the next iteration
    }
}

// keep the last (P-1)*M pre-rotated samples for the next call: dst[0..H) <- src[n .. n+H)
__global__ void k_copy_tail(const float2 *__restrict__ src, float2 *__restrict__ dst, long long offset, int count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[offset + i];
}

}  // namespace csdr
