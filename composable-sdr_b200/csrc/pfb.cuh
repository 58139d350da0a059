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
#include "platform.cuh"

namespace csdr {

struct PfbParams {
    const float2 *xr;       // pre-rotated samples: xr[0 .. (P-1)*M) = history frames, then nf*M new samples
    float2 *y;              // [M][y_stride] channel-major output, frame t at column t
    long long y_stride;
    int M, P, nf, F;        // channels, taps per branch, frames in this call, frames per CTA
    int log2M;              // >= 0 when M is a power of two, else -1
    const float *h;         // prototype, P*M taps
    const float2 *tw;       // M twiddles exp(-j 2 pi t / M)
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
        const float2 *src = p.xr + (long long)(t0 + f + P - 1) * M + n;     // newest sample of this branch
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
        p.y[(long long)c * p.y_stride + t0 + f] = buf[f * M + c];
    }
}

// keep the last (P-1)*M pre-rotated samples for the next call: dst[0..H) <- src[n .. n+H)
__global__ void k_copy_tail(const float2 *__restrict__ src, float2 *__restrict__ dst, long long offset, int count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[offset + i];
}

}  // namespace csdr
