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
    float2 tw[kPfbTileMaxM / 2];           // exp(-j 2 pi t / M)
    float h[kPfbTileP * kPfbTileMaxM];     // h[k * M + n] = prototype[(M - 1 - n) + k * M]
};

template <int LM> __host__ __device__ constexpr int pfb_rev(int i)
{
    int r = 0;
    for (int b = 0; b < LM; b++) r |= ((i >> b) & 1) << (LM - 1 - b);
    return r;
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
            ffma2(a[2 * j], p.h[k * M + 2 * j], cf(q.x, q.y));
            ffma2(a[2 * j + 1], p.h[k * M + 2 * j + 1], cf(q.z, q.w));
        }
    }
    // M-point forward DFT, radix-2 decimation in time
    float2 b[M];
#pragma unroll
    for (int i = 0; i < M; i++) b[i] = a[pfb_rev<LM>(i)];
#pragma unroll
    for (int s = 1; s <= LM; s++) {
        const int len = 1 << s, half = len >> 1, tws = M / len;
#pragma unroll
        for (int grp = 0; grp < M; grp += len) {
#pragma unroll
            for (int j = 0; j < half; j++) {
                const float2 w = p.tw[j * tws];
                const float2 u = b[grp + j], v = b[grp + j + half];
                const float tr = v.x * w.x - v.y * w.y, ti = v.x * w.y + v.y * w.x;
                b[grp + j] = cf(u.x + tr, u.y + ti);
                b[grp + j + half] = cf(u.x - tr, u.y - ti);
            }
        }
    }
    float2 *yo = p.y + t0 + f;
#pragma unroll
    for (int c = 0; c < M; c++) yo[(long long)c * p.y_stride] = b[c];
}

// ---- large power-of-two M (128..1024): a CTA slides over its frames with a ring of rows --------------------------
// Shared memory holds the last P + 1 = 15 rows of M samples (ring), two M-point DFT buffers and an [M][TF] output
// tile.  Two frames are processed per iteration by the two halves of the CTA (M/4 threads each): the two new rows are
// loaded once (coalesced, prefetched through registers one iteration ahead), thread i of a half evaluates the four
// polyphase branches n = i + j M/4 of its frame (4 x 14 taps in registers for the whole kernel) straight into
// digit-reversed DFT order, the DFT runs in shared memory (radix-4 DIT stages, one butterfly per thread and stage,
// plus one radix-2 stage when log2 M is odd), and the frame is parked in the output tile; every TF frames the tile is
// written out channel-major, TF consecutive frames (64 bytes) per channel.  Each input sample is read from HBM once
// (plus 13 rows of halo per CTA), each output written once.
constexpr int kPfbRingP = 14, kPfbRingTF = 8, kPfbRingCPT = 4, kPfbRingRows = kPfbRingP + 1;

struct PfbRingParams {
    const float2 *xr; float2 *y; long long y_stride;
    int nf, T;               // frames in this call, frames per CTA (multiple of TF)
    int M, log2M;
    const float *h;          // prototype, P*M taps
    const float2 *tw;        // M twiddles exp(-j 2 pi t / M)
};
inline size_t pfb_ring_smem(int M) { return (size_t)M * sizeof(float2) * (kPfbRingRows + 2 + (kPfbRingTF + 1) + 1); }

// position of element n in the input order of the mixed-radix DIT below (stage radices 2?, 4, 4, ...; the LAST stage
// splits n by its lowest base-4 digit): pos = digits of n in reverse order, i.e. the bit reversal of n with the two
// bits of every base-4 digit swapped back; with odd log2 M the top bit of n (the radix-2 stage) lands in bit 0
__device__ __forceinline__ int pfb_ring_perm(int n, int lm)
{
    const unsigned r = __brev((unsigned)n) >> (32 - lm);
    if (lm & 1) {
        const unsigned low = r & 1u, up = r >> 1;
        return (int)(((((up & 0x55555555u) << 1) | ((up & 0xAAAAAAAAu) >> 1)) << 1) | low);
    }
    return (int)(((r & 0x55555555u) << 1) | ((r & 0xAAAAAAAAu) >> 1));
}

__global__ void __launch_bounds__(512, 1) k_pfb_ring(const PfbRingParams p)
{
    constexpr int P = kPfbRingP, TF = kPfbRingTF, CPT = kPfbRingCPT, TFP = TF + 1, RR = kPfbRingRows;
    CSDR_DYN_SMEM(smem_raw);
    const int M = p.M, lm = p.log2M, NT = M / CPT;                // blockDim.x = 2 NT
    const int half_id = threadIdx.x / NT, tid = threadIdx.x - half_id * NT, gtid = threadIdx.x;
    float2 *ring = reinterpret_cast<float2 *>(smem_raw);          // [RR][M]
    float2 *work = ring + RR * M + half_id * M;                    // [2][M]
    float2 *obuf = ring + RR * M + 2 * M;                          // [M][TFP]
    float2 *stw = obuf + M * TFP;                                  // [M] twiddles
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
#pragma unroll
            for (int j = 0; j < CPT; j++) work[pfb_ring_perm(tid + j * NT, lm)] = acc[j];
        }
        __syncthreads();
        int s = 0;                                                  // bits done so far
        if (lm & 1) {
            // one radix-2 stage on adjacent elements: two butterflies per thread
            if (have) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    float2 *a = work + 2 * (tid + q * NT);
                    const float2 u = a[0], v = a[1];
                    a[0] = cf(u.x + v.x, u.y + v.y);
                    a[1] = cf(u.x - v.x, u.y - v.y);
                }
            }
            __syncthreads();
            s = 1;
        }
        for (; s < lm; s += 2) {
            // radix-4 DIT stage: sub-transforms of length 2^s -> 2^(s+2); one butterfly per thread
            if (have) {
                const int quarter = 1 << s, jj = tid & (quarter - 1), grp = tid >> s;
                float2 *a = work + (grp << (s + 2)) + jj;
                const int tws = jj << (lm - s - 2);                 // W_{4 quarter}^{jj} = W_M^{jj M / (4 quarter)}
                const float2 w1 = stw[tws], w2 = stw[2 * tws], w3 = stw[3 * tws];
                const float2 x0 = a[0], x1 = a[quarter], x2 = a[2 * quarter], x3 = a[3 * quarter];
                const float2 b1 = cf(x1.x * w1.x - x1.y * w1.y, x1.x * w1.y + x1.y * w1.x);
                const float2 b2 = cf(x2.x * w2.x - x2.y * w2.y, x2.x * w2.y + x2.y * w2.x);
                const float2 b3 = cf(x3.x * w3.x - x3.y * w3.y, x3.x * w3.y + x3.y * w3.x);
                const float2 s02 = cf(x0.x + b2.x, x0.y + b2.y), d02 = cf(x0.x - b2.x, x0.y - b2.y);
                const float2 s13 = cf(b1.x + b3.x, b1.y + b3.y), d13 = cf(b1.x - b3.x, b1.y - b3.y);
                a[0] = cf(s02.x + s13.x, s02.y + s13.y);
                a[quarter] = cf(d02.x + d13.y, d02.y - d13.x);      // d02 - j d13
                a[2 * quarter] = cf(s02.x - s13.x, s02.y - s13.y);
                a[3 * quarter] = cf(d02.x - d13.y, d02.y + d13.x);  // d02 + j d13
            }
            __syncthreads();
        }
        if (have) {
            const int tf = (my_t - t0) & (TF - 1);
#pragma unroll
            for (int j = 0; j < CPT; j++) { const int c = tid + j * NT; obuf[c * TFP + tf] = work[c]; }
        }
        const int last = min(t + 1, t1 - 1);                        // last frame parked so far
        const int tfl = (last - t0) & (TF - 1);
        if (tfl == TF - 1 || last == t1 - 1) {
            __syncthreads();
            const int cnt = tfl + 1, tb = last - tfl;               // frames parked in the tile, first of them
            for (int e = gtid; e < M * TF; e += 2 * NT) {
                const int c = e >> 3, f = e & (TF - 1);
                if (f < cnt) p.y[(long long)c * p.y_stride + tb + f] = obuf[c * TFP + f];
            }
        }
    }
}

// keep the last (P-1)*M pre-rotated samples for the next call: dst[0..H) <- src[n .. n+H)
__global__ void k_copy_tail(const float2 *__restrict__ src, float2 *__restrict__ dst, long long offset, int count)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[offset + i];
}

}  // namespace csdr
