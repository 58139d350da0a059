// frontend_std.cuh -- the fused front end (see frontend.cuh) specialised at compile time for the half-band plan
// msresamp_crcf_create(r, 60 dB) always produces (m = 10, 5, 3, 3, ...; the reference hard-codes As = 60,
// apps/SoapySDR.hs:194).  The whole tile geometry is constexpr (frontend_geom.hpp), so every shared-memory access is
// base register + immediate, loop trip counts are constants and the half-band taps are read straight from the
// kernel-parameter constant bank by the FFMAs.  Other plans use the generic k_frontend.
//
// Three kernels share the stage, bookkeeping and resampler code (CSDR_OPT_FRONTEND_VARIANT):
//   k_frontend_direct<S> (1, default): the raw tile arrives one tile ahead by a TMA tensor copy (cp.async.bulk.tensor,
//       128-byte swizzle, mbarrier) and the first half-band stage reads it where it lands, mixing in registers.
//       3 CTAs per SM.
//   k_frontend_std<S> (0): raw samples are prefetched into registers one tile ahead, mixed and written de-interleaved
//       by a separate loader pass.  2 CTAs per SM.  Kept as the cross-check of the direct kernel.
//   k_frontend_ws<S> (2): the direct kernel split into producer / consumer warp groups (measured slower; experiment).
#pragma once
#include "frontend.cuh"

namespace csdr {

template <int S, int V = 0> struct FeStd { static constexpr FeGeom G = fe_make_geom_std(S, V); };

__device__ __forceinline__ float4 fe_ldg_stream(const float4 *p)
{
#ifdef CSDR_EMU
    return *p;
#else
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}

// v * conj(phasor) (down) or v * phasor (up)
template <int MIX>
__device__ __forceinline__ float2 fe_mix(float2 v, unsigned th, int quantize)
{
    if ((MIX & 3) == 0) return v;
    const float2 w = fe_phasor(th, (MIX & 4) ? 0 : ((MIX & 8) ? 1 : quantize));     // bits 2/3: flag known at compile time
    constexpr int MODE = MIX & 3;
    const float s = (MODE == 1) ? -w.y : w.y;
    return cf(v.x * w.x - v.y * s, v.y * w.x + v.x * s);
}

// MIX: bits 0-1 = mode (0 none, 1 down: v * conj(w), 2 up: v * w), bit 3 = quantised NCO
template <int MIX>
__device__ __forceinline__ float2 fe_mix_q(float2 v, unsigned th)
{
    if constexpr ((MIX & 3) == 0) return v;
    else {
        const float2 w = fe_phasor_q<(MIX >> 3) & 1>(th);
        const float s = ((MIX & 3) == 1) ? -w.y : w.y;
        return cf(v.x * w.x - v.y * s, v.y * w.x + v.x * s);
    }
}

// ---- top level: (mix) -> shared, in the consumer's de-interleaved layout ------------------------------------
// A tile that lies inside the chunk at a 16-byte aligned address is "bulk": its raw samples are prefetched into
// registers (fe_prefetch) while the previous tile is being filtered.
template <int S, int V>
__device__ __forceinline__ bool fe_tile_is_bulk(const FrontendParams &p, const float2 *xs, long long lo)
{
    constexpr int NS = FeStd<S, V>::G.n[S];
    const long long rel0 = lo - p.n0;
    if (V >= 1) {
        // the tensor map covers whole 16-sample rows of x from sample tma_r on; tiles start on row boundaries
        const long long rr = rel0 - p.tma_r;
        return p.tma_ok && rr >= 0 && (rr & (kFeRawRow - 1)) == 0 && (rr / kFeRawRow) + NS / kFeRawRow <= p.tma_rows;
    }
    return rel0 >= 0 && rel0 + NS <= p.nx && (reinterpret_cast<uintptr_t>(xs + rel0) & 15) == 0;
}

template <int S> struct FePrefetch {
    static constexpr int NP = FeStd<S, 0>::G.n[S] / 2, IT = (NP + kFeNT - 1) / kFeNT;
    float4 v[IT];
};

// issue the 16-byte loads of a bulk tile; they complete while the previous tile is being filtered
template <int S>
__device__ __forceinline__ void fe_prefetch(FePrefetch<S> &pre, const FrontendParams &p, const float2 *__restrict__ xs,
                                            long long lo)
{
    const float4 *src = reinterpret_cast<const float4 *>(xs + (lo - p.n0));
#pragma unroll
    for (int k = 0; k < FePrefetch<S>::IT; k++) {
        const int pi = threadIdx.x + kFeNT * k;
        if (pi < FePrefetch<S>::NP) pre.v[k] = fe_ldg_stream(src + pi);
    }
}

template <int S, int MIX>
__device__ __forceinline__ void fe_load_top(const FrontendParams &p, const float2 *__restrict__ xs,
                                            const float2 *__restrict__ hs, float2 *__restrict__ dst, long long lo,
                                            const FePrefetch<S> &pre, bool bulk)
{
    constexpr FeGeom G = FeStd<S, 0>::G;
    constexpr int NS = G.n[S], STR = G.stride[S], D = G.R[S - 1];
    static_assert(D == 8 && NS % 2 == 0, "loader assumes an 8-way layout of the top level");
    const int tid = threadIdx.x;
    const long long rel0 = lo - p.n0;
    const unsigned th0 = p.theta0 + (unsigned)lo * p.dtheta;
    const bool inside = rel0 >= 0 && rel0 + NS <= p.nx;
    if (bulk) {
        // one 16-byte load = the (even, odd) pair p; pair -> sub-array p & 7, index p >> 3
        float2 *dE = dst + (tid & 7) * STR + (tid >> 3);
        float2 *dO = dE + D * STR + kFePlanePad;
#pragma unroll
        for (int k = 0; k < FePrefetch<S>::IT; k++) {
            const int pi = tid + kFeNT * k;
            if (pi < FePrefetch<S>::NP) {
                const float4 v = pre.v[k];
                const unsigned th = th0 + (unsigned)(2 * pi) * p.dtheta;
                dE[(kFeNT / 8) * k] = fe_mix<MIX>(cf(v.x, v.y), th, p.quantize);
                dO[(kFeNT / 8) * k] = fe_mix<MIX>(cf(v.z, v.w), th + p.dtheta, p.quantize);
            }
        }
    } else if (inside) {
        // chunk not 16-byte aligned at this tile: 8-byte global loads, sample i -> plane i & 1, pair i >> 1
        const float2 *src = xs + rel0;
        float2 *d0 = dst + (tid & 1) * (D * STR + kFePlanePad) + ((tid >> 1) & 7) * STR + (tid >> 4);
        constexpr int IT = (NS + kFeNT - 1) / kFeNT;
#pragma unroll 4
        for (int k = 0; k < IT; k++) {
            const int i = tid + kFeNT * k;
            if (i < NS) d0[(kFeNT / 16) * k] = fe_mix<MIX>(src[i], th0 + (unsigned)i * p.dtheta, p.quantize);
        }
    } else {
        // edge tile: samples before the chunk come from the carried history, samples after it are zero
        for (int i = tid; i < NS; i += kFeNT) {
            const long long rel = rel0 + i;
            float2 v = cf(0.f, 0.f);
            if (rel >= 0) { if (rel < p.nx) v = xs[rel]; }
            else if (rel >= -(long long)p.hcap) v = hs[p.hcap + rel];
            dst[fe_addr<D>(i, STR)] = fe_mix<MIX>(v, th0 + (unsigned)i * p.dtheta, p.quantize);
        }
    }
}

// ---- one half-band stage with compile-time geometry --------------------------------------------------------
//   out[q] = C[q+M] + sum_{u<2M} g[u] * T[q+u+SH]     T/C = tap/centre planes (odd/even samples; swapped when the
//   input level is shifted by one sample), R consecutive outputs per thread slot
template <int M, int R, int STR, int SH, int NOUT, bool LAST, int D2, int STR2, int NTG = kFeNT>
__device__ __forceinline__ void fe_stage_c(const float2 *__restrict__ in, float2 *__restrict__ out,
                                           const float *__restrict__ g, float zeta, int tid = threadIdx.x)
{
    const float2 *T = in + (SH ? 0 : R * STR + kFePlanePad);
    const float2 *C = in + (SH ? R * STR + kFePlanePad : 0);
    constexpr int NSLOTS = NOUT / R;
    for (int t = tid; t < NSLOTS; t += NTG) {
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = C[((M + r) % R) * STR + t + (M + r) / R];
#pragma unroll
        for (int c = 0; c < R + 2 * M - 1; c++) {
            const float2 v = T[((c + SH) % R) * STR + t + (c + SH) / R];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int u = c - r;
                if (u >= 0 && u < 2 * M) ffma2(acc[r], g[u], v);
            }
        }
        if constexpr (LAST) {
            float2 *o = out + t * R;
#pragma unroll
            for (int r = 0; r < R; r++) o[r] = cf(acc[r].x * zeta, acc[r].y * zeta);
        } else {
            // q = R t + r -> plane r & 1, pair p = H t + (r >> 1) (H = R/2) -> sub-array p & (D2-1), index p / D2
            constexpr int H = R / 2;
            static_assert(H >= 1 && (D2 & (D2 - 1)) == 0, "layout factors are powers of two");
            if constexpr (H >= D2) {
                // sub-array = (r >> 1) & (D2-1) for every t, index = (H / D2) t + (r >> 1) / D2
                float2 *o = out + t * (H / D2);
#pragma unroll
                for (int r = 0; r < R; r++) o[(r & 1) * (D2 * STR2 + kFePlanePad) + ((r >> 1) & (D2 - 1)) * STR2 + (r >> 1) / D2] = acc[r];
            } else {
                // K = D2 / H slots share one index: sub-array = H (t & (K-1)) + (r >> 1), index = t / K
                constexpr int K = D2 / H;
                float2 *o = out + (H * (t & (K - 1))) * STR2 + t / K;
#pragma unroll
                for (int r = 0; r < R; r++) o[(r & 1) * (D2 * STR2 + kFePlanePad) + (r >> 1) * STR2] = acc[r];
            }
        }
    }
}

template <int S, int V, int s>
__device__ __forceinline__ void fe_run_stages(const FrontendParams &p, float2 *smem)
{
    constexpr FeGeom G = FeStd<S, V>::G;
    constexpr bool LAST = (s == 0);
    constexpr int D2 = LAST ? 1 : G.R[LAST ? 0 : s - 1];
    constexpr int SH = (s == S - 1) ? G.shift : 0;
    fe_stage_c<G.m[s], G.R[s], G.stride[s + 1], SH, G.n[s], LAST, D2, G.stride[s]>(
        smem + G.off[s + 1], smem + G.off[s], p.taps[s], p.zeta);
    __syncthreads();
    if constexpr (s > 0) fe_run_stages<S, V, s - 1>(p, smem);
}

// ceil(num / st) for num < 2^57, st < 2^26 without a 64-bit division: fp64 estimate + exact integer correction
__device__ __forceinline__ long long fe_ceil_div(unsigned long long num, unsigned st, double inv_st)
{
    long long q = (long long)((double)num * inv_st);
    long long r = (long long)num - q * (long long)st;
    while (r < 0) { q--; r += st; }
    while (r >= (long long)st) { q++; r -= st; }
    return q + (r != 0);
}

struct FeTileInfo {
    long long lo; int bulk, oA, oB;
    unsigned phA;     // timing phase of output oA relative to the tile's first push, in [0, step)
    float fA;         // phA / 2^24
    int pad;
};

// everything one tile needs that is not per-thread work: computed by ONE thread, one tile ahead
template <int S, int V>
__device__ __forceinline__ void fe_tile_info(const FrontendParams &p, const float2 *xs, int tile, double inv_st, FeTileInfo &ti)
{
    constexpr FeGeom G = FeStd<S, V>::G;
    const long long kArel = (long long)tile * G.Tc;                 // pushes relative to K0
    const long long kBrel = min(kArel + (long long)G.Tc, p.K1 - p.K0);
    ti.lo = (p.K0 + kArel - kHcPad) * (1LL << S) + G.d[S];
    ti.bulk = fe_tile_is_bulk<S, V>(p, xs, ti.lo) ? 1 : 0;
    // outputs emitted by pushes [kA, kB): o' with kArel*2^24 <= ph0 + o'*step < kBrel*2^24
    const unsigned long long a = (unsigned long long)kArel << 24, b = (unsigned long long)kBrel << 24;
    ti.oA = (a > p.ph0) ? (int)fe_ceil_div(a - p.ph0, p.step, inv_st) : 0;
    ti.oB = (b > p.ph0) ? (int)fe_ceil_div(b - p.ph0, p.step, inv_st) : 0;
    ti.phA = (unsigned)(p.ph0 + (unsigned long long)ti.oA * p.step - a);
    ti.fA = (float)ti.phA * (1.0f / 16777216.0f);
}

// ---- arbitrary resampler over one tile (rate_arb < 1: every push emits at most one output) --------------------
// One thread per PAIR of pushes: the 16 c-samples both windows need are fetched with eight conflict-free 16-byte
// loads and stay in registers.  Output oA + j has timing phase phA + j*step relative to the tile, belongs to local
// push (phase >> 24) and uses branch = the next `bits` phase bits.  The first j of a thread is estimated in fp32 and
// corrected exactly in integers (no division).  Taps come from the padded copy of the bank in shared memory.
template <int TC, int NTG = kFeNT>
__device__ __forceinline__ void fe_resample_tile(const FrontendParams &p, const float2 *__restrict__ cb, float2 *__restrict__ ys,
                                                 const FeTileInfo &ti, int npush, const float *__restrict__ bank_s, float rate_f,
                                                 int tid = threadIdx.x)
{
    const unsigned mask = (1u << p.bits) - 1u;
    const int sh = 24 - p.bits;
    const unsigned phA = ti.phA;
    const float fA = ti.fA;
    float2 *yo = ys + ti.oA;
    for (int t = tid; t < TC / 2; t += NTG) {
        if (2 * t >= npush) break;
        float2 w[16];                               // local indices 2t+2 .. 2t+17; push e ends at w[14 + e]
        const float4 *src = reinterpret_cast<const float4 *>(cb + 2 * t + 2);
#pragma unroll
        for (int i = 0; i < 8; i++) { const float4 q = src[i]; w[2 * i] = cf(q.x, q.y); w[2 * i + 1] = cf(q.z, q.w); }
        // first output at or after local push 2t
        const unsigned long long target = (unsigned long long)(2 * t) << 24;
        const float jf = ((float)(2 * t) - fA) * rate_f;
        int j = jf > 0.f ? (int)jf : 0;
        unsigned long long P = phA + (unsigned long long)(unsigned)j * p.step;
        while (P < target) { j++; P += p.step; }
        while (j > 0 && P - p.step >= target) { j--; P -= p.step; }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            if (2 * t + e < npush && (int)(P >> 24) == 2 * t + e) {
                const unsigned br = ((unsigned)P >> sh) & mask;
                float h[kHsub];
                const float *hr = bank_s + br * (kHsub + 1);
#pragma unroll
                for (int i = 0; i < kHsub; i++) h[i] = hr[i];
                float ar = 0.f, ai = 0.f;
#pragma unroll
                for (int jj = 0; jj < kHsub; jj++) {
                    ar = fmaf(h[jj], w[14 + e - jj].x, ar);
                    ai = fmaf(h[jj], w[14 + e - jj].y, ai);
                }
                yo[j] = cf(ar, ai);
                j++; P += p.step;
            }
        }
    }
}

template <int S>
__global__ void __launch_bounds__(kFeNT, 2) k_frontend_std(const CSDR_GRID_CONSTANT FrontendParams p)
{
    constexpr FeGeom G = FeStd<S, 0>::G;
    CSDR_DYN_SMEM(smem_raw);
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    float *bank_s = reinterpret_cast<float *>(smem_raw) + 2 * G.total_f2;
    __shared__ FeTileInfo s_info[3];

    const int npfb = 1 << p.bits;
    for (int i = threadIdx.x; i < npfb * kHsub; i += kFeNT) {
        const int row = i / kHsub, col = i - row * kHsub;
        bank_s[row * (kHsub + 1) + col] = p.bank[i];
    }
    const float2 *xs = p.x + (long long)blockIdx.y * p.x_stride;
    const float2 *hs = p.hist + (long long)blockIdx.y * p.hcap;
    float2 *ys = p.y + (long long)blockIdx.y * p.y_stride;
    const double inv_st = 1.0 / (double)p.step;
    const float rate_f = 16777216.0f / (float)p.step;
    const int gstep = (int)gridDim.x;

    // Software pipeline over the tiles of this (persistent) CTA: the raw samples of tile i+1 are loaded into
    // registers while tile i is filtered, so the sequential load -> mix -> filter chain never waits for HBM.
    // Tile bookkeeping (absolute position, alignment, output range) is done by one thread, two tiles ahead.
    if (threadIdx.x == 0) {
        const int t0 = (int)blockIdx.x;
        if (t0 < p.ntiles) fe_tile_info<S, 0>(p, xs, t0, inv_st, s_info[0]);
        if (t0 + gstep < p.ntiles) fe_tile_info<S, 0>(p, xs, t0 + gstep, inv_st, s_info[1]);
    }
    __syncthreads();
    FePrefetch<S> pre;
    if ((int)blockIdx.x < p.ntiles && s_info[0].bulk) fe_prefetch<S>(pre, p, xs, s_info[0].lo);

    int cur = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gstep) {
        const int nxt = cur == 2 ? 0 : cur + 1, nxt2 = nxt == 2 ? 0 : nxt + 1;
        const long long lo = s_info[cur].lo;
        const bool bulk = s_info[cur].bulk != 0;
        float2 *top = smem + G.off[S];
        if (p.mix_mode == 0)      fe_load_top<S, 0>(p, xs, hs, top, lo, pre, bulk);
        else if (p.quantize) { if (p.mix_mode == 1) fe_load_top<S, 1 | 8>(p, xs, hs, top, lo, pre, bulk);
                               else                 fe_load_top<S, 2 | 8>(p, xs, hs, top, lo, pre, bulk); }
        else                 { if (p.mix_mode == 1) fe_load_top<S, 1 | 4>(p, xs, hs, top, lo, pre, bulk);
                               else                 fe_load_top<S, 2 | 4>(p, xs, hs, top, lo, pre, bulk); }
        if (threadIdx.x == 0 && tile + 2 * gstep < p.ntiles) fe_tile_info<S, 0>(p, xs, tile + 2 * gstep, inv_st, s_info[nxt2]);
        __syncthreads();
        if (tile + gstep < p.ntiles && s_info[nxt].bulk) fe_prefetch<S>(pre, p, xs, s_info[nxt].lo);

        fe_run_stages<S, 0, S - 1>(p, smem);
        {
            const int npush = (int)min((long long)G.Tc, p.K1 - p.K0 - (long long)tile * G.Tc);
            fe_resample_tile<G.Tc>(p, smem + G.off[0], ys, s_info[cur], npush, bank_s, rate_f);
        }
        __syncthreads();   // smem is reused by the next tile
        cur = nxt;
    }
}

// =============================================================================================================
// k_frontend_direct: the raw tile is brought in by ONE TMA tensor copy (rows of 16 samples, 128-byte swizzle) and the
// first half-band stage reads it where it lands: one 16-byte load fetches the (even, odd) pair -- the even sample is a
// tap input, the odd one a centre input (the top level starts one sample early, shift = 1) -- and both are multiplied
// by the NCO phasor in registers.  8 outputs per thread slot = one row per slot: the swizzle puts the same chunk
// position of eight consecutive rows into eight different banks, so the loads of a quarter-warp are conflict-free, and
// the results go to the next level with the conflict-free store pattern of fe_stage_c.  No mixing pass, no
// de-interleaving pass, no LSU work for the copy; the lower stages and the resampler are those of k_frontend_std.

// 16-byte chunk index of pair P in the swizzled tile
__device__ __forceinline__ int fe_swz(int P) { return (P & ~7) | ((P ^ (P >> 3)) & 7); }

template <int M, int NOUT, bool LAST, int D2, int STR2, int MIX, int NTG = kFeNT>
__device__ __forceinline__ void fe_stage_top(const float2 *__restrict__ raw, float2 *__restrict__ out,
                                             const float *__restrict__ g, float zeta, unsigned thb, unsigned dth, int tid = threadIdx.x)
{
    constexpr int R = kFeTopR, NSLOTS = NOUT / R, NC = R + 2 * M - 1;
    static_assert(R == 8 && NOUT % R == 0 && kFeRawRow == 2 * R, "one slot per 128-byte row");
    const float4 *src = reinterpret_cast<const float4 *>(raw);
    for (int t = tid; t < NSLOTS; t += NTG) {
        // output q = R t + r:  centre = odd sample of pair q + M, tap u = even sample of pair q + u + 1.
        // Pair 8 t + k sits in row t + (k >> 3), chunk (k & 7) ^ ((t + (k >> 3)) & 7).
        const unsigned ths = thb + (unsigned)(2 * (R * t + 1)) * dth;
        float2 acc[R];
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = cf(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const int k = c + 1, row = t + (k >> 3);
            const float4 pr = src[row * 8 + ((k & 7) ^ (row & 7))];
            const float2 v = fe_mix_q<MIX>(cf(pr.x, pr.y), ths + (unsigned)(2 * c) * dth);
#pragma unroll
            for (int r = 0; r < R; r++) {
                const int u = c - r;
                if (u >= 0 && u < 2 * M) ffma2(acc[r], g[u], v);
            }
            if (c >= M - 1 && c < M - 1 + R) {
                const float2 e = fe_mix_q<MIX>(cf(pr.z, pr.w), ths + (unsigned)(2 * c + 1) * dth);
                acc[c - (M - 1)].x += e.x; acc[c - (M - 1)].y += e.y;
            }
        }
        if constexpr (LAST) {
            float2 *o = out + t * R;
#pragma unroll
            for (int r = 0; r < R; r++) o[r] = cf(acc[r].x * zeta, acc[r].y * zeta);
        } else {
            // q = R t + r -> plane r & 1, pair p = H t + (r >> 1) (H = R/2) -> sub-array p & (D2-1), index p / D2
            constexpr int H = R / 2;
            if constexpr (H >= D2) {
                float2 *o = out + t * (H / D2);
#pragma unroll
                for (int r = 0; r < R; r++) o[(r & 1) * (D2 * STR2 + kFePlanePad) + ((r >> 1) & (D2 - 1)) * STR2 + (r >> 1) / D2] = acc[r];
            } else {
                constexpr int K = D2 / H;
                float2 *o = out + (H * (t & (K - 1))) * STR2 + t / K;
#pragma unroll
                for (int r = 0; r < R; r++) o[(r & 1) * (D2 * STR2 + kFePlanePad) + (r >> 1) * STR2] = acc[r];
            }
        }
    }
}

template <int S, int MIX>
__device__ __forceinline__ void fe_run_top(const FrontendParams &p, float2 *smem, unsigned thb)
{
    constexpr FeGeom G = FeStd<S, 1>::G;
    constexpr bool LAST = (S == 1);
    constexpr int D2 = LAST ? 1 : G.R[LAST ? 0 : S - 2];
    static_assert(G.shift == 1, "pairs are (tap, centre) only when the top level starts one sample early");
    fe_stage_top<G.m[S - 1], G.n[S - 1], LAST, D2, G.stride[S - 1], MIX>(smem + G.off[S], smem + G.off[S - 1],
                                                                          p.taps[S - 1], p.zeta, thb, p.dtheta);
}
// staging[i] = raw sample lo + i for a tile the bulk copy cannot fetch (it reaches into the carried history, past
// the end of the chunk, or the chunk is not 16-byte aligned there)
template <int S, int NTG = kFeNT, int V = 1>
__device__ __forceinline__ void fe_fill_staging(const FrontendParams &p, const float2 *__restrict__ xs,
                                                const float2 *__restrict__ hs, float2 *__restrict__ raw, long long lo,
                                                int tid = threadIdx.x)
{
    constexpr int NS = FeStd<S, V>::G.n[S];
    const long long rel0 = lo - p.n0;
    for (int i = tid; i < NS; i += NTG) {
        const long long rel = rel0 + i;
        float2 v = cf(0.f, 0.f);
        if (rel >= 0) { if (rel < p.nx) v = xs[rel]; }
        else if (rel >= -(long long)p.hcap) v = hs[p.hcap + rel];
        raw[2 * fe_swz(i >> 1) + (i & 1)] = v;
    }
}

// one tensor copy per tile (two when the tile has more than 256 rows), issued by one thread
template <int S, int V = 1>
__device__ __forceinline__ void fe_copy_staging(const FrontendParams &p, const FeTmap *tm, float2 *raw, long long lo,
                                                unsigned long long *bar)
{
    constexpr int ROWS = FeStd<S, V>::G.n[S] / kFeRawRow, NB = (ROWS + 255) / 256, BOX = ROWS / NB;
    static_assert(BOX * NB == ROWS && BOX <= 256, "tile = NB boxes of BOX rows");
    const int row0 = (int)((lo - p.n0 - p.tma_r) / kFeRawRow);
    if (NB == 1) tma_load_rows(raw, tm, row0, (int)blockIdx.y, ROWS, bar);
    else {
        // a barrier phase takes ONE arrival: announce the whole tile with the first box
        tma_load_rows(raw, tm, row0, (int)blockIdx.y, BOX, bar, ROWS);
        for (int b = 1; b < NB; b++) tma_load_rows(raw + b * BOX * kFeRawRow, tm, row0 + b * BOX, (int)blockIdx.y, BOX, bar, 0);
    }
}

#if defined(CSDR_FE_SKIP) && (CSDR_FE_SKIP & 16)
#define FE_NOFILL 1        // timing experiments only: leave the staging buffer as it is
#else
#define FE_NOFILL 0
#endif
#ifndef CSDR_FE_MINB
#define CSDR_FE_MINB 2
#endif
template <int S>
__global__ void __launch_bounds__(kFeNT, CSDR_FE_MINB) k_frontend_direct(const CSDR_GRID_CONSTANT FrontendParams p, const CSDR_GRID_CONSTANT FeTmap tmap)
{
    constexpr FeGeom G = FeStd<S, 1>::G;
    constexpr int NS = G.n[S];
    CSDR_DYN_SMEM_1K(smem_raw1k);
    unsigned char *smem_raw = smem_raw1k;
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    float *bank_s = reinterpret_cast<float *>(smem_raw) + 2 * G.total_f2;
    float2 *raw = smem + G.off[S];
    __shared__ FeTileInfo s_info[3];
    __shared__ __align__(8) unsigned long long s_bar;

    const int npfb = 1 << p.bits;
    for (int i = threadIdx.x; i < npfb * kHsub; i += kFeNT) {
        const int row = i / kHsub, col = i - row * kHsub;
        bank_s[row * (kHsub + 1) + col] = p.bank[i];
    }
    const float2 *xs = p.x + (long long)blockIdx.y * p.x_stride;
    const float2 *hs = p.hist + (long long)blockIdx.y * p.hcap;
    float2 *ys = p.y + (long long)blockIdx.y * p.y_stride;
    const double inv_st = 1.0 / (double)p.step;
    const float rate_f = 16777216.0f / (float)p.step;
    const int gstep = (int)gridDim.x;

    unsigned parity = 0;
    if (threadIdx.x == 0) {
        const int t0 = (int)blockIdx.x;
        bulk_init(&s_bar);
        if (t0 < p.ntiles) fe_tile_info<S, 1>(p, xs, t0, inv_st, s_info[0]);
        if (t0 + gstep < p.ntiles) fe_tile_info<S, 1>(p, xs, t0 + gstep, inv_st, s_info[1]);
    }
    __syncthreads();
    if ((int)blockIdx.x < p.ntiles) {
        if (FE_NOFILL) {} else if (s_info[0].bulk) { if (threadIdx.x == 0) fe_copy_staging<S>(p, &tmap, raw, s_info[0].lo, &s_bar); }
        else fe_fill_staging<S>(p, xs, hs, raw, s_info[0].lo);
    }
    __syncthreads();

    int cur = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gstep) {
        const int nxt = cur == 2 ? 0 : cur + 1, nxt2 = nxt == 2 ? 0 : nxt + 1;
        if (s_info[cur].bulk && !FE_NOFILL) { bulk_wait(&s_bar, parity); parity ^= 1u; }
        // phase word of the tile's local sample 0 (+ the table-rounding offset of the quantised NCO)
        const unsigned thb = p.theta0 + (unsigned)s_info[cur].lo * p.dtheta + (p.quantize ? (1u << 21) : 0u) + kFePhaseBias;
#ifdef CSDR_FE_SKIP
        if (CSDR_FE_SKIP & 2) {}
        else if (CSDR_FE_SKIP & 1) fe_run_top<S, 0>(p, smem, thb);
        else
#endif
        if (p.mix_mode == 0)      fe_run_top<S, 0>(p, smem, thb);
        else if (p.quantize) { if (p.mix_mode == 1) fe_run_top<S, 1 | 8>(p, smem, thb);
                               else                 fe_run_top<S, 2 | 8>(p, smem, thb); }
        else                 { if (p.mix_mode == 1) fe_run_top<S, 1>(p, smem, thb);
                               else                 fe_run_top<S, 2>(p, smem, thb); }
        __syncthreads();
        // the staging buffer has been consumed: start fetching the next tile of this CTA
        if (tile + gstep < p.ntiles) {
            if (FE_NOFILL) {} else if (s_info[nxt].bulk) { if (threadIdx.x == 0) fe_copy_staging<S>(p, &tmap, raw, s_info[nxt].lo, &s_bar); }
            else fe_fill_staging<S>(p, xs, hs, raw, s_info[nxt].lo);
        }
        // bookkeeping for the tile after next: by a thread of the last warp, which has no slot in the lower stages
        if (threadIdx.x == kFeNT - 32 && tile + 2 * gstep < p.ntiles) fe_tile_info<S, 1>(p, xs, tile + 2 * gstep, inv_st, s_info[nxt2]);

#ifdef CSDR_FE_SKIP
        if (!(CSDR_FE_SKIP & 4))
#endif
        if constexpr (S >= 2) fe_run_stages<S, 1, S - 2>(p, smem);
#ifdef CSDR_FE_SKIP
        if (!(CSDR_FE_SKIP & 8))
#endif
        {
            const int npush = (int)min((long long)G.Tc, p.K1 - p.K0 - (long long)tile * G.Tc);
            fe_resample_tile<G.Tc>(p, smem + G.off[0], ys, s_info[cur], npush, bank_s, rate_f);
        }
        __syncthreads();   // lower levels may be overwritten; a synchronously filled staging buffer is complete
        cur = nxt;
    }
}

// =============================================================================================================
// k_frontend_ws: k_frontend_direct with the CTA split into two groups of four warps that work on DIFFERENT tiles.
// Group A (producer) runs the first half-band stage of tile i+1 -- the NCO mix, bound by instruction issue -- while
// group B (consumer) runs the lower stages and the resampler of tile i -- bound by shared memory and short of slots
// for 256 threads.  The two halves are about the same amount of work, every stage fills at least 75 % of its group,
// and the hand-over level (S-1) is double-buffered; named barriers (bar.sync / bar.arrive) replace the CTA-wide ones.
enum { kBarA = 1, kBarB = 2, kBarFull0 = 3, kBarEmpty0 = 5 };       // FULL: 3, 4; EMPTY: 5, 6
constexpr int kFeWsGroup = 128;

template <int S, int MIX>
__device__ __forceinline__ void fe_ws_top(const FrontendParams &p, float2 *smem, float2 *dst, unsigned thb, int tid)
{
    constexpr FeGeom G = FeStd<S, 2>::G;
    constexpr int D2 = G.R[S - 2];
    fe_stage_top<G.m[S - 1], G.n[S - 1], false, D2, G.stride[S - 1], MIX, kFeWsGroup>(smem + G.off[S], dst, p.taps[S - 1], p.zeta,
                                                                                      thb, p.dtheta, tid);
}
template <int S, int s>
__device__ __forceinline__ void fe_ws_lower(const FrontendParams &p, float2 *smem, const float2 *src, int tid, int full_id)
{
    constexpr FeGeom G = FeStd<S, 2>::G;
    constexpr bool LAST = (s == 0);
    constexpr int D2 = LAST ? 1 : G.R[LAST ? 0 : s - 1];
    fe_stage_c<G.m[s], G.R[s], G.stride[s + 1], 0, G.n[s], LAST, D2, G.stride[s], kFeWsGroup>(
        (s == S - 2) ? src : smem + G.off[s + 1], smem + G.off[s], p.taps[s], p.zeta, tid);
    named_bar_sync(kBarB, kFeWsGroup);
    // the hand-over buffer has been consumed by the first lower stage: give it back to the producers
    if (s == S - 2) named_bar_arrive(kBarEmpty0 + (full_id - kBarFull0), 2 * kFeWsGroup);
    if constexpr (s > 0) fe_ws_lower<S, s - 1>(p, smem, src, tid, full_id);
}

template <int S>
__global__ void __launch_bounds__(2 * kFeWsGroup, 3) k_frontend_ws(const CSDR_GRID_CONSTANT FrontendParams p, const CSDR_GRID_CONSTANT FeTmap tmap)
{
    constexpr FeGeom G = FeStd<S, 2>::G;
    static_assert(S >= 2, "the consumer group needs at least one stage of its own");
    CSDR_DYN_SMEM_1K(smem_ws1k);
    unsigned char *smem_raw = smem_ws1k;
    float2 *smem = reinterpret_cast<float2 *>(smem_raw);
    float *bank_s = reinterpret_cast<float *>(smem_raw) + 2 * G.total_f2;
    float2 *raw = smem + G.off[S];
    __shared__ FeTileInfo s_infoA[3], s_infoB[3];
    __shared__ __align__(8) unsigned long long s_bar;

    const int npfb = 1 << p.bits;
    for (int i = threadIdx.x; i < npfb * kHsub; i += 2 * kFeWsGroup) {
        const int row = i / kHsub, col = i - row * kHsub;
        bank_s[row * (kHsub + 1) + col] = p.bank[i];
    }
    const float2 *xs = p.x + (long long)blockIdx.y * p.x_stride;
    const float2 *hs = p.hist + (long long)blockIdx.y * p.hcap;
    float2 *ys = p.y + (long long)blockIdx.y * p.y_stride;
    const double inv_st = 1.0 / (double)p.step;
    const float rate_f = 16777216.0f / (float)p.step;
    const int gstep = (int)gridDim.x, t0 = (int)blockIdx.x;
    const bool groupA = threadIdx.x < kFeWsGroup;
    const int tid = groupA ? threadIdx.x : threadIdx.x - kFeWsGroup;
    if (threadIdx.x == 0) bulk_init(&s_bar);
    __syncthreads();                                              // bank and barrier ready; from here on the groups part

    if (groupA) {
        // ---------------- producers: TMA tile -> first stage (mix) -> hand-over buffer
        unsigned parity = 0;
        if (tid == 0) {
            if (t0 < p.ntiles) fe_tile_info<S, 2>(p, xs, t0, inv_st, s_infoA[0]);
            if (t0 + gstep < p.ntiles) fe_tile_info<S, 2>(p, xs, t0 + gstep, inv_st, s_infoA[1]);
        }
        named_bar_sync(kBarA, kFeWsGroup);
        if (t0 < p.ntiles) {
            if (s_infoA[0].bulk) { if (tid == 0) fe_copy_staging<S, 2>(p, &tmap, raw, s_infoA[0].lo, &s_bar); }
            else fe_fill_staging<S, kFeWsGroup, 2>(p, xs, hs, raw, s_infoA[0].lo, tid);
        }
        named_bar_sync(kBarA, kFeWsGroup);
        int cur = 0, it = 0;
        for (int tile = t0; tile < p.ntiles; tile += gstep, it++) {
            const int nxt = cur == 2 ? 0 : cur + 1, nxt2 = nxt == 2 ? 0 : nxt + 1, b = it & 1;
            if (s_infoA[cur].bulk) { bulk_wait(&s_bar, parity); parity ^= 1u; }
            if (it >= 2) named_bar_sync(kBarEmpty0 + b, 2 * kFeWsGroup);      // consumers are done with this buffer
            float2 *dst = smem + (b ? G.off_alt : G.off[S - 1]);
            const unsigned thb = p.theta0 + (unsigned)s_infoA[cur].lo * p.dtheta + (p.quantize ? (1u << 21) : 0u) + kFePhaseBias;
            if (p.mix_mode == 0)      fe_ws_top<S, 0>(p, smem, dst, thb, tid);
            else if (p.quantize) { if (p.mix_mode == 1) fe_ws_top<S, 1 | 8>(p, smem, dst, thb, tid);
                                   else                 fe_ws_top<S, 2 | 8>(p, smem, dst, thb, tid); }
            else                 { if (p.mix_mode == 1) fe_ws_top<S, 1>(p, smem, dst, thb, tid);
                                   else                 fe_ws_top<S, 2>(p, smem, dst, thb, tid); }
            named_bar_sync(kBarA, kFeWsGroup);                     // staging consumed, hand-over buffer complete
            named_bar_arrive(kBarFull0 + b, 2 * kFeWsGroup);
            if (tile + gstep < p.ntiles) {
                if (s_infoA[nxt].bulk) { if (tid == 0) fe_copy_staging<S, 2>(p, &tmap, raw, s_infoA[nxt].lo, &s_bar); }
                else fe_fill_staging<S, kFeWsGroup, 2>(p, xs, hs, raw, s_infoA[nxt].lo, tid);
            }
            if (tid == kFeWsGroup - 32 && tile + 2 * gstep < p.ntiles) fe_tile_info<S, 2>(p, xs, tile + 2 * gstep, inv_st, s_infoA[nxt2]);
            named_bar_sync(kBarA, kFeWsGroup);                     // a synchronously filled tile / the new info are visible
            cur = nxt;
        }
        // match the consumers' last arrivals so that every barrier phase is complete when the CTA exits
        for (int k = (it >= 2 ? it - 2 : 0); k < it; k++) named_bar_sync(kBarEmpty0 + (k & 1), 2 * kFeWsGroup);
    } else {
        // ---------------- consumers: lower stages + resampler of the tile the producers finished last
        if (tid == 0) {
            if (t0 < p.ntiles) fe_tile_info<S, 2>(p, xs, t0, inv_st, s_infoB[0]);
            if (t0 + gstep < p.ntiles) fe_tile_info<S, 2>(p, xs, t0 + gstep, inv_st, s_infoB[1]);
        }
        named_bar_sync(kBarB, kFeWsGroup);
        int cur = 0, it = 0;
        for (int tile = t0; tile < p.ntiles; tile += gstep, it++) {
            const int nxt = cur == 2 ? 0 : cur + 1, nxt2 = nxt == 2 ? 0 : nxt + 1, b = it & 1;
            named_bar_sync(kBarFull0 + b, 2 * kFeWsGroup);
            if (tid == kFeWsGroup - 32 && tile + 2 * gstep < p.ntiles) fe_tile_info<S, 2>(p, xs, tile + 2 * gstep, inv_st, s_infoB[nxt2]);
            fe_ws_lower<S, S - 2>(p, smem, smem + (b ? G.off_alt : G.off[S - 1]), tid, kBarFull0 + b);
            {
                const int npush = (int)min((long long)G.Tc, p.K1 - p.K0 - (long long)tile * G.Tc);
                fe_resample_tile<G.Tc, kFeWsGroup>(p, smem + G.off[0], ys, s_infoB[cur], npush, bank_s, rate_f, tid);
            }
            named_bar_sync(kBarB, kFeWsGroup);                     // lower levels may be overwritten; new info visible
            cur = nxt;
        }
    }
}

}  // namespace csdr
