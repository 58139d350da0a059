// backend.cuh -- the sequential-state blocks that run at the decimated / per-channel rate:
//     iirfilt_crcf dc blocker   (Liquid.chs:575-589, liquid iirfilt.c "normal" form: v0 = x - a1 v1, y = v0 - v1)
//     agc_crcf + squelch gate   (Liquid.chs:693-717, liquid agc.c; Haskell zeroes y unless status == SIGNALHI)
//     freqdem                   (Liquid.chs:324-334, liquid freqdem.c)
// for `nlanes` independent sample sequences (streams or channelizer channels) of n samples each.
//
// Parallelisation over TIME (the reference is one sequential loop per stream).  Only the AGC gain is a genuinely
// sequential, non-linear recurrence; everything else is arranged around it as coalesced, fully parallel passes:
//   1. dc blocker (linear recurrence): k_dc_scan reduces every G-sample group to its zero-state response (one warp
//      per group, fp64), scans the groups of a CTA and looks back over the preceding CTAs: the exact filter state at every
//      group boundary is then one multiply-add away.
//   2. k_be_prep (one warp per group): dc-blocked samples y_dc[n] and their power p[n] = |y_dc[n]|^2.
//   3. k_agc_emit: the gain loop sees its input only through p[n]  (|x g|^2 = g^2 p), so the chain is
//      g^2 p -> one-pole filter -> g *= y2'^(-alpha/2): ~11 instructions per sample.  The loop is contractive, so
//      L-sample segments are run speculatively after a W-sample warm-up, each by its own thread.  The gains of every
//      32-sample block stay in shared memory and the CTA emits them at once (one sample per thread, one 32-sample word
//      per warp): ungated output y = y_dc g, discriminator value, "rssi > threshold" bit and sign bits by warp ballot.
//   4. k_be_finish (one cooperative launch): start states are verified against the predecessors' end states, misses
//      are refined in parallel and, as a last resort, repaired in stream order; the squelch state machine never feeds
//      back into the gain and is resolved EXACTLY on one threshold bit per sample; the gate is applied last as a mask.
// The result equals the sequential loop to within the stated tolerance in all cases (gate positions exactly, given
// the threshold bits); only the speed depends on the signal.
#pragma once
#include "platform.cuh"
#include "frontend.cuh"     // fe_phasor (optional rotation in k_be_prep)
#include <algorithm>

namespace csdr {

enum { SQ_UNKNOWN = 0, SQ_ENABLED, SQ_RISE, SQ_SIGNALHI, SQ_FALL, SQ_SIGNALLO, SQ_TIMEOUT, SQ_DISABLED };

struct LaneState {                 // carried across calls, one per lane
    float dc_re, dc_im;            // dc blocker v1
    float g, y2p;                  // agc gain, filtered output energy
    int mode; unsigned timer;      // squelch FSM
    float fm_re, fm_im;            // freqdem r_prime: the last UNGATED agc output sample
};
struct SegState { float g, y2p; };

struct FsmState;
struct BackendParams {
    const float2 *in; long long in_lane_stride;
    void *out; long long out_lane_stride;      // float (demod != 0) or float2 elements
    void *const *out_table;                    // optional [nlanes]: lane i writes to out_table[i] instead of out + i * stride
                                               // (the caller's own per-channel buffers: no copy behind the back end)
    int n, nlanes;
    int L, W, G, nseg, ngrp;
    int has_dc, has_agc, demod;                // demod: 0 none (cf32 out), 1 fm (float out)
    float alpha; float one_minus_alpha_f; float neg_half_alpha;
    float g_thr;                               // rssi > threshold  <=>  g < g_thr  (bisected on the host)
    unsigned timeout; float fm_ref;
    int squelch_enabled;
    int exact_math;                            // 1: library expf/logf/atan2f instead of the SFU forms (slower)
    int gate;                                  // 1: zero the output unless squelch status == SIGNALHI (Liquid.chs:700-704)
    LaneState *lane;
    const float2 *ydc; long long ydc_stride;   // dc-blocked samples (= in when there is no dc blocker)
    const float *pw; long long pw_stride;      // [nlanes][pw_stride] power of the dc-blocked samples
    float *g_first; float2 *y_first;           // [nlanes] gain / ungated output before the chunk's first sample
    float2 *y_end;                             // [nlanes] ungated output of the chunk's last sample
    float2 *seg_ylast;                         // [nlanes][nseg] ungated output of every segment's last sample
    unsigned *barrier;                         // [4] grid barrier of k_be_finish: arrivals, generation, exit count
    SegState *seg_start, *seg_end;             // [nlanes][nseg] gain-loop state at segment boundaries
    int nwords, FW;                            // 32-sample words per lane; FSM replay length in segments
    unsigned *exbits, *gatebits;               // [nlanes][nwords] threshold-exceeded / gate-open bit per sample
    unsigned *sgnr, *sgni;                     // [nlanes][nwords] sign bits of the ungated agc output (discriminator
                                               // values next to a closed gate are signed-zero artefacts: +-pi or 0)
    unsigned *prev_sign;                       // [nlanes] sign bits (re | im << 1) of the sample before this chunk
    FsmState *fsm_start, *fsm_end;             // [nlanes][nseg]
    unsigned *prev_gate;                       // [nlanes] gate of the sample before this chunk
    unsigned *first_bad;                       // [nlanes][2] first segment whose start state does not continue its
                                               // predecessor (gain loop, squelch FSM); 0xffffffff = none
    unsigned *bad_list; unsigned *bad_count; unsigned bad_cap;   // [2][bad_cap] segments to refine after the first / second
                                               // verification; [3] counters: round 0, round 1, squelch-FSM misses
    unsigned long long *fixups;                // [3] segments re-run in order: gain loop, squelch FSM; refined in parallel
};

// ------------------------------------------------------------------------------------------ dc blocker
// v[n] = x[n] + c v[n-1] (c = 1 - alpha), y[n] = v[n] - v[n-1]: a linear recurrence, evaluated in fp64 as affine maps
// (k_dc_scan below); the float32 recurrence of iirfilt then runs over G/32 samples per lane from the exact state.
#ifndef CSDR_DC_GB          // block-geometry experiments (scripts/exp_build.sh): groups per block, warps and CTAs per SM
#define CSDR_DC_GB 8
#endif
#ifndef CSDR_DC_WARPS
#define CSDR_DC_WARPS 2
#endif
#ifndef CSDR_DC_MINB
#define CSDR_DC_MINB 8
#endif
constexpr int kDcGB = CSDR_DC_GB;
constexpr int kDcWarps = CSDR_DC_WARPS;
constexpr int kDcMinB = CSDR_DC_MINB;

struct DcParams {
    const float2 *in; long long in_lane_stride;
    float2 *out; long long out_lane_stride;    // dc-blocked samples (nullptr: not wanted)
    float *pw; long long pw_stride;            // |y|^2 (nullptr: not wanted)
    int n, nlanes, G, ngrp, nblk, has_dc;
    int rot; unsigned rot_theta, rot_dtheta; int rot_quantize;   // rot = 1: multiply the output by conj(NCO phasor) (channelizer pre-rotation)
    double c;                                  // 1 - alpha  (= -a1)
    double cS[5];                              // c^(S d), S = G/32 samples per lane, d = 1, 2, 4, 8, 16
    float a1;
    const double *powA;                        // [kDcGB + 1]     A^k, A = c^G
    // k_dc_scan: blocks of kDcGB groups publish their zero-state response and look back over their predecessors
    SelfValid16 *agg;                          // [nlanes][nblk]  zero-state response of block b (fp64 pair); all-ones until
                                               // it is published (reset by a memset before every launch)
    unsigned *ticket;                          // [2] next block to hand out, CTAs that have left (reset by the last one)
    const double *powAB; int depth;            // [depth + 1] (c^(G kDcGB))^k; blocks further back than `depth` have decayed
                                               // below 1e-13 of the state
    const float2 *dc_in; float2 *dc_out;       // [nlanes] filter state v1 before the chunk / after it (two buffers: blocks
                                               // of one launch read the old state while the last block writes the new one)
};

// the S = G/32 consecutive samples of this lane in group j (zero past the end of the chunk); 16-byte loads when the
// lane's samples are 16-byte aligned
template <int S>
__device__ __forceinline__ void dc_load(const float2 *__restrict__ x, int n, int i0, float2 (&v)[S])
{
    if (S >= 2 && i0 + S <= n && (reinterpret_cast<uintptr_t>(x + i0) & 15) == 0) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x + i0);
#pragma unroll
        for (int k = 0; k < S / 2; k++) { const float4 q = __ldg(x4 + k); v[2 * k] = cf(q.x, q.y); v[2 * k + 1] = cf(q.z, q.w); }
    } else {
#pragma unroll
        for (int k = 0; k < S; k++) v[k] = (i0 + k < n) ? __ldg(x + i0 + k) : cf(0.f, 0.f);
    }
}

// Inclusive warp scan of the zero-state responses: every lane covers S samples, i.e. its map is  s -> q s + a  with the
// same q = c^S for all lanes, so after the step with distance d a lane that has a partner (l >= d) covers exactly d
// lanes of its own and the partner's value enters with q^d = cS[log2 d].  Afterwards lane l holds the zero-state
// response at the end of lane l's samples.  float32 is enough INSIDE a group: an error e of the state a lane starts
// from reaches its outputs as alpha e c^n (y = x - alpha v1), 1e-6 of the state's rounding is 1e-9 of the signal; what
// accumulates over long distances -- the group and block totals -- is kept in fp64.
__device__ __forceinline__ void dc_warp_scan(float &ar, float &ai, const float (&qS)[5])
{
    const int l = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int d = 1 << s;
        const float pr = __shfl_up_sync(0xffffffffu, ar, d), pi = __shfl_up_sync(0xffffffffu, ai, d);
        if (l >= d) { ar = fmaf(pr, qS[s], ar); ai = fmaf(pi, qS[s], ai); }
    }
}
// q^e for 0 <= e < 32
__device__ __forceinline__ double dc_pow32(const double (&cS)[5], int e)
{
    double r = 1.0;
#pragma unroll
    for (int s = 0; s < 5; s++) if ((e >> s) & 1) r *= cS[s];
    return r;
}

// One pass over the samples: a CTA takes blocks of kDcGB groups (1024 samples at G = 128) in stream order (ticket).
// Small CTAs (two warps, eight per SM): the block-wide steps below are separated by barriers that idle the whole CTA, so
// many small CTAs overlap them better than a few large ones (measured per 2^27 samples: 64 groups x 8 warps 540 us,
// 16 x 4 503 us, 8 x 2 472 us).
//   0. The block arrives in shared memory by ONE bulk copy (cp.async.bulk, mbarrier completion) issued while the previous
//      block is still being worked on: no registers, no load instructions, HBM latency hidden.
//   1. Every lane takes its 4 consecutive samples, a float32 warp scan gives each lane the zero-state response up to its
//      samples (kept in registers) and the group's total; the totals of the block's groups are combined in fp64 and
//      the block's zero-state response is published (a self-validating 16-byte value: no flag, no fence).
//   2. Look-back: the filter state before the block is  sum_k AB^(k-1) agg[b-k]  (+ AB^b x the state carried from the
//      previous call), AB = c^1024 = 0.60 for alpha = 5e-4, so `depth` = 59 predecessors settle it to 1e-13 (one
//      16-byte load per thread, all in flight at once).
//   3. Every lane gets the exact state before its samples, runs iirfilt's float32 recurrence over them and writes the
//      dc-blocked samples (optionally pre-rotated for the channelizer) and their power.
// The loop is software-pipelined over two blocks: steps 0-1 of the CTA's NEXT block run before steps 2-3 of the current
// one, so every block's response has been published a whole block time before anybody looks back for it (with the steps
// in order the CTAs fall into lock-step with the slowest of their seven predecessors: 26 % of the kernel's time, measured
// by taking the wait out).  Both blocks' samples stay in registers; the one staging tile is refilled for the block after
// next while the scan and the output pass run.
// Each sample is read once and written once; nothing else travels through HBM.  out may alias in.
constexpr size_t kDcSmem = sizeof(float2) * kDcGB * 128;       // the staged block (G <= 128)
template <int S>
__global__ void __launch_bounds__(32 * kDcWarps, kDcMinB) k_dc_scan(const DcParams p)
{
    constexpr int GW = kDcGB / kDcWarps;               // groups per warp
    CSDR_DYN_SMEM(smem_raw);
    float2 *tile = reinterpret_cast<float2 *>(smem_raw);
    __shared__ double sr[2][kDcGB], si[2][kDcGB];      // zero-state response at the END of each group, block-local (current / next block)
    __shared__ double s_red[2][kDcWarps];
    __shared__ double s_cr, s_ci;
    __shared__ int s_ticket;
    __shared__ __align__(8) unsigned long long s_bar;
    const int t = threadIdx.x, w = t >> 5, l = t & 31;
    const double A = p.powA[1];
    const double ql = dc_pow32(p.cS, l);                                      // q^l, q = c^S
    const float cf32 = (float)p.c;
    const float qS[5] = {(float)p.cS[0], (float)p.cS[1], (float)p.cS[2], (float)p.cS[3], (float)p.cS[4]};
    const int total = p.nlanes * p.nblk;
    const int blk = kDcGB * p.G;                                              // samples per block
    // a block that lies inside the chunk at a 16-byte aligned address is fetched by the bulk copy; the others (the last
    // block of a lane, unaligned chunks) are read with ordinary loads
    auto block_src = [&](int tkk) { const int ln = tkk / p.nblk, bb = tkk - ln * p.nblk; return p.in + (long long)ln * p.in_lane_stride + (long long)bb * blk; };
    auto block_bulk = [&](int tkk) {
        const int bb = tkk % p.nblk;
        return (long long)(bb + 1) * blk <= p.n && (reinterpret_cast<uintptr_t>(block_src(tkk)) & 15) == 0;
    };
    unsigned parity = 0;
    // ---- 0. samples of block tkk into registers (from the staged tile, or straight from global memory)
    auto load_block = [&](int tkk, float2 (&v)[GW][S]) {
#ifdef CSDR_EMU
        __syncthreads();                                    // the emulated bulk copy is a memcpy by thread 0
#endif
        if (block_bulk(tkk)) {
            bulk_wait(&s_bar, parity); parity ^= 1u;
#pragma unroll
            for (int k = 0; k < GW; k++) {
                const float4 *src = reinterpret_cast<const float4 *>(tile + (w * GW + k) * p.G + l * S);
#pragma unroll
                for (int q = 0; q < S / 2; q++) { const float4 f = src[q]; v[k][2 * q] = cf(f.x, f.y); v[k][2 * q + 1] = cf(f.z, f.w); }
                if (S == 1) v[k][0] = tile[(w * GW + k) * p.G + l];
            }
        } else {
            const int lane = tkk / p.nblk, b = tkk - lane * p.nblk;
            const float2 *__restrict__ xx = p.in + (long long)lane * p.in_lane_stride;
#pragma unroll
            for (int k = 0; k < GW; k++) {
                const int j = b * kDcGB + w * GW + k;
                dc_load<S>(xx, (j < p.ngrp) ? p.n : 0, j * p.G + l * S, v[k]);
            }
        }
    };
    // ---- 1. scan of a block held in registers (float32 inside a group, fp64 across groups); publishes its response
    auto scan_block = [&](int tkk, const float2 (&v)[GW][S], float (&er)[GW], float (&ei)[GW], double *srb, double *sib) {
        double Rr = 0.0, Ri = 0.0;                          // zero-state response of this warp's run of groups
#pragma unroll
        for (int k = 0; k < GW; k++) {
            float ar = 0.f, ai = 0.f;
#pragma unroll
            for (int q = 0; q < S; q++) { ar = fmaf(ar, cf32, v[k][q].x); ai = fmaf(ai, cf32, v[k][q].y); }
            dc_warp_scan(ar, ai, qS);
            er[k] = __shfl_up_sync(0xffffffffu, ar, 1); ei[k] = __shfl_up_sync(0xffffffffu, ai, 1);
            if (l == 0) { er[k] = 0.f; ei[k] = 0.f; }
            const float Tr = __shfl_sync(0xffffffffu, ar, 31), Ti = __shfl_sync(0xffffffffu, ai, 31);
            Rr = Rr * A + (double)Tr; Ri = Ri * A + (double)Ti;
            if (l == 0) { srb[w * GW + k] = Rr; sib[w * GW + k] = Ri; }
        }
        __syncthreads();
        double vr = 0.0, vi = 0.0;
        if (t < kDcGB) {
            const int tw = t / GW, tkk2 = t - tw * GW;
            // state at the start of warp tw's run inside this block
            double pr = 0.0, pi = 0.0;
            for (int q = 0; q < tw; q++) {
                const double a = p.powA[GW * (tw - 1 - q)];
                pr += srb[q * GW + GW - 1] * a; pi += sib[q * GW + GW - 1] * a;
            }
            const double a = p.powA[tkk2 + 1];
            vr = srb[t] + pr * a; vi = sib[t] + pi * a;
        }
        __syncthreads();
        if (t < kDcGB) { srb[t] = vr; sib[t] = vi; }
        if (t == kDcGB - 1) sv16_store(p.agg + tkk, vr, vi);
    };
    // the staged tile has been read by everybody: take the next ticket and start its copy
    auto next_ticket = [&]() {
        if (t == 0) s_ticket = (int)atomicAdd(p.ticket, 1u);
        __syncthreads();
        const int tkk = s_ticket;
        if (t == 0 && tkk < total && block_bulk(tkk)) bulk_copy_g2s(tile, block_src(tkk), (unsigned)(blk * sizeof(float2)), &s_bar);
        return tkk;
    };

    if (t == 0) bulk_init(&s_bar);
    __syncthreads();
    int tk = next_ticket();
    float2 v[GW][S];
    float er[GW], ei[GW];
    int cb = 0;                                               // which half of sr / si belongs to the current block
    int tk_next = total;
    if (tk < total) {
        load_block(tk, v);
        __syncthreads();                                      // (s_ticket has been read by everybody)
        tk_next = next_ticket();
        scan_block(tk, v, er, ei, sr[0], si[0]);
    }
    float2 vn[GW][S];
    float ern[GW], ein[GW];
    // one iteration of the pipeline: (v, er, ei) = current block, (vn, ern, ein) = where the next block goes.  Called with
    // the two register sets alternating, so the next block becomes the current one without a copy.
    auto iteration = [&](float2 (&v)[GW][S], float (&er)[GW], float (&ei)[GW], float2 (&vn)[GW][S], float (&ern)[GW], float (&ein)[GW]) {
        const int lane = tk / p.nblk, b = tk - lane * p.nblk;
        const bool bulk = block_bulk(tk);
        // ---- steps 0-1 of the next block
        int tk_next2 = total;
        if (tk_next < total) {
            load_block(tk_next, vn);
            __syncthreads();
            tk_next2 = next_ticket();
            scan_block(tk_next, vn, ern, ein, sr[cb ^ 1], si[cb ^ 1]);
        }
        // ---- 2. look back (thread k waits for block b - 1 - k, b - 1 - k - 256, ...)
        double cr = 0.0, ci = 0.0;
        for (int k = t; k < p.depth && k < b; k += blockDim.x) {
            const int src = tk - 1 - k;
            double ax, ay;
#if defined(CSDR_DC_SKIP) && (CSDR_DC_SKIP & 2)
            sv16_load(p.agg + src, ax, ay);
#else
            while (!sv16_load(p.agg + src, ax, ay)) {}
#endif
            const double m = p.powAB[k];
            cr += ax * m; ci += ay * m;
        }
        if (t == 0 && b <= p.depth) { const float2 v0 = p.dc_in[lane]; const double m = p.powAB[b]; cr += (double)v0.x * m; ci += (double)v0.y * m; }
        if (w == 0 || p.depth > 32) {
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) { cr += __shfl_xor_sync(0xffffffffu, cr, d); ci += __shfl_xor_sync(0xffffffffu, ci, d); }
        }
        if (l == 0) { s_red[0][w] = cr; s_red[1][w] = ci; }
        __syncthreads();                                    // (also: sr / si of the current block are complete)
        if (t == 0) {
            double a = 0.0, bb = 0.0;
            for (int q = 0; q < kDcWarps; q++) { a += s_red[0][q]; bb += s_red[1][q]; }
            s_cr = a; s_ci = bb;
        }
        __syncthreads();
        const double carry_r = s_cr, carry_i = s_ci;
        const double *srb = sr[cb], *sib = si[cb];
        // ---- 3. apply
        float2 *__restrict__ yo = p.out ? p.out + (long long)lane * p.out_lane_stride : nullptr;
        float *__restrict__ wo = p.pw ? p.pw + (long long)lane * p.pw_stride : nullptr;
#pragma unroll
        for (int k = 0; k < GW; k++) {
            const int jl = w * GW + k, j = b * kDcGB + jl;
            if (j >= p.ngrp) break;                                           // warp-uniform
            const int i0 = j * p.G + l * S;
            // state before group j (fp64), then before this lane's samples
            const double pa = p.powA[jl];
            const double Vr = (jl ? srb[jl - 1] : 0.0) + carry_r * pa, Vi = (jl ? sib[jl - 1] : 0.0) + carry_i * pa;
            float v1r = (float)((double)er[k] + Vr * ql), v1i = (float)((double)ei[k] + Vi * ql);
#pragma unroll
            for (int q = 0; q < S; q++) {
                if (bulk || i0 + q < p.n) {
                    const float v0r = __fsub_rn(v[k][q].x, __fmul_rn(p.a1, v1r));
                    const float v0i = __fsub_rn(v[k][q].y, __fmul_rn(p.a1, v1i));
                    v[k][q] = cf(__fsub_rn(v0r, v1r), __fsub_rn(v0i, v1i));
                    v1r = v0r; v1i = v0i;
                }
            }
            // the lane that holds the chunk's last sample stores the filter state for the next call
            if (i0 < p.n && i0 + S >= p.n) p.dc_out[lane] = cf(v1r, v1i);
#if defined(CSDR_DC_SKIP) && (CSDR_DC_SKIP & 1)
            if (false) {
#else
            if (p.rot) {
#endif
                // the channelizer's pre-rotation (nco_crcf_mix_block_down, Liquid.chs:847) rides on this pass
#pragma unroll
                for (int q = 0; q < S; q++) {
                    const float2 ph = fe_phasor(p.rot_theta + (unsigned)(i0 + q) * p.rot_dtheta, p.rot_quantize);
                    v[k][q] = cf(v[k][q].x * ph.x + v[k][q].y * ph.y, v[k][q].y * ph.x - v[k][q].x * ph.y);
                }
            }
#if defined(CSDR_DC_SKIP) && (CSDR_DC_SKIP & 4)
            if (yo && v[k][0].x == 1.2345e-30f) {
#else
            if (yo) {
#endif
                float2 *y = yo + i0;
                if (S == 4 && (bulk || i0 + S <= p.n) && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
                    reinterpret_cast<float4 *>(y)[0] = make_float4(v[k][0].x, v[k][0].y, v[k][1 % S].x, v[k][1 % S].y);
                    reinterpret_cast<float4 *>(y)[1] = make_float4(v[k][2 % S].x, v[k][2 % S].y, v[k][3 % S].x, v[k][3 % S].y);
                } else {
#pragma unroll
                    for (int q = 0; q < S; q++) if (i0 + q < p.n) y[q] = v[k][q];
                }
            }
            if (wo) {
                float *wp = wo + i0;
                float e[S];
#pragma unroll
                for (int q = 0; q < S; q++) e[q] = __fadd_rn(__fmul_rn(v[k][q].x, v[k][q].x), __fmul_rn(v[k][q].y, v[k][q].y));
                if (S == 4 && (bulk || i0 + S <= p.n)) *reinterpret_cast<float4 *>(wp) = make_float4(e[0], e[1 % S], e[2 % S], e[3 % S]);   // pw_stride % 4 == 0
                else {
#pragma unroll
                    for (int q = 0; q < S; q++) if (i0 + q < p.n) wp[q] = e[q];
                }
            }
        }
        __syncthreads();      // s_red / s_cr and this block's half of sr / si are reused
        tk = tk_next; tk_next = tk_next2; cb ^= 1;
    };
    while (tk < total) {
        iteration(v, er, ei, vn, ern, ein);
        if (tk >= total) break;
        iteration(vn, ern, ein, v, er, ei);
    }
    // tickets for the next launch: reset by the last CTA to leave
    if (t == 0) {
        __threadfence();
        if (atomicAdd(p.ticket + 1, 1u) == gridDim.x - 1) { p.ticket[0] = 0; p.ticket[1] = 0; }
    }
}

// power of the samples (no dc blocker in this back end: per-channel lanes behind a channelizer whose kernels did not
// write it themselves), one warp per group
template <int S>
__global__ void __launch_bounds__(256) k_be_prep(const DcParams p)
{
    const int lane = blockIdx.y, l = threadIdx.x & 31;
    const float2 *__restrict__ x = p.in + (long long)lane * p.in_lane_stride;
    float *__restrict__ wo = p.pw + (long long)lane * p.pw_stride;
    const int jstep = (int)((gridDim.x * blockDim.x) >> 5);
    for (int j = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); j < p.ngrp; j += jstep) {     // warp-uniform
        const int i0 = j * p.G + l * S;
        float2 v[S];
        dc_load<S>(x, p.n, i0, v);
        float *w = wo + i0;
        float e[S];
#pragma unroll
        for (int q = 0; q < S; q++) e[q] = __fadd_rn(__fmul_rn(v[q].x, v[q].x), __fmul_rn(v[q].y, v[q].y));
        if (S == 4 && i0 + S <= p.n) *reinterpret_cast<float4 *>(w) = make_float4(e[0], e[1 % S], e[2 % S], e[3 % S]);   // pw_stride % 4 == 0
        else {
#pragma unroll
            for (int q = 0; q < S; q++) if (i0 + q < p.n) w[q] = e[q];
        }
    }
}

// ------------------------------------------------------------------------------------------ agc gain loop
__device__ __forceinline__ void fsm_step(int &mode, unsigned &timer, bool ex, unsigned timeout)
{
    // AGC(_squelch_update_mode), liquid agc.c
    switch (mode) {
    case SQ_ENABLED:  mode = ex ? SQ_RISE : SQ_ENABLED; break;
    case SQ_RISE:     mode = ex ? SQ_SIGNALHI : SQ_FALL; break;
    case SQ_SIGNALHI: mode = ex ? SQ_SIGNALHI : SQ_FALL; break;
    case SQ_FALL:     mode = ex ? SQ_SIGNALHI : SQ_SIGNALLO; timer = timeout; break;
    case SQ_SIGNALLO:
        timer--;
        if (timer == 0) mode = SQ_TIMEOUT;
        else if (ex)    mode = SQ_SIGNALHI;
        break;
    case SQ_TIMEOUT:  mode = SQ_ENABLED; break;
    default: break;
    }
}

__device__ __forceinline__ float be_rcp(float x)
{
#ifdef CSDR_EMU
    return 1.0f / x;
#else
    float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}

// arg(x + jy) with a degree-8 minimax polynomial for atan on [0, 1] (max error 1.1e-7 rad, float32-limited; the
// library atan2f is ~3x the instructions).  Exact zeros keep the library's signed-zero semantics.
// Branch-free: |a| = min / max is 0 for two zeros (the quotient is replaced, not computed), the quadrant comes from the SIGN
// BITS, so the signed-zero cases come out as the library's (atan2(+-0, -0) = +-pi, atan2(+-0, +0) = +-0); an infinite
// operand gives min / max = 0 or NaN exactly where atan2f gives a multiple of pi/2 or pi/4 -- only the latter (both
// infinite) differs (NaN instead of +-pi/4, +-3pi/4), as does a NaN operand's payload.
__device__ __forceinline__ float be_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    // min / max by one reciprocal.  Subnormal operands are real here (the gain loop overshoots after a stretch of digital
    // silence and the products of two tiny outputs underflow), so both are lifted by 2^24 when the larger one is subnormal;
    // the clamp then keeps two exact zeros at 0 * rcp(FLT_MIN) = 0, the quotient the signed-zero cases need, without a
    // select of its own (7 instructions; the library division with its zero guard: 9).
    const float lift = (mx < 1.17549435e-38f) ? 16777216.f : 1.f;
    const float a = (mn * lift) * be_rcp(fmaxf(mx * lift, 1.17549435e-38f));
    const float s = a * a;
    float r = 0.0028340641874819994f;
    r = fmaf(r, s, -0.016005029901862144f);
    r = fmaf(r, s, 0.042587608098983765f);
    r = fmaf(r, s, -0.07495445758104324f);
    r = fmaf(r, s, 0.10636754333972931f);
    r = fmaf(r, s, -0.14202570915222168f);
    r = fmaf(r, s, 0.19992484152317047f);
    r = fmaf(r, s, -0.3333306610584259f);
    r = fmaf(r, s, 1.0f);
    r *= a;
    if (ay > ax) r = 1.57079637f - r;
    if (__float_as_int(x) < 0) r = 3.14159274f - r;
    return copysignf(r, y);
}


// one step of AGC(_execute), liquid agc.c, seen from the gain:  y = x g, y2 = |y|^2 = g^2 |x|^2,
// y2' = (1-alpha) y2' + alpha y2 (liquid evaluates this in double and rounds to float; the float32 FMA differs from it
// by at most one ulp of y2'), g *= y2'^(-alpha/2) unless y2' <= 1e-6, g <= 1e6.  Default: SFU exp2/log2 (each step
// good to ~3e-7 relative; the loop is contractive, so the gain stays within ~1e-6 of the libm evaluation).
__device__ __forceinline__ float be_lg2(float x)
{
#ifdef CSDR_EMU
    return log2f(x);
#else
    float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}
__device__ __forceinline__ float be_ex2(float x)
{
#ifdef CSDR_EMU
    return exp2f(x);
#else
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#endif
}
struct AgcCoef { float alpha, oma, nha; };          // alpha, 1 - alpha, -alpha / 2: kept in registers by the loops
template <bool EXACT>
__device__ __forceinline__ void agc_step(const AgcCoef &p, float &g, float &g2, float &y2p, float pw)
{
    if (EXACT) {
        const float y2 = __fmul_rn(__fmul_rn(g, g), pw);
        y2p = fmaf(p.oma, y2p, __fmul_rn(p.alpha, y2));
        const float f = expf(p.nha * logf(y2p));
        g *= (y2p > 1e-6f) ? f : 1.0f;
        g = fminf(g, 1e6f);
    } else {
        // the recurrence closes through g^2, so g^2 is carried as a state of its own and updated with the squared
        // factor: the dependent chain per sample is FMA -> lg2 -> mul -> ex2 -> mul -> mul -> min, and the gain itself
        // only trails it.  (g2 / g^2 drifts by rounding, ~1e-7 sqrt(samples); it is re-derived from g at the start of
        // every segment.)
        y2p = fmaf(__fmul_rn(p.alpha, pw), g2, __fmul_rn(p.oma, y2p));
        const float lg = be_lg2(y2p);
        const bool ok = y2p > 1e-6f;
        const float f = be_ex2(__fmul_rn(p.nha, lg)), f2 = __fmul_rn(f, f);
        g2 = fminf(ok ? __fmul_rn(g2, f2) : g2, 1e12f);
        g = fminf(ok ? __fmul_rn(g, f) : g, 1e6f);
    }
}

__device__ __forceinline__ bool be_close(float a, float b, float atol = 0.f)
{
    return fabsf(a - b) <= 1e-5f * fmaxf(fabsf(a), fabsf(b)) + atol;   // fp32 rounding keeps two runs ~1e-6 apart
}
__device__ __forceinline__ bool be_match(const SegState &a, const SegState &b)
{
    return be_close(a.g, b.g) && be_close(a.y2p, b.y2p);
}

// what the chunk's first sample needs from the previous chunk (the lane state is overwritten before k_be_emit runs)
__device__ __forceinline__ void be_save_first(const BackendParams &p, int lane)
{
    const LaneState ls = p.lane[lane];
    p.g_first[lane] = ls.g;
    p.y_first[lane] = cf(ls.fm_re, ls.fm_im);
    p.prev_sign[lane] = ((unsigned)__float_as_int(ls.fm_re) >> 31) | (((unsigned)__float_as_int(ls.fm_im) >> 31) << 1);
}

// start state of a speculative segment whose warm-up window starts at w0 > 0: unit output energy for the mean power
// e of the window's first 16 samples
__device__ __forceinline__ void agc_guess(float e, float &g, float &y2p)
{
    g = (e > 1e-30f) ? rsqrtf(e) : 1e6f;
    if (g > 1e6f) g = 1e6f;
    y2p = 1.0f;
}

__device__ __forceinline__ float *be_out_f(const BackendParams &p, int lane)
{
    return p.out_table ? (float *)p.out_table[lane] : (float *)p.out + (long long)lane * p.out_lane_stride;
}
__device__ __forceinline__ float2 *be_out_c(const BackendParams &p, int lane)
{
    return p.out_table ? (float2 *)p.out_table[lane] : (float2 *)p.out + (long long)lane * p.out_lane_stride;
}

// what the emission of a lane needs, fetched once per kernel
struct BeEmitCtx {
    float *of; float2 *oc;                     // output of this lane (discriminator / cf32)
    unsigned *exb, *sgr, *sgi;                 // bit planes of this lane
    float g_thr, fm_ref;
    int n;
    bool skip_closed;                          // words without a threshold-exceeding sample are left to the gate pass
};
__device__ __forceinline__ BeEmitCtx be_emit_ctx(const BackendParams &p, int lane)
{
    BeEmitCtx c;
    c.of = be_out_f(p, lane); c.oc = be_out_c(p, lane);
    c.exb = p.exbits + (long long)lane * p.nwords; c.sgr = p.sgnr + (long long)lane * p.nwords; c.sgi = p.sgni + (long long)lane * p.nwords;
    c.g_thr = p.g_thr; c.fm_ref = p.fm_ref; c.n = p.n;
    c.skip_closed = p.gate != 0;
    return c;
}

// ------------------------------------------------------------------------------------------ emission of one word
// 32 consecutive samples of a lane, one per thread of a warp: ungated output y = y_dc * (gain before the sample),
// threshold bit = (gain after the sample) < g_thr, sign bits, discriminator m = arg(conj(y[n-1]) y[n]) / (2 pi kf)
// (freqdem_demodulate) or the cf32 sample itself.  `yprev0`: ungated output of the sample before the word (lane 0).
// FULL: all 32 samples exist (no bounds checks).  Returns this lane's y (the caller carries the last one on).
template <bool AGC, bool FM, bool EXACT, bool FULL>
__device__ __forceinline__ float2 be_emit_word(const BeEmitCtx &c, int u0, float2 xv, float g0, float ga, float2 yprev0)
{
    const int l = threadIdx.x & 31;
    const bool in = FULL || (u0 + l < c.n);
    float2 y = xv;
    if (AGC) y = cf(__fmul_rn(xv.x, g0), __fmul_rn(xv.y, g0));
    float2 yp;
    yp.x = __shfl_up_sync(0xffffffffu, y.x, 1); yp.y = __shfl_up_sync(0xffffffffu, y.y, 1);
    if (l == 0) yp = yprev0;
    if (AGC) {
        const int wd = u0 >> 5;
        const unsigned ex = __ballot_sync(0xffffffffu, in && ga < c.g_thr);       // rssi = -20 log10(g) > threshold
        if (FM) {
            const unsigned sr = __ballot_sync(0xffffffffu, in && (__float_as_int(y.x) < 0));
            const unsigned si = __ballot_sync(0xffffffffu, in && (__float_as_int(y.y) < 0));
            if (l < 3) { unsigned *dst = l == 0 ? c.exb : l == 1 ? c.sgr : c.sgi; dst[wd] = l == 0 ? ex : l == 1 ? sr : si; }
        } else if (l == 0) c.exb[wd] = ex;
        // The squelch can only be open on a sample that exceeds the threshold (every path into SIGNALHI takes the
        // "exceeded" branch of the state machine), so a word without such a sample is closed throughout: the gate pass
        // only has to fix the signed-zero artefact next to an open sample; zeros are written here.
        if (c.skip_closed && ex == 0) {
            if (in) { if (FM) c.of[u0 + l] = 0.f; else c.oc[u0 + l] = cf(0.f, 0.f); }
            return y;
        }
    }
    if (in) {
        if (FM) {
            const float re = __fadd_rn(__fmul_rn(yp.x, y.x), __fmul_rn(yp.y, y.y));
            const float im = __fsub_rn(__fmul_rn(yp.x, y.y), __fmul_rn(yp.y, y.x));
            c.of[u0 + l] = (EXACT ? atan2f(im, re) : be_atan2(im, re)) * c.fm_ref;
        } else {
            c.oc[u0 + l] = y;
        }
    }
    return y;
}

// ------------------------------------------------------------------------------------------ gain loop + emission
// One thread per L-sample segment, kAgcT consecutive segments of a lane per CTA.  The threads walk their windows
// [b0 - W, b1) in lock-step, 32 samples at a time: each warp fetches the 32-sample blocks of its 32 threads with one
// coalesced 128-byte load per thread-row into a padded shared-memory tile (row stride 33: column access is
// conflict-free) and every thread runs the recurrence over its row.  In the emitting part of the window the rows are
// overwritten with the gains, and the CTA turns them into output right away: warp w takes rows 32 w .. 32 w + 31, one
// 32-sample word per row and iteration, lane = sample (coalesced 256-byte reads of y_dc, coalesced stores): the gains
// never travel through HBM.  17 KB of shared memory per CTA whatever L and W are, so every chain of a call is resident at
// once and the kernel is bound by the latency of one window, not by waves.
// The first sample of a segment takes its gain and its predecessor's output from the segment's OWN warm-up (verified to
// 1e-5 against the predecessor's end state by k_be_finish; typically equal to ~1e-7).
constexpr int kAgcT = 128, kAgcB = 32;
constexpr int kAgcRow = kAgcB + 1;      // a row of the power / output tile: [0] threshold word of the block, [1 + k] power -> output of sample k
constexpr int kAgcXRow = kAgcB + 2;     // a row of the sample tile (float2): 272 bytes, 16-byte aligned, at most 2-way bank conflicts
constexpr size_t kAgcSmem = sizeof(float) * kAgcT * kAgcRow + sizeof(float2) * kAgcT * kAgcXRow;

// per-thread state of the emission: previous ungated output and the bit words of the current block
struct AgcEmitState { float2 yp; unsigned ex, sr, si; };
// (w << 1) | sign bit of v
__device__ __forceinline__ unsigned be_shift_in_sign(unsigned w, float v)
{
#ifdef CSDR_EMU
    return (w << 1) | ((unsigned)__float_as_int(v) >> 31);
#else
    return __funnelshift_l((unsigned)__float_as_int(v), w, 1);
#endif
}

// samples [K0, K0 + NK) of the current block of this thread's row: gain loop and demodulation in one register-resident
// loop (the demodulation's independent instructions fill the latency of the gain recurrence)
template <bool EXACT, bool FM, int K0, int NK>
__device__ __forceinline__ void agc_emit_run(const AgcCoef &co, float g_thr, float fm_ref, float *myrow, float2 *myx, float &g, float &g2,
                                             float &y2p, AgcEmitState &e)
{
#pragma unroll
    for (int k = K0; k < K0 + NK; k++) {
        const float2 xv = myx[k];
        const float2 y = cf(__fmul_rn(xv.x, g), __fmul_rn(xv.y, g));          // gain BEFORE the sample
        agc_step<EXACT>(co, g, g2, y2p, myrow[1 + k]);                          // g: gain after the sample
        e.ex |= (g < g_thr ? 1u : 0u) << k;                                     // rssi = -20 log10(g) > threshold
        if (FM) {
            const float re = __fadd_rn(__fmul_rn(e.yp.x, y.x), __fmul_rn(e.yp.y, y.y));
            const float im = __fsub_rn(__fmul_rn(e.yp.x, y.y), __fmul_rn(e.yp.y, y.x));
            myrow[1 + k] = (EXACT ? atan2f(im, re) : be_atan2(im, re)) * fm_ref;
            // (one funnel shift per plane: the sign enters at bit 0, sample k of the block ends up at bit 31 - k; the word is
            // reversed once when it is stored)
            e.sr = be_shift_in_sign(e.sr, y.x);
            e.si = be_shift_in_sign(e.si, y.y);
        } else {
            myx[k] = y;
        }
        e.yp = y;
    }
}
// the same for a ragged block (chunk end): cnt < 32 samples
template <bool EXACT, bool FM>
__device__ __forceinline__ void agc_emit_ragged(const AgcCoef &co, float g_thr, float fm_ref, float *myrow, float2 *myx, float &g, float &g2,
                                             float &y2p, AgcEmitState &e, int cnt)
{
    for (int k = 0; k < cnt; k++) {
        const float2 xv = myx[k];
        const float2 y = cf(__fmul_rn(xv.x, g), __fmul_rn(xv.y, g));
        agc_step<EXACT>(co, g, g2, y2p, myrow[1 + k]);
        e.ex |= (g < g_thr ? 1u : 0u) << k;
        if (FM) {
            const float re = __fadd_rn(__fmul_rn(e.yp.x, y.x), __fmul_rn(e.yp.y, y.y));
            const float im = __fsub_rn(__fmul_rn(e.yp.x, y.y), __fmul_rn(e.yp.y, y.x));
            myrow[1 + k] = (EXACT ? atan2f(im, re) : be_atan2(im, re)) * fm_ref;
            e.sr |= ((unsigned)__float_as_int(y.x) >> 31) << k;
            e.si |= ((unsigned)__float_as_int(y.y) >> 31) << k;
        } else {
            myx[k] = y;
        }
        e.yp = y;
    }
}

template <bool EXACT, bool FM>
__global__ void __launch_bounds__(kAgcT, 4) k_agc_emit(const BackendParams p)
{
    CSDR_DYN_SMEM(smem_raw);
    float *sm = reinterpret_cast<float *>(smem_raw);                                         // [kAgcT][kAgcRow]
    float2 *s_x = reinterpret_cast<float2 *>(sm + kAgcT * kAgcRow);                          // [kAgcT][kAgcXRow] dc-blocked samples of the current block
    const int lane = blockIdx.y, tid = threadIdx.x, w = tid >> 5, l = tid & 31;
    const int seg0 = blockIdx.x * kAgcT, seg = seg0 + tid;
    const float *__restrict__ pw = p.pw + (long long)lane * p.pw_stride;
    const float2 *__restrict__ x = p.ydc + (long long)lane * p.ydc_stride;
    float *const of = be_out_f(p, lane);
    float2 *const oc = be_out_c(p, lane);
    unsigned *const exb = p.exbits + (long long)lane * p.nwords, *const sgr = p.sgnr + (long long)lane * p.nwords,
             *const sgi = p.sgni + (long long)lane * p.nwords;
    const float g_thr = p.g_thr, fm_ref = p.fm_ref;
    const bool skip_closed = p.gate != 0;
    const int b0 = seg * p.L;
    const bool live = seg < p.nseg;
    const int nsteps = (p.W + p.L) / kAgcB, wsteps = p.W / kAgcB;
    float g = 1.f, g2 = 1.f, y2p = 1.f, g30 = 1.f;
    AgcEmitState es; es.yp = cf(0.f, 0.f); es.ex = es.sr = es.si = 0u;
    bool started = false;
    // rows of this warp: thread r = 32 w + i, block start u_r = (seg0 + r) L - W + 32 s.  The 32 loads of the NEXT step
    // are issued before the recurrence of the current one runs, so their latency is never waited for.
    float nxt[32];
    const AgcCoef co{p.alpha, p.one_minus_alpha_f, p.neg_half_alpha};
    const int L = p.L, n = p.n;
    const int ubase = (seg0 + 32 * w) * L - p.W + l;           // int: n < 2^31
    // a CTA whose windows lie inside the chunk (all but the first and the last few) skips the bounds checks
    const bool interior = (seg0 * L - p.W >= 0) && ((seg0 + kAgcT) * L <= n);
    const bool full_rows = (seg0 + kAgcT) * L <= n;            // every row of every emitted block is complete
    auto fetch = [&](int s) {
        if (interior) {
            // (32-bit element offsets from one base pointer: one wide multiply-add per load)
            int off = ubase + s * kAgcB;
#pragma unroll
            for (int i = 0; i < 32; i++, off += L) nxt[i] = __ldg(pw + off);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const int u = ubase + i * L + s * kAgcB;
                nxt[i] = (u >= 0 && u < n) ? __ldg(pw + u) : 0.f;
            }
        }
    };
    // The samples a block's emission needs (rows of this warp) are copied into shared memory asynchronously, half a block
    // at a time and one block ahead: the first half of block s + 1 when the loop of step s is half way (its slots are free
    // then), the second half when it ends; four rows (4 x 128 bytes) per instruction.
    const bool x_async = full_rows && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    auto stage_x = [&](int eb, int half) {
        if (x_async) {
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int row = 32 * w + 4 * it + (l >> 3), c2 = 16 * half + 2 * (l & 7);
                cp_async16(s_x + row * kAgcXRow + c2, x + (long long)(seg0 + row) * L + eb + c2);
            }
        } else {
            for (int i = 0; i < 16; i++) {
                const int row = 32 * w + 2 * i + (l >> 4), c = 16 * half + (l & 15);
                const long long u = (long long)(seg0 + row) * L + eb + c;
                s_x[row * kAgcXRow + c] = (seg0 + row < p.nseg && u < n) ? x[u] : cf(0.f, 0.f);
            }
        }
        cp_async_commit();
    };
    float *myrow = sm + tid * kAgcRow;
    float2 *myx = s_x + tid * kAgcXRow;
    fetch(0);
    for (int s = 0; s < nsteps; s++) {
#pragma unroll
        for (int i = 0; i < 32; i++) sm[(32 * w + i) * kAgcRow + 1 + l] = nxt[i];
        const bool emit = s >= wsteps;
        const bool next_emit = s + 1 >= wsteps && s + 1 < nsteps;       // block s + 1 is emitted: its samples are staged during this step
        const int eb_next = (s + 1 - wsteps) * kAgcB;
        // first half of this block has landed (issued half a step ago; cf32 output: the whole block, see below)
        if (emit) { if (FM) cp_async_wait_group<1>(); else cp_async_wait_group<0>(); }
        __syncthreads();
        if (s + 1 < nsteps) fetch(s + 1);
        const int u0 = b0 - p.W + s * kAgcB;                    // time of this thread's myrow[1]
        const bool active = live && u0 >= 0 && u0 < n;
        const int cnt = active ? min(kAgcB, n - u0) : 0;
        if (active) {
            if (!started) {
                started = true;
                if (b0 - p.W <= 0) { const LaneState ls = p.lane[lane]; g = ls.g; y2p = ls.y2p; }   // time 0: exact
                else {
                    float e = 0.f;
#pragma unroll
                    for (int k = 0; k < 16; k++) e += myrow[1 + k];
                    agc_guess(e * (1.0f / 16.0f), g, y2p);
                }
                g2 = __fmul_rn(g, g);
            }
            if (s == wsteps) {
                SegState s0; s0.g = g; s0.y2p = y2p;
                p.seg_start[(long long)lane * p.nseg + seg] = s0;
                g2 = __fmul_rn(g, g);
                // ungated output of the sample before the segment: y[b0-1] = y_dc[b0-1] * (gain before it = gain after
                // b0-2, kept from the warm-up).  Segment 0 continues the previous call (lane state).
                if (b0 == 0) { const LaneState ls = p.lane[lane]; es.yp = cf(ls.fm_re, ls.fm_im); }
                else { const float2 xv = x[b0 - 1]; es.yp = cf(__fmul_rn(xv.x, g30), __fmul_rn(xv.y, g30)); }
            }
        }
        // ---- first half of the block
        if (emit) {
            es.ex = es.sr = es.si = 0u;
            if (cnt == kAgcB) agc_emit_run<EXACT, FM, 0, 16>(co, g_thr, fm_ref, myrow, myx, g, g2, y2p, es);
            else if (cnt > 0) agc_emit_ragged<EXACT, FM>(co, g_thr, fm_ref, myrow, myx, g, g2, y2p, es, cnt);
        } else if (cnt == kAgcB) {
#pragma unroll
            for (int k = 0; k < 16; k++) agc_step<EXACT>(co, g, g2, y2p, myrow[1 + k]);
        } else {
            for (int k = 0; k < cnt; k++) agc_step<EXACT>(co, g, g2, y2p, myrow[1 + k]);
        }
        // (cf32 output is written over the samples in place and copied out at the end of the step: the next block is staged
        // behind the copy-out then)
        __syncwarp();
        if (FM) {
            if (next_emit) stage_x(eb_next, 0); else cp_async_commit();
            cp_async_wait_group<1>();                                   // second half of this block has landed
        }
        __syncwarp();
        // ---- second half
        if (emit) {
            if (cnt == kAgcB) agc_emit_run<EXACT, FM, 16, 16>(co, g_thr, fm_ref, myrow, myx, g, g2, y2p, es);
        } else if (cnt == kAgcB) {
#pragma unroll
            for (int k = 16; k < 32; k++) { if (k == 31) g30 = g; agc_step<EXACT>(co, g, g2, y2p, myrow[1 + k]); }
        }
        __syncwarp();
        if (FM) { if (next_emit) stage_x(eb_next, 1); else cp_async_commit(); }
        if (emit) {
            // bit words of this thread's block; the threshold word also goes into the tile (row slot 0) for the copy-out
            if (cnt > 0) {
                const int wd = (b0 + (s - wsteps) * kAgcB) >> 5;
                exb[wd] = es.ex;
                // full blocks collected the sign bits most recent sample first (agc_emit_run), ragged ones in place
                if (FM) { sgr[wd] = cnt == kAgcB ? __brev(es.sr) : es.sr; sgi[wd] = cnt == kAgcB ? __brev(es.si) : es.si; }
                if (u0 + cnt >= min(b0 + L, n)) p.seg_ylast[(long long)lane * p.nseg + seg] = es.yp;     // the segment's last sample
            }
            myrow[0] = __uint_as_float(cnt > 0 ? es.ex : 0u);
            __syncthreads();
            // ---- copy-out: warp w writes rows 32 w .. 32 w + 31, lane = sample (coalesced).  The squelch can only be open
            // on a sample that exceeds the threshold (every path into SIGNALHI takes the "exceeded" branch of the state
            // machine), so a word without such a sample is closed throughout and left to the gate pass.
            const int eb = (s - wsteps) * kAgcB;
            const int ub0 = (seg0 + 32 * w) * L + eb;
            const float *srow = sm + (32 * w) * kAgcRow;
            // (a word without a threshold-exceeding sample is closed throughout -- the squelch can only be open on a sample
            // that exceeds the threshold: every path into SIGNALHI takes the "exceeded" branch of the state machine -- so it
            // is written as zeros right here and the gate pass does not have to touch it)
            if (full_rows) {
                if (FM) {
                    float *dst = of + ub0 + l;
#pragma unroll 8
                    for (int i = 0; i < 32; i++, dst += L, srow += kAgcRow)
                        *dst = (skip_closed && __float_as_uint(srow[0]) == 0u) ? 0.f : srow[1 + l];
                } else {
                    float2 *dst = oc + ub0 + l;
                    const float2 *xr = s_x + (32 * w) * kAgcXRow + l;
#pragma unroll 8
                    for (int i = 0; i < 32; i++, dst += L, srow += kAgcRow, xr += kAgcXRow)
                        *dst = (skip_closed && __float_as_uint(srow[0]) == 0u) ? cf(0.f, 0.f) : *xr;
                }
            } else {
                int ub = ub0;
                for (int i = 0; i < 32; i++, ub += L) {
                    const int row = 32 * w + i;
                    if (seg0 + row >= p.nseg || ub >= n) break;                                       // warp-uniform
                    const bool zero = skip_closed && __float_as_uint(sm[row * kAgcRow]) == 0u;        // warp-uniform
                    if (ub + l < n) {
                        if (FM) of[ub + l] = zero ? 0.f : sm[row * kAgcRow + 1 + l];
                        else    oc[ub + l] = zero ? cf(0.f, 0.f) : s_x[row * kAgcXRow + l];
                    }
                }
            }
        }
        if (!FM && next_emit) { __syncwarp(); stage_x(eb_next, 0); stage_x(eb_next, 1); }
        __syncthreads();
    }
    if (live) {
        SegState s1; s1.g = g; s1.y2p = y2p;
        p.seg_end[(long long)lane * p.nseg + seg] = s1;
    }
}

// no AGC: the first-sample state is saved by a launch of its own
__global__ void k_be_first(const BackendParams p)
{
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane < p.nlanes) be_save_first(p, lane);
}

// ------------------------------------------------------------------------------------------ emission without AGC
// One sample per thread, one 32-sample word per warp and iteration; gain 1, no bits.
constexpr int kEmitRun = 16;     // consecutive 32-sample words per warp: the previous sample comes from the neighbour lane
template <bool FM, bool EXACT>
__global__ void __launch_bounds__(256) k_be_emit(const BackendParams p)
{
    const int lane = blockIdx.y, l = threadIdx.x & 31;
    const float2 *__restrict__ x = p.ydc + (long long)lane * p.ydc_stride;
    const BeEmitCtx ctx = be_emit_ctx(p, lane);
    const int n = p.n, nwords = p.nwords;
    const int w0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kEmitRun, w1 = min(w0 + kEmitRun, nwords);
    if (w0 >= nwords) return;                                    // whole warps leave together
    float2 ylast = (w0 == 0) ? p.y_first[lane] : x[w0 * 32 - 1];
    for (int w = w0; w < w1; w++) {
        const int i = w * 32 + l;
        const bool in = i < n;
        const float2 xv = in ? x[i] : cf(0.f, 0.f);
        const float2 y = be_emit_word<false, FM, EXACT, false>(ctx, w * 32, xv, 1.f, 1.f, ylast);
        if (in && i == n - 1) { p.lane[lane].fm_re = y.x; p.lane[lane].fm_im = y.y; }      // (read by k_be_first, an earlier launch)
        ylast.x = __shfl_sync(0xffffffffu, y.x, 31); ylast.y = __shfl_sync(0xffffffffu, y.y, 31);
    }
}

// ------------------------------------------------------------------------------------------ squelch FSM on the bits
struct FsmState { int mode; unsigned timer; };
__device__ __forceinline__ bool fsm_same(const FsmState &a, const FsmState &b)
{
    return a.mode == b.mode && (a.mode != SQ_SIGNALLO || a.timer == b.timer);
}

// run the FSM over bits [i0, i1) of a lane (bits[w - wbase] = word w); when gate != nullptr write one gate bit
// (mode == SIGNALHI) per sample
__device__ __forceinline__ void fsm_run(const BackendParams &p, const unsigned *bits, int wbase, unsigned *gate, FsmState &s,
                                        int i0, int i1)
{
    int i = i0;
    while (i < i1) {
        const int w = i >> 5, lo = i & 31, cnt = min(32 - lo, i1 - i);
        const unsigned full = (cnt == 32) ? 0xffffffffu : ((1u << cnt) - 1u);
        const unsigned word = (bits[w - wbase] >> lo) & full;
        unsigned gw = 0;
        if (word == full && s.mode == SQ_SIGNALHI) gw = full;                              // stays open
        else if (word == 0 && s.mode == SQ_ENABLED) gw = 0;                                // stays closed
        else if (word == 0 && s.mode == SQ_SIGNALLO && s.timer > (unsigned)cnt) s.timer -= (unsigned)cnt;
        else {
            for (int b = 0; b < cnt; b++) {
                fsm_step(s.mode, s.timer, (word >> b) & 1u, p.timeout);
                gw |= (s.mode == SQ_SIGNALHI ? 1u : 0u) << b;
            }
        }
        if (gate) {
            // words are owned by one segment except for a ragged chunk end; lo != 0 only happens at i0 of a replay
            if (lo == 0) gate[w] = gw; else gate[w] = (gate[w] & ((1u << lo) - 1u)) | (gw << lo);
        }
        i += cnt;
    }
}

// ------------------------------------------------------------------------------------------ k_be_finish
// Everything behind the fused gain/emission kernel, in ONE cooperative launch (all CTAs co-resident, software grid
// barrier between phases):
//   A  verify the gain speculation (segments whose start state does not continue their predecessor's end state are
//      listed) and resolve the squelch FSM speculatively on the threshold bits (every segment replays timeout + 8 bits)
//   -- only if the list is not empty (digital silence, an un-damped stretch of the loop): refine the listed segments in
//      parallel from their predecessors' end states (two rounds), repair what is left in stream order, redo A's FSM pass
//   B  verify the FSM speculation; misses (a measure-zero coincidence of the time-out with a threshold crossing) are
//      repaired in stream order
//   C  apply the gate (Liquid.chs:700-704) to the output, store the lanes' states, reset the counters
// Under the test-only CPU emulation (CTAs run one after the other) the phases are launched one by one instead.
constexpr int kFinT = 256;
constexpr int kFinWords = 8320;                 // shared staging: threshold words of 256 segments + replay, or a segment's gains
enum { FIN_A = 0, FIN_R0, FIN_V1, FIN_R1, FIN_V2, FIN_FIX, FIN_A2, FIN_B, FIN_BFIX, FIN_C, FIN_NPHASE };

__device__ __forceinline__ void be_grid_barrier(unsigned *bar, unsigned nblocks)
{
    __syncthreads();
#ifndef CSDR_EMU
    if (threadIdx.x == 0) {
        volatile unsigned *vb = bar;
        const unsigned gen = vb[1];
        __threadfence();
        if (atomicAdd(&bar[0], 1u) == nblocks - 1) { vb[0] = 0; __threadfence(); atomicAdd(&bar[1], 1u); }
        else { while (vb[1] == gen) {} }
        __threadfence();
    }
    __syncthreads();
#else
    (void)bar; (void)nblocks;
#endif
}

// Re-run one segment from `st` (the predecessor's end state) and emit it again; CTA-cooperative: thread 0 runs the
// recurrence into shared memory, then every warp emits whole words.  s_g: L + 1 floats.
template <bool EXACT>
__device__ void be_redo_segment(const BackendParams &p, int lane, int seg, SegState st, float *s_g)
{
    const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n), len = b1 - b0;
    const float *__restrict__ pw = p.pw + (long long)lane * p.pw_stride;
    const float2 *__restrict__ x = p.ydc + (long long)lane * p.ydc_stride;
    const long long t = (long long)lane * p.nseg + seg;
    __syncthreads();
    if (threadIdx.x == 0) {
        const AgcCoef co{p.alpha, p.one_minus_alpha_f, p.neg_half_alpha};
        float g = st.g, y2p = st.y2p, g2 = __fmul_rn(g, g);
        p.seg_start[t] = st;
        s_g[0] = g;
        for (int i = 0; i < len; i++) { agc_step<EXACT>(co, g, g2, y2p, pw[b0 + i]); s_g[i + 1] = g; }
        SegState s1; s1.g = g; s1.y2p = y2p;
        p.seg_end[t] = s1;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, l = threadIdx.x & 31;
    const BeEmitCtx ctx = be_emit_ctx(p, lane);
    const float2 yprev_seg = (seg == 0) ? p.y_first[lane] : p.seg_ylast[t - 1];
    for (int wd = warp; wd * 32 < len; wd += nwarps) {
        const int idx = wd * 32 + l, u = b0 + idx;
        const bool in = idx < len;
        const float2 xv = in ? x[u] : cf(0.f, 0.f);
        const float g0 = in ? s_g[idx] : 1.f, ga = in ? s_g[idx + 1] : 1.f;
        // output of the sample before the word: from this segment's own gains, or the predecessor's last sample
        float2 yp0 = yprev_seg;
        if (wd > 0) { const float2 xp = x[b0 + wd * 32 - 1]; const float gp = s_g[wd * 32 - 1]; yp0 = cf(__fmul_rn(xp.x, gp), __fmul_rn(xp.y, gp)); }
        float2 y;
        if (p.demod == 1) y = be_emit_word<true, true, EXACT, false>(ctx, b0 + wd * 32, xv, g0, ga, yp0);
        else              y = be_emit_word<true, false, EXACT, false>(ctx, b0 + wd * 32, xv, g0, ga, yp0);
        if (idx == len - 1) p.seg_ylast[t] = y;
    }
    __syncthreads();
}

template <bool EXACT>
__global__ void __launch_bounds__(kFinT, 2) k_be_finish(const BackendParams p, int phase_lo, int phase_hi)
{
    __shared__ unsigned s_buf[kFinWords];
    __shared__ unsigned s_next;
    __shared__ int s_cur;
    const int tid = threadIdx.x, nthr = gridDim.x * blockDim.x, gtid = blockIdx.x * blockDim.x + tid;
    const long long nsegs = (long long)p.nlanes * p.nseg;
    const int blocks_per_lane = (p.nseg + kFinT - 1) / kFinT;
    const bool stage_ok = (long long)(p.FW + kFinT) * p.L / 32 <= kFinWords;

    bool pending = false;                        // a phase has run since the last grid barrier
    for (int ph = phase_lo; ph <= phase_hi; ph++) {
        if (pending) { be_grid_barrier(p.barrier, gridDim.x); pending = false; }
        // phases that only exist to repair a failed speculation are skipped (same decision in every CTA: the counters
        // are complete, no phase has run since the last barrier)
        const unsigned cnt0 = p.bad_count[0];
        if ((ph >= FIN_R0 && ph <= FIN_A2) && cnt0 == 0) continue;
        if (ph == FIN_R1 && p.bad_count[1] == 0) continue;
        if (ph == FIN_BFIX && p.bad_count[2] == 0) continue;
        pending = true;

        if (ph == FIN_A || ph == FIN_V1 || ph == FIN_V2) {
            // ---- gain speculation: start state of every segment against its predecessor's end state
            const int round = ph == FIN_A ? 0 : ph == FIN_V1 ? 1 : 2;
            for (long long t = gtid; t < nsegs; t += nthr) {
                const int seg = (int)(t % p.nseg);
                if (seg == 0 || be_match(p.seg_start[t], p.seg_end[t - 1])) continue;
                if (round < 2) {
                    const unsigned idx = atomicAdd(p.bad_count + round, 1u);
                    if (idx < p.bad_cap) p.bad_list[(long long)round * p.bad_cap + idx] = (unsigned)t;
                } else {
                    atomicMin(&p.first_bad[2 * (int)(t / p.nseg)], (unsigned)seg);
                }
            }
        }
        if (ph == FIN_A) {
            // what the chunk's first sample needs from the previous call, before the lane states are overwritten
            for (int lane = gtid; lane < p.nlanes; lane += nthr) {
                be_save_first(p, lane);
                p.prev_gate[lane] = (p.lane[lane].mode == SQ_SIGNALHI) ? 1u : 0u;
            }
        }
        if (ph == FIN_R0 || ph == FIN_R1) {
            // ---- second chance, in parallel: a start state that is slightly off (the warm-up met an un-damped stretch
            // of the loop) is replaced by the predecessor's END state, which is accurate because the predecessor's own L
            // samples damped its error; the segment is re-run and re-emitted from there, one CTA per segment.
            const int round = ph == FIN_R0 ? 0 : 1;
            const unsigned count = min(p.bad_count[round], p.bad_cap);
            for (unsigned idx = blockIdx.x; idx < count; idx += gridDim.x) {
                const long long t = p.bad_list[(long long)round * p.bad_cap + idx];
                const int lane = (int)(t / p.nseg), seg = (int)(t - (long long)lane * p.nseg);
                be_redo_segment<EXACT>(p, lane, seg, p.seg_end[t - 1], reinterpret_cast<float *>(s_buf));
            }
            if (blockIdx.x == 0 && tid == 0) atomicAdd(p.fixups + 2, (unsigned long long)count);
        }
        if (ph == FIN_FIX) {
            // ---- last resort: re-run the remaining misses of a lane in stream order (a gain frozen by digital silence)
            for (int lane = blockIdx.x; lane < p.nlanes; lane += gridDim.x) {
                SegState *E = p.seg_end + (long long)lane * p.nseg;
                const SegState *S0 = p.seg_start + (long long)lane * p.nseg;
                bool first = true;
                if (tid == 0) s_cur = 1;
                __syncthreads();
                while (true) {
                    if (tid == 0) s_next = first ? p.first_bad[2 * lane] : 0xffffffffu;
                    __syncthreads();
                    if (!first) {
                        // after a repair: parallel search for the next segment >= s_cur that does not continue its predecessor
                        const int cur = s_cur;
                        for (int j = cur + tid; j < p.nseg; j += blockDim.x)
                            if (!be_match(S0[j], E[j - 1])) { atomicMin(&s_next, (unsigned)j); break; }
                        __syncthreads();
                    }
                    first = false;
                    const unsigned nxt = s_next;
                    if (nxt == 0xffffffffu) break;
                    int seg = (int)nxt;
                    unsigned long long redone = 0;
                    while (seg < p.nseg) {
                        __syncthreads();
                        const SegState pe = E[seg - 1];
                        if (be_match(S0[seg], pe)) break;       // uniform: every thread reads the same two states
                        be_redo_segment<EXACT>(p, lane, seg, pe, reinterpret_cast<float *>(s_buf));
                        redone++;
                        seg++;      // the successor is re-checked against the new end state on the next iteration
                    }
                    __syncthreads();
                    if (tid == 0) { s_cur = seg; atomicAdd(p.fixups, redone); }
                    __syncthreads();
                }
                __syncthreads();
            }
        }
        if (ph == FIN_A || ph == FIN_A2) {
            // ---- squelch FSM, speculative: 256 consecutive segments of a lane per CTA; their threshold words and the
            // FW segments of replay in front of them are staged in shared memory with coalesced loads
            for (int item = blockIdx.x; item < p.nlanes * blocks_per_lane; item += gridDim.x) {
                const int lane = item / blocks_per_lane, sb = (item - lane * blocks_per_lane) * kFinT;
                const unsigned *bits = p.exbits + (long long)lane * p.nwords;
                unsigned *gate = p.gatebits + (long long)lane * p.nwords;
                const int wlo = max(0, (sb - p.FW) * (p.L / 32)), whi = min(p.nwords, (sb + kFinT) * (p.L / 32));
                __syncthreads();
                if (stage_ok) for (int wd = wlo + tid; wd < whi; wd += blockDim.x) s_buf[wd - wlo] = bits[wd];
                __syncthreads();
                const int seg = sb + tid;
                if (seg < p.nseg) {
                    const long long t = (long long)lane * p.nseg + seg;
                    const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
                    FsmState s;
                    int w0 = b0 - p.FW * p.L;
                    if (w0 <= 0) { w0 = 0; s.mode = p.lane[lane].mode; s.timer = p.lane[lane].timer; }   // exact
                    else { s.mode = SQ_ENABLED; s.timer = 0; }
                    const unsigned *src = stage_ok ? s_buf : bits;
                    const int wbase = stage_ok ? wlo : 0;
                    fsm_run(p, src, wbase, nullptr, s, w0, b0);
                    p.fsm_start[t] = s;
                    fsm_run(p, src, wbase, gate, s, b0, b1);
                    p.fsm_end[t] = s;
                }
            }
        }
        if (ph == FIN_B) {
            for (long long t = gtid; t < nsegs; t += nthr) {
                const int seg = (int)(t % p.nseg);
                if (seg > 0 && !fsm_same(p.fsm_start[t], p.fsm_end[t - 1])) {
                    atomicMin(&p.first_bad[2 * (int)(t / p.nseg) + 1], (unsigned)seg);
                    atomicAdd(p.bad_count + 2, 1u);
                }
            }
        }
        if (ph == FIN_BFIX) {
            for (int lane = blockIdx.x; lane < p.nlanes; lane += gridDim.x) {
                FsmState *E = p.fsm_end + (long long)lane * p.nseg;
                const FsmState *S0 = p.fsm_start + (long long)lane * p.nseg;
                const unsigned *bits = p.exbits + (long long)lane * p.nwords;
                unsigned *gate = p.gatebits + (long long)lane * p.nwords;
                bool first = true;
                if (tid == 0) s_cur = 1;
                __syncthreads();
                while (true) {
                    if (tid == 0) s_next = first ? p.first_bad[2 * lane + 1] : 0xffffffffu;
                    __syncthreads();
                    if (!first) {
                        const int cur = s_cur;
                        for (int j = cur + tid; j < p.nseg; j += blockDim.x)
                            if (!fsm_same(S0[j], E[j - 1])) { atomicMin(&s_next, (unsigned)j); break; }
                        __syncthreads();
                    }
                    first = false;
                    const unsigned nxt = s_next;
                    if (nxt == 0xffffffffu) break;
                    if (tid == 0) {
                        int seg = (int)nxt;
                        unsigned long long redone = 0;
                        while (seg < p.nseg) {
                            FsmState s = E[seg - 1];
                            if (fsm_same(S0[seg], s)) break;
                            const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
                            fsm_run(p, bits, 0, gate, s, b0, b1);
                            E[seg] = s;
                            redone++;
                            seg++;
                        }
                        s_cur = seg;
                        atomicAdd(p.fixups + 1, redone);
                    }
                    __syncthreads();
                }
                __syncthreads();
            }
        }
        if (ph == FIN_C) {
            // ---- gate.  cf32 output: closed samples become 0+0j (Liquid.chs:704).  Discriminator output: m[i] =
            // arg(conj(r'[i-1]) r'[i]) with r' the GATED samples; when either neighbour is closed the product is a signed
            // zero whose argument is 0 or +-pi exactly as in the sequential code, so those values are recomputed here
            // from the sign bits of the ungated samples.  One warp per run of 32 words, lane = sample.
            if (p.gate) {
                const int l = tid & 31, warp = gtid >> 5, nwarp = nthr >> 5;
                const int runs_per_lane = (p.nwords + 31) / 32;
                // (the words of the NEXT run are fetched before the current one is processed: the loop is a chain of
                // dependent global loads otherwise)
                struct Run { unsigned gw, fullw, gtop, srw, siw, srt, sit, exw; int lane, wd0; bool have; };
                auto load_run = [&](long long item) {
                    Run r{};
                    r.lane = (int)(item / runs_per_lane); r.wd0 = (int)(item - (long long)r.lane * runs_per_lane) * 32;
                    const unsigned *gate = p.gatebits + (long long)r.lane * p.nwords;
                    const int wd = r.wd0 + l;
                    r.have = wd < p.nwords;
                    const int cntw = r.have ? min(32, p.n - wd * 32) : 0;
                    r.fullw = (cntw == 32) ? 0xffffffffu : ((1u << cntw) - 1u);
                    r.gw = r.have ? (gate[wd] & r.fullw) : 0u;
                    r.gtop = r.have ? (wd ? (gate[wd - 1] >> 31) : p.prev_gate[r.lane]) : 0u;
                    r.exw = r.have ? p.exbits[(long long)r.lane * p.nwords + wd] : 0u;
                    if (p.demod == 1 && r.have) {
                        const unsigned *sr = p.sgnr + (long long)r.lane * p.nwords, *si = p.sgni + (long long)r.lane * p.nwords;
                        r.srw = sr[wd]; r.siw = si[wd];
                        r.srt = wd ? (sr[wd - 1] >> 31) : (p.prev_sign[r.lane] & 1u);
                        r.sit = wd ? (si[wd - 1] >> 31) : ((p.prev_sign[r.lane] >> 1) & 1u);
                    }
                    return r;
                };
                const long long nitems = (long long)p.nlanes * runs_per_lane;
                Run nxt_run{};
                if (warp < nitems) nxt_run = load_run(warp);
                for (long long item = warp; item < nitems; item += nwarp) {
                    const Run cur = nxt_run;
                    if (item + nwarp < nitems) nxt_run = load_run(item + nwarp);
                    const int lane = cur.lane, wd0 = cur.wd0;
                    const bool have = cur.have;
                    const unsigned fullw = cur.fullw, gw = cur.gw, gtop = cur.gtop, srw = cur.srw, siw = cur.siw, srt = cur.srt, sit = cur.sit;
                    const int cntw = (fullw == 0xffffffffu) ? 32 : 0;      // only "whole word" matters below
                    // words that need no write (open throughout, predecessor open) are skipped; words that are closed
                    // throughout (the bulk of a squelched stream) take a short path: one coalesced store of zeros
                    const unsigned gprevw = (gw << 1) | gtop;
                    // (words without a threshold-exceeding sample were written as zeros by the emission already)
                    const bool zeroed = cur.exw == 0u && gw == 0u && (p.demod != 1 || gtop == 0u);
                    const bool need = have && !zeroed && (p.demod == 1 ? ((gw & gprevw & fullw) != fullw) : (gw != fullw));
                    const bool closed = have && cntw == 32 && gw == 0 && (p.demod != 1 || gtop == 0);
                    const unsigned m_closed = __ballot_sync(0xffffffffu, closed);
                    for (unsigned m = __ballot_sync(0xffffffffu, need); m; m &= m - 1) {
                        const int j = __ffs(m) - 1;
                        const int i = (wd0 + j) * 32 + l;
                        if ((m_closed >> j) & 1u) {
                            if (p.demod == 1) be_out_f(p, lane)[i] = 0.f;            // arg(+0 + j0) = 0
                            else              be_out_c(p, lane)[i] = cf(0.f, 0.f);
                            continue;
                        }
                        const unsigned g = __shfl_sync(0xffffffffu, gw, j), full = __shfl_sync(0xffffffffu, fullw, j);
                        const bool in = (full >> l) & 1u;
                        if (p.demod != 1) {
                            if (in && !((g >> l) & 1u)) be_out_c(p, lane)[i] = cf(0.f, 0.f);
                            continue;
                        }
                        const unsigned gprev = __shfl_sync(0xffffffffu, gprevw, j);
                        float *of = be_out_f(p, lane);
                        const unsigned r = __shfl_sync(0xffffffffu, srw, j), im = __shfl_sync(0xffffffffu, siw, j);
                        const unsigned rprev = (r << 1) | __shfl_sync(0xffffffffu, srt, j), iprev = (im << 1) | __shfl_sync(0xffffffffu, sit, j);
                        const bool open = (g >> l) & 1u, popen = (gprev >> l) & 1u;
                        if (!in || (open && popen)) continue;
                        // unit-magnitude stand-ins carry the signs; a closed sample is +0+0j
                        const float yr = open ? (((r >> l) & 1u) ? -1.f : 1.f) : 0.f;
                        const float yi = open ? (((im >> l) & 1u) ? -1.f : 1.f) : 0.f;
                        const float fr = popen ? (((rprev >> l) & 1u) ? -1.f : 1.f) : 0.f;
                        const float fi = popen ? (((iprev >> l) & 1u) ? -1.f : 1.f) : 0.f;
                        const float re = __fadd_rn(__fmul_rn(fr, yr), __fmul_rn(fi, yi));
                        const float ii = __fsub_rn(__fmul_rn(fr, yi), __fmul_rn(fi, yr));
                        of[i] = atan2f(ii, re) * p.fm_ref;
                    }
                }
            }
            // lane states for the next call (nobody reads the old ones any more: the barrier in front of this phase)
            for (int lane = gtid; lane < p.nlanes; lane += nthr) {
                const SegState e = p.seg_end[(long long)lane * p.nseg + p.nseg - 1];
                const FsmState f = p.fsm_end[(long long)lane * p.nseg + p.nseg - 1];
                const float2 y = p.seg_ylast[(long long)lane * p.nseg + p.nseg - 1];      // ungated output of the chunk's last sample
                LaneState &ls = p.lane[lane];
                ls.g = e.g; ls.y2p = e.y2p; ls.mode = f.mode; ls.timer = f.timer; ls.fm_re = y.x; ls.fm_im = y.y;
                p.first_bad[2 * lane] = 0xffffffffu; p.first_bad[2 * lane + 1] = 0xffffffffu;
            }
        }
    }
    // counters for the next call: by the last CTA to get here (every CTA has read them by then)
    if (phase_hi == FIN_C) {
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            if (atomicAdd(&p.barrier[2], 1u) == gridDim.x - 1) { p.bad_count[0] = 0; p.bad_count[1] = 0; p.bad_count[2] = 0; p.barrier[2] = 0; }
        }
    }
}

// stand-alone freqdem (freqdem_demodulate_block): fully parallel, r' = previous input sample
__global__ void k_freqdem(const float2 *__restrict__ r, float *__restrict__ m, long long n, float2 prev, float ref)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float2 a = (i == 0) ? prev : r[i - 1];
        float2 b = r[i];
        float re = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
        float im = __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x));
        m[i] = atan2f(im, re) * ref;
    }
}

// mix: out[t] = sum over lanes (channel order, left fold -- Trans.hs:119-122)
__global__ void k_lane_sum(const float *__restrict__ in, long long lane_stride, int nlanes, float *__restrict__ out,
                           long long n /* floats */)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        float acc = in[i];
        // same left fold, loads issued 32 lanes ahead of the adds (few output columns, many lanes: the loop is bound
        // by how many loads are in flight)
        int c = 1;
        for (; c + 32 <= nlanes; c += 32) {
            float v[32];
#pragma unroll
            for (int k = 0; k < 32; k++) v[k] = in[(long long)(c + k) * lane_stride + i];
#pragma unroll
            for (int k = 0; k < 32; k++) acc = __fadd_rn(acc, v[k]);
        }
        for (; c < nlanes; c++) acc = __fadd_rn(acc, in[(long long)c * lane_stride + i]);
        out[i] = acc;
    }
}

// The same fold over the back end's own gated output, without reading what the gate has closed: a 32-sample word whose gate
// bits are all zero (and, for the discriminator, whose predecessor sample is closed too: next to an open sample a closed one
// is arg(+-0 +- j0) = 0 or +-pi, not 0) holds exact zeros, so the fold skips it (x + 0 = x) -- a wide-band
// input with few occupied channels is mostly closed words (config 4: 64 of 1024 channels carry a signal).  fpw = floats per
// sample (1 discriminator, 2 cf32); all threads of a warp look at the same gate word (one broadcast load).
__global__ void __launch_bounds__(256) k_lane_sum_gated(const float *__restrict__ in, long long lane_stride, int nlanes, float *__restrict__ out,
                                                        long long nsamp, const unsigned *__restrict__ gate, int nwords,
                                                        const unsigned *__restrict__ prev_gate, int fpw, int need_prev)
{
    // one warp per 32-sample word: lane k inspects the gate word of channel c0 + k, the ballot lists the channels that have
    // to be read, and those are added in channel order (four loads in flight)
    const int l = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long wd = warp; wd < nwords; wd += nwarp) {
        const long long sidx = wd * 32 + l;
        const bool valid = sidx < nsamp;
        float a0 = 0.f, a1 = 0.f;
        for (int c0 = 0; c0 < nlanes; c0 += 32) {
            const int c = c0 + l;
            unsigned g = 0u;
            if (c < nlanes) {
                const unsigned *gl = gate + (long long)c * nwords;
                g = gl[wd];
                if (need_prev) g |= wd ? (gl[wd - 1] >> 31) : prev_gate[c];
            }
            unsigned m = __ballot_sync(0xffffffffu, g != 0u);
            while (m) {
                int ch[4]; int cnt = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) if (m) { ch[k] = c0 + __ffs(m) - 1; m &= m - 1; cnt = k + 1; } else ch[k] = 0;
                if (fpw == 2) {
                    float2 v[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = (k < cnt && valid) ? reinterpret_cast<const float2 *>(in + (long long)ch[k] * lane_stride)[sidx] : cf(0.f, 0.f);
#pragma unroll
                    for (int k = 0; k < 4; k++) if (k < cnt) { a0 = __fadd_rn(a0, v[k].x); a1 = __fadd_rn(a1, v[k].y); }
                } else {
                    float v[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) v[k] = (k < cnt && valid) ? in[(long long)ch[k] * lane_stride + sidx] : 0.f;
#pragma unroll
                    for (int k = 0; k < 4; k++) if (k < cnt) a0 = __fadd_rn(a0, v[k]);
                }
            }
        }
        if (valid) { if (fpw == 2) reinterpret_cast<float2 *>(out)[sidx] = cf(a0, a1); else out[sidx] = a0; }
    }
}

// ------------------------------------------------------------------------------------------ launch sequences
// `launch(kernel, grid, block, smem_bytes, args...)` is supplied by the caller (CUDA stream launcher in
// csdr_b200.cu, thread emulator in the CPU-only tests) so that both run the same sequence with the same grids.
template <class Launch>
inline void be_launch_prep(Launch &launch, const DcParams &d)
{
    const dim3 grid((unsigned)std::max(1, std::min(65535, (d.ngrp + 63) / 64)), d.nlanes), block(256);   // ~8 groups per warp
    if (d.G == 128)     launch(k_be_prep<4>, grid, block, 0, d);
    else if (d.G == 64) launch(k_be_prep<2>, grid, block, 0, d);
    else                launch(k_be_prep<1>, grid, block, 0, d);
}
// dc blocker in one pass (states, dc-blocked samples, power): persistent CTAs, `max_ctas` of them
template <class Launch>
inline void be_launch_dc(Launch &launch, const DcParams &d, int max_ctas)
{
    const dim3 grid((unsigned)std::max(1, std::min(max_ctas, d.nblk * d.nlanes))), block(32 * kDcWarps);
    if (d.G == 128)     launch(k_dc_scan<4>, grid, block, kDcSmem, d);
    else if (d.G == 64) launch(k_dc_scan<2>, grid, block, kDcSmem, d);
    else                launch(k_dc_scan<1>, grid, block, kDcSmem, d);
}

// everything after the dc/power pass: gain loop + emission, then verification / squelch FSM / gate in one cooperative launch
// (`launch.coop(kernel, block, params, phase_lo, phase_hi)` sizes the grid to what is co-resident; the CPU emulation runs
// the phases one by one)
template <class Launch>
inline void be_launch(Launch &launch, const BackendParams &b)
{
    if (b.has_agc) {
        const dim3 grid((b.nseg + kAgcT - 1) / kAgcT, b.nlanes), block(kAgcT);      // L, W: multiples of 32
        if (b.exact_math) { if (b.demod == 1) launch(k_agc_emit<true, true>, grid, block, kAgcSmem, b); else launch(k_agc_emit<true, false>, grid, block, kAgcSmem, b); }
        else              { if (b.demod == 1) launch(k_agc_emit<false, true>, grid, block, kAgcSmem, b); else launch(k_agc_emit<false, false>, grid, block, kAgcSmem, b); }
        launch.debug_after_verify(b);
        if (b.exact_math) launch.coop(k_be_finish<true>, dim3(kFinT), b, (int)FIN_A, (int)FIN_C);
        else              launch.coop(k_be_finish<false>, dim3(kFinT), b, (int)FIN_A, (int)FIN_C);
    } else {
        launch(k_be_first, dim3((b.nlanes + 127) / 128), dim3(128), 0, b);
        const dim3 grid((unsigned)std::max(1, (b.nwords + 8 * kEmitRun - 1) / (8 * kEmitRun)), b.nlanes), block(256);   // kEmitRun words per warp
        if (b.exact_math) { if (b.demod == 1) launch(k_be_emit<true, true>, grid, block, 0, b); else launch(k_be_emit<false, true>, grid, block, 0, b); }
        else              { if (b.demod == 1) launch(k_be_emit<true, false>, grid, block, 0, b); else launch(k_be_emit<false, false>, grid, block, 0, b); }
    }
}

}  // namespace csdr
