// backend.cuh -- the sequential-state blocks that run at the decimated / per-channel rate:
//     iirfilt_crcf dc blocker   (Liquid.chs:575-589, liquid iirfilt.c "normal" form: v0 = x - a1 v1, y = v0 - v1)
//     agc_crcf + squelch gate   (Liquid.chs:693-717, liquid agc.c; Haskell zeroes y unless status == SIGNALHI)
//     freqdem                   (Liquid.chs:324-334, liquid freqdem.c)
// for `nlanes` independent sample sequences (streams or channelizer channels) of n samples each.
//
// Parallelisation over TIME (the reference is one sequential loop per stream):
//   * dc blocker: a linear recurrence.  k_dc_local reduces every G-sample group to its zero-state response and
//     scans the groups of a CTA, k_dc_carry chains the CTAs (fp64): the exact filter state at every group boundary
//     is then one multiply-add away, and consumers restart the float32 recurrence from those states.
//   * AGC + squelch: the gain loop is contractive, so L-sample segments are run speculatively after a W-sample
//     warm-up and verified against their predecessors; the squelch state machine never feeds back into the gain
//     and is resolved EXACTLY afterwards on one threshold bit per sample; the gate is applied last as a mask.
//     Details at "agc + fm" below.  The result equals the sequential loop to within the stated tolerance in all
//     cases (gate positions exactly); only the speed depends on the signal.
#pragma once
#include "platform.cuh"

namespace csdr {

enum { SQ_UNKNOWN = 0, SQ_ENABLED, SQ_RISE, SQ_SIGNALHI, SQ_FALL, SQ_SIGNALLO, SQ_TIMEOUT, SQ_DISABLED };

struct LaneState {                 // carried across calls, one per lane
    float dc_re, dc_im;            // dc blocker v1
    float g, y2p;                  // agc gain, filtered output energy
    int mode; unsigned timer;      // squelch FSM
    float fm_re, fm_im;            // freqdem r_prime
};
struct SegState { float g, y2p; int mode; unsigned timer; float fm_re, fm_im; };

struct FsmState;
struct BackendParams {
    const float2 *in; long long in_lane_stride;
    void *out; long long out_lane_stride;      // float (demod != 0) or float2 elements
    int n, nlanes;
    int L, W, G, nseg, ngrp;
    int has_dc, has_agc, demod;                // demod: 0 none (cf32 out), 1 fm (float out)
    float dc_a1;                               // a[1] = -1 + alpha_dc
    float alpha; float one_minus_alpha_f; float neg_half_alpha;
    float g_thr;                               // rssi > threshold  <=>  g < g_thr  (bisected on the host)
    unsigned timeout; float fm_ref;
    int squelch_enabled;
    int exact_math;                            // 1: library expf/logf/atan2f in the per-sample loop (slower)
    int gate;                                  // 1: zero the output unless squelch status == SIGNALHI (Liquid.chs:700-704)
    LaneState *lane;
    SegState *seg_start, *seg_end;             // [nlanes][nseg] gain-loop state at segment boundaries
    const double2 *dcVloc, *dcCarry; const double *dcPowA; int nblk;   // dc state at group boundaries (see dc blocker)
    int nwords, FW;                            // 32-sample words per lane; FSM replay length in segments
    unsigned *exbits, *gatebits;               // [nlanes][nwords] threshold-exceeded / gate-open bit per sample
    unsigned *sgnr, *sgni;                     // [nlanes][nwords] sign bits of the ungated agc output (discriminator
                                               // values next to a closed gate are signed-zero artefacts: +-pi or 0)
    unsigned *prev_sign;                       // [nlanes] sign bits (re | im << 1) of the sample before this chunk
    FsmState *fsm_start, *fsm_end;      // [nlanes][nseg]
    unsigned *prev_gate;                       // [nlanes] gate of the sample before this chunk
    unsigned *first_bad;                       // [nlanes][2] first segment whose start state does not continue its
                                               // predecessor (gain loop, squelch FSM); 0xffffffff = none
    unsigned *bad_list; unsigned *bad_count; unsigned bad_cap;   // segments to refine after the first verification
    unsigned long long *fixups;                // [3] segments re-run in order: gain loop, squelch FSM; refined in parallel
};

// ------------------------------------------------------------------------------------------ dc blocker
// v[n] = x[n] + c v[n-1] (c = 1 - alpha), y[n] = v[n] - v[n-1].  The state at every G-sample group boundary is
//   V(j) = Vloc[j-1] + carry[b] * A^k ,   A = c^G, b = (j-1) / kDcGB, k = (j-1) % kDcGB + 1
// Vloc = in-block inclusive scan of the groups' zero-state responses (k_dc_local, one group per thread, kDcGB
// groups per CTA), carry[b] = state at the start of block b (k_dc_carry, sequential over the few blocks), all fp64.
constexpr int kDcGB = 256;

struct DcParams {
    const float2 *in; long long in_lane_stride;
    float2 *out; long long out_lane_stride;
    int n, nlanes, G, ngrp, nblk;
    double c;                                  // 1 - alpha  (= -a1)
    float a1;
    double2 *Vloc;                             // [nlanes][ngrp]  block-local state at the END of group j
    double2 *carry;                            // [nlanes][nblk]  state at the start of block b
    const double *powA;                        // [kDcGB + 1]     A^k
    LaneState *lane;
};

__device__ __forceinline__ double2 dc_state_at(const double2 *Vloc, const double2 *carry, const double *powA,
                                               int ngrp, int nblk, int lane, int j)
{
    // filter state before group j (= after j groups)
    const double2 *cl = carry + (long long)lane * nblk;
    if (j == 0) return cl[0];
    const int b = (j - 1) / kDcGB, k = (j - 1) - b * kDcGB + 1;
    const double2 v = Vloc[(long long)lane * ngrp + (j - 1)], cb = cl[b];
    const double a = powA[k];
    return make_double2(v.x + cb.x * a, v.y + cb.y * a);
}

__global__ void __launch_bounds__(kDcGB) k_dc_local(const DcParams p)
{
    __shared__ double sr[kDcGB], si[kDcGB];
    const int lane = blockIdx.y, t = threadIdx.x, j = blockIdx.x * kDcGB + t;
    const float2 *x = p.in + (long long)lane * p.in_lane_stride;
    double ar = 0.0, ai = 0.0;
    if (j < p.ngrp) {
        const int i0 = j * p.G, i1 = min(i0 + p.G, p.n);
        for (int i = i0; i < i1; i++) {
            const float2 v = x[i];
            ar = ar * p.c + (double)v.x;
            ai = ai * p.c + (double)v.y;
        }
        // a short last group is completed with zero input so that every group advances the state by c^G
        for (int i = i1; i < i0 + p.G; i++) { ar *= p.c; ai *= p.c; }
    }
    sr[t] = ar; si[t] = ai;
    __syncthreads();
    double f = p.powA[1];
    for (int d = 1; d < kDcGB; d <<= 1) {
        double vr = 0.0, vi = 0.0;
        if (t >= d) { vr = sr[t - d] * f; vi = si[t - d] * f; }
        __syncthreads();
        if (t >= d) { sr[t] += vr; si[t] += vi; }
        __syncthreads();
        f *= f;
    }
    if (j < p.ngrp) p.Vloc[(long long)lane * p.ngrp + j] = make_double2(sr[t], si[t]);
}

// one CTA per lane: carries of the (few) blocks, then the state after the last sample goes into the lane state
__global__ void __launch_bounds__(256) k_dc_carry(const DcParams p)
{
    const int lane = blockIdx.x, t = threadIdx.x;
    __shared__ double ar[1024], ai[1024];
    double2 *cl = p.carry + (long long)lane * p.nblk;
    // block aggregates (state contribution of a full block), fetched in parallel
    for (int b = t; b < p.nblk && b < 1024; b += blockDim.x) {
        const int jl = (b + 1) * kDcGB - 1;
        double2 v = make_double2(0.0, 0.0);
        if (jl < p.ngrp) v = p.Vloc[(long long)lane * p.ngrp + jl];
        ar[b] = v.x; ai[b] = v.y;
    }
    __syncthreads();
    if (t == 0) {
        double cr = (double)p.lane[lane].dc_re, ci = (double)p.lane[lane].dc_im;
        const double AB = p.powA[kDcGB];
        for (int b = 0; b < p.nblk; b++) {
            cl[b] = make_double2(cr, ci);
            double vr, vi;
            if (b < 1024) { vr = ar[b]; vi = ai[b]; }
            else {
                const int jl = (b + 1) * kDcGB - 1;
                double2 v = make_double2(0.0, 0.0);
                if (jl < p.ngrp) v = p.Vloc[(long long)lane * p.ngrp + jl];
                vr = v.x; vi = v.y;
            }
            cr = cr * AB + vr; ci = ci * AB + vi;
        }
        // state after n samples: restart from the last full-group boundary
        const int jf = p.n / p.G;
        const double2 v = dc_state_at(p.Vloc, p.carry, p.powA, p.ngrp, p.nblk, lane, jf);
        float v1r = (float)v.x, v1i = (float)v.y;
        const float2 *x = p.in + (long long)lane * p.in_lane_stride;
        for (int i = jf * p.G; i < p.n; i++) {
            const float2 s = x[i];
            v1r = __fsub_rn(s.x, __fmul_rn(p.a1, v1r));
            v1i = __fsub_rn(s.y, __fmul_rn(p.a1, v1i));
        }
        p.lane[lane].dc_re = v1r; p.lane[lane].dc_im = v1i;
    }
}

// stand-alone dc blocker output (iirfilt_crcf_execute_block): one thread per group restarts the float32
// recurrence from the exact boundary state.  May run in place (k_dc_carry has already read what it needs).
__global__ void k_dc_apply(const DcParams p)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.ngrp) return;
    int lane = (int)(t / p.ngrp), j = (int)(t - (long long)lane * p.ngrp);
    const float2 *x = p.in + (long long)lane * p.in_lane_stride;
    float2 *y = p.out + (long long)lane * p.out_lane_stride;
    const double2 v = dc_state_at(p.Vloc, p.carry, p.powA, p.ngrp, p.nblk, lane, j);
    float v1r = (float)v.x, v1i = (float)v.y;
    int i0 = j * p.G, i1 = min(i0 + p.G, p.n);
    for (int i = i0; i < i1; i++) {
        float2 s = x[i];
        float v0r = __fsub_rn(s.x, __fmul_rn(p.a1, v1r));
        float v0i = __fsub_rn(s.y, __fmul_rn(p.a1, v1i));
        y[i] = cf(__fsub_rn(v0r, v1r), __fsub_rn(v0i, v1i));
        v1r = v0r; v1i = v0i;
    }
}

// ------------------------------------------------------------------------------------------ agc + fm
// The squelch FSM does not feed back into the gain loop, so the work is split:
//   pass G (k_backend_spec / k_backend_fixup): the gain trajectory.  Each L-sample segment is run by its own thread
//       after a W-sample warm-up (the loop is contractive: perturbations decay like (1-alpha)^(k/2)); it writes the
//       UNGATED outputs (agc samples, or discriminator values of ungated neighbours) and one "rssi > threshold" bit
//       per sample.  Segment start states are verified against the predecessor's end state (1e-6 relative) and the
//       rare misses (e.g. a gain frozen by digital silence) are re-run in stream order.
//   pass F (k_backend_fsm / k_backend_fsm_fix): the squelch state machine, EXACT, on the bit stream.  Each segment
//       re-derives its entry state by replaying the bits of the preceding FW segments (any entry state is forgotten
//       after timeout+4 samples except for measure-zero coincidences), emits one gate bit per sample, and entry states
//       are again verified against the predecessor's exit state and repaired in order where they differ.
//   pass A (k_backend_gate): zero the outputs where the gate is closed: y[i] if !gate[i]; the discriminator output
//       m[i] = arg(conj(r[i-1]) r[i]) if !(gate[i-1] && gate[i])  (arg(0) = 0, as in liquid).
struct AgcRun { float g, y2p; float fr, fi; };

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

// arg(x + jy) with a degree-8 minimax polynomial for atan on [0, 1] (max error 1.1e-7 rad, float32-limited; the
// library atan2f is ~3x the instructions).  Exact zeros keep the library's signed-zero semantics.
__device__ __forceinline__ float be_atan2(float y, float x)
{
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    if (mx == 0.f || !(mx < 1e37f)) return atan2f(y, x);
    const float a = __fdividef(mn, mx);
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
    if (x < 0.f) r = 3.14159274f - r;
    return copysignf(r, y);
}

// compile-time configuration of the per-sample loop (runtime flags inside it would fence the instruction scheduler)
enum { BE_DC = 1, BE_AGC = 2, BE_FM = 4, BE_EXACT = 8, BE_NCFG = 16 };

// one sample of the gain loop: ungated agc output (yr, yi), threshold bit, ungated discriminator value
template <int CFG>
__device__ __forceinline__ bool be_step(const BackendParams &p, AgcRun &s, float xr, float xi, float &yr, float &yi,
                                        float &m)
{
    bool ex = true;
    if (CFG & BE_AGC) {
        // AGC(_execute), liquid agc.c
        yr = __fmul_rn(xr, s.g); yi = __fmul_rn(xi, s.g);
        float y2 = __fadd_rn(__fmul_rn(yr, yr), __fmul_rn(yi, yi));
        // liquid evaluates (1.0 - alpha) * y2' + alpha * y2 in double and rounds to float; the float32 FMA below differs
        // from it by at most one ulp of y2' (6e-8 relative), i.e. 3e-9 per step in the gain: far below the 1e-6 at
        // which two float32 runs of this loop settle anyway
        s.y2p = fmaf(p.one_minus_alpha_f, s.y2p, __fmul_rn(p.alpha, y2));
        // g *= y2'^(-alpha/2) unless y2' <= 1e-6.  Default: SFU exp2/log2 (each step is good to ~3e-7 relative and
        // the loop is contractive, so the gain stays within ~1e-6 of the libm evaluation); BE_EXACT: expf/logf.
        const float f = (CFG & BE_EXACT) ? expf(p.neg_half_alpha * logf(s.y2p)) : __expf(p.neg_half_alpha * __logf(s.y2p));
        s.g *= (s.y2p > 1e-6f) ? f : 1.0f;
        s.g = fminf(s.g, 1e6f);
        ex = s.g < p.g_thr;                       // rssi = -20 log10(g) > threshold
    } else { yr = xr; yi = xi; }
    if (CFG & BE_FM) {
        // freqdem_demodulate: arg(conj(r') r) / (2 pi kf)
        float re = __fadd_rn(__fmul_rn(s.fr, yr), __fmul_rn(s.fi, yi));
        float im = __fsub_rn(__fmul_rn(s.fr, yi), __fmul_rn(s.fi, yr));
        m = ((CFG & BE_EXACT) ? atan2f(im, re) : be_atan2(im, re)) * p.fm_ref;
        s.fr = yr; s.fi = yi;
    }
    return ex;
}

__device__ __forceinline__ bool be_close(float a, float b, float atol = 0.f)
{
    return fabsf(a - b) <= 1e-5f * fmaxf(fabsf(a), fabsf(b)) + atol;   // fp32 rounding keeps two runs ~1e-6 apart
}
__device__ __forceinline__ bool be_match(const SegState &a, const SegState &b, int has_agc, int demod)
{
    // the discriminator history is the previous ungated sample x * g: it continues whenever the gain does (and is
    // exactly the previous input sample when there is no AGC)
    (void)demod;
    return !has_agc || (be_close(a.g, b.g) && be_close(a.y2p, b.y2p));
}

// run samples [i0, i1) of one lane from state s / dc state (v1r, v1i).  EMIT: write outputs, threshold bits and
// sign bits (the caller owns whole 32-sample words: i0 % 32 == 0 and i1 is a multiple of 32 or the chunk end).
// Samples are fetched eight at a time, one block ahead of the recurrence, so that the sequential chain never waits
// for memory.
template <bool EMIT, int CFG>
__device__ __forceinline__ void be_run(const BackendParams &p, int lane, AgcRun &s, float &v1r, float &v1i, int i0,
                                       int i1)
{
    const float2 *__restrict__ x = p.in + (long long)lane * p.in_lane_stride;
    float *__restrict__ of = (float *)p.out + (long long)lane * p.out_lane_stride;
    float2 *__restrict__ oc = (float2 *)p.out + (long long)lane * p.out_lane_stride;
    unsigned *bits = p.exbits + (long long)lane * p.nwords;
    unsigned *sr = p.sgnr + (long long)lane * p.nwords, *si = p.sgni + (long long)lane * p.nwords;
    unsigned word = 0, wr = 0, wi = 0;
    constexpr int B = 8;
    float2 cur[B], nxt[B];
    // 16-byte loads (two samples each) when the lane's samples are 16-byte aligned: every lane reads its own cache
    // line, so the number of load instructions is what the LSU pays for
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((i0 & 1) == 0);
    auto fetch = [&](float2 (&dst)[B], int i) {
        if (vec && i + B <= i1) {
            const float4 *x4 = reinterpret_cast<const float4 *>(x + i);
#pragma unroll
            for (int k = 0; k < B / 2; k++) { const float4 q = __ldg(x4 + k); dst[2 * k] = cf(q.x, q.y); dst[2 * k + 1] = cf(q.z, q.w); }
        } else {
#pragma unroll
            for (int k = 0; k < B; k++) dst[k] = (i + k < i1) ? __ldg(x + i + k) : cf(0.f, 0.f);
        }
    };
    const bool vec_out = EMIT && (CFG & BE_FM) && ((reinterpret_cast<uintptr_t>(of) & 15) == 0) && ((i0 & 3) == 0);
    float mbuf[B];
    if (i0 < i1) fetch(nxt, i0);
    for (int i = i0; i < i1; i += B) {
#pragma unroll
        for (int k = 0; k < B; k++) cur[k] = nxt[k];
        if (i + B < i1) fetch(nxt, i + B);
#pragma unroll
        for (int k = 0; k < B; k++) {
            if (i + k >= i1) break;
            float xr = cur[k].x, xi = cur[k].y;
            if (CFG & BE_DC) {
                float v0r = __fsub_rn(xr, __fmul_rn(p.dc_a1, v1r));
                float v0i = __fsub_rn(xi, __fmul_rn(p.dc_a1, v1i));
                xr = __fsub_rn(v0r, v1r); xi = __fsub_rn(v0i, v1i);
                v1r = v0r; v1i = v0i;
            }
            float yr, yi, m = 0.f;
            const bool ex = be_step<CFG>(p, s, xr, xi, yr, yi, m);
            if (EMIT) {
                const int ii = i + k;
                if (CFG & BE_FM) { if (vec_out) mbuf[k] = m; else of[ii] = m; } else oc[ii] = cf(yr, yi);
                word |= (ex ? 1u : 0u) << (ii & 31);
                wr |= ((unsigned)__float_as_int(yr) >> 31) << (ii & 31);
                wi |= ((unsigned)__float_as_int(yi) >> 31) << (ii & 31);
                if ((ii & 31) == 31 || ii == i1 - 1) {
                    bits[ii >> 5] = word; word = 0;
                    if (CFG & BE_FM) { sr[ii >> 5] = wr; si[ii >> 5] = wi; }
                    wr = 0; wi = 0;
                }
            }
        }
        if (EMIT && vec_out) {
            if (i + B <= i1) {
                float4 *o4 = reinterpret_cast<float4 *>(of + i);
                o4[0] = make_float4(mbuf[0], mbuf[1], mbuf[2], mbuf[3]);
                o4[1] = make_float4(mbuf[4], mbuf[5], mbuf[6], mbuf[7]);
            } else {
#pragma unroll
                for (int k = 0; k < B; k++) if (i + k < i1) of[i + k] = mbuf[k];
            }
        }
    }
}

__device__ __forceinline__ void be_dc_state(const BackendParams &p, int lane, int i, float &v1r, float &v1i)
{
    // i is a multiple of G
    if (p.has_dc) {
        const double2 v = dc_state_at(p.dcVloc, p.dcCarry, p.dcPowA, p.ngrp, p.nblk, lane, i / p.G);
        v1r = (float)v.x; v1i = (float)v.y;
    } else { v1r = 0.f; v1i = 0.f; }
}

__device__ __forceinline__ SegState be_pack(const AgcRun &s)
{
    SegState st; st.g = s.g; st.y2p = s.y2p; st.mode = 0; st.timer = 0; st.fm_re = s.fr; st.fm_im = s.fi;
    return st;
}

template <int CFG>
__global__ void __launch_bounds__(128) k_backend_spec(const BackendParams p)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.nseg) return;
    const int lane = (int)(t / p.nseg), seg = (int)(t - (long long)lane * p.nseg);
    const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
    if (t == 0) *p.bad_count = 0;
    AgcRun s; float v1r, v1i;
    int w0 = b0 - p.W;
    const LaneState ls = p.lane[lane];
    if (w0 <= 0) {
        // the warm-up reaches the chunk start: run from the true carried state (exact)
        w0 = 0;
        s.g = ls.g; s.y2p = ls.y2p; s.fr = ls.fm_re; s.fi = ls.fm_im;
    } else {
        // equilibrium guess: unit output energy for the first samples of the warm-up window
        const float2 *x = p.in + (long long)lane * p.in_lane_stride;
        float e = 0.f;
        for (int i = 0; i < 16; i++) { float2 v = x[w0 + i]; e += v.x * v.x + v.y * v.y; }
        e *= (1.0f / 16.0f);
        s.g = (e > 1e-30f) ? rsqrtf(e) : 1e6f;
        if (s.g > 1e6f) s.g = 1e6f;
        s.y2p = 1.0f; s.fr = 0.f; s.fi = 0.f;
    }
    be_dc_state(p, lane, w0, v1r, v1i);
    be_run<false, CFG>(p, lane, s, v1r, v1i, w0, b0);  // warm-up, nothing emitted
    p.seg_start[t] = be_pack(s);
    be_run<true, CFG>(p, lane, s, v1r, v1i, b0, b1);
    p.seg_end[t] = be_pack(s);
}

// grid-wide verification of the gain speculation.  pass 0: collect the segments whose start state does not continue
// their predecessor's end state; pass 1 (after k_backend_refine): first such segment per lane, for the in-order repair.
__global__ void k_backend_verify(const BackendParams p, int pass)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.nseg) return;
    const int lane = (int)(t / p.nseg), seg = (int)(t - (long long)lane * p.nseg);
    if (seg == 0 || be_match(p.seg_start[t], p.seg_end[t - 1], p.has_agc, p.demod)) return;
    if (pass == 0) {
        const unsigned idx = atomicAdd(p.bad_count, 1u);
        if (idx < p.bad_cap) p.bad_list[idx] = (unsigned)t;
    } else {
        atomicMin(&p.first_bad[2 * lane], (unsigned)seg);
    }
}

// second chance, in parallel: a start state that is slightly off (the warm-up met an un-damped stretch of the
// loop) is replaced by the predecessor's END state, which is accurate because the predecessor's own L samples
// damped its error; the segment is re-run from there.  (A predecessor that is being refined at the same time may
// be read before or after its update: both values are valid to well below the tolerance.)
template <int CFG>
__global__ void __launch_bounds__(128) k_backend_refine(const BackendParams p)
{
    const unsigned count = min(*p.bad_count, p.bad_cap);
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < count; idx += gridDim.x * blockDim.x) {
        const long long t = p.bad_list[idx];
        const int lane = (int)(t / p.nseg), seg = (int)(t - (long long)lane * p.nseg);
        const SegState pe = p.seg_end[t - 1];
        AgcRun s; s.g = pe.g; s.y2p = pe.y2p; s.fr = pe.fm_re; s.fi = pe.fm_im;
        float v1r, v1i;
        const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
        be_dc_state(p, lane, b0, v1r, v1i);
        p.seg_start[t] = pe;
        be_run<true, CFG>(p, lane, s, v1r, v1i, b0, b1);
        p.seg_end[t] = be_pack(s);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.fixups + 2, (unsigned long long)count);
}

// the list is consumed: empty it for the next verification round
__global__ void k_backend_list_reset(const BackendParams p) { *p.bad_count = 0; }

// one CTA per lane: re-run the misses in stream order (rare), store the lane's gain state
template <int CFG>
__global__ void k_backend_fixup(const BackendParams p)
{
    const int lane = blockIdx.x;
    __shared__ unsigned s_next;
    __shared__ int s_cur;
    SegState *E = p.seg_end + (long long)lane * p.nseg;
    const SegState *S0 = p.seg_start + (long long)lane * p.nseg;
    bool first = true;
    if (threadIdx.x == 0) s_cur = 1;
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) s_next = first ? p.first_bad[2 * lane] : 0xffffffffu;
        __syncthreads();
        if (!first) {
            // after a repair: parallel search for the next segment >= s_cur that does not continue its predecessor
            const int cur = s_cur;
            for (int j = cur + threadIdx.x; j < p.nseg; j += blockDim.x)
                if (!be_match(S0[j], E[j - 1], p.has_agc, p.demod)) { atomicMin(&s_next, (unsigned)j); break; }
            __syncthreads();
        }
        first = false;
        const unsigned nxt = s_next;
        if (nxt == 0xffffffffu) break;
        if (threadIdx.x == 0) {
            int seg = (int)nxt;
            unsigned long long redone = 0;
            while (seg < p.nseg) {
                const SegState pe = E[seg - 1];
                if (be_match(S0[seg], pe, p.has_agc, p.demod)) break;
                AgcRun s; s.g = pe.g; s.y2p = pe.y2p; s.fr = pe.fm_re; s.fi = pe.fm_im;
                float v1r, v1i;
                const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
                be_dc_state(p, lane, b0, v1r, v1i);
                be_run<true, CFG>(p, lane, s, v1r, v1i, b0, b1);
                E[seg] = be_pack(s);
                redone++;
                seg++;      // the successor is re-checked against the new end state on the next iteration
            }
            s_cur = seg;
            atomicAdd(p.fixups, redone);
        }
        __syncthreads();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const SegState e = E[p.nseg - 1];
        LaneState ls = p.lane[lane];
        p.prev_sign[lane] = ((unsigned)__float_as_int(ls.fm_re) >> 31) | (((unsigned)__float_as_int(ls.fm_im) >> 31) << 1);
        ls.g = e.g; ls.y2p = e.y2p; ls.fm_re = e.fm_re; ls.fm_im = e.fm_im;
        p.lane[lane] = ls;
        p.first_bad[2 * lane] = 0xffffffffu;
    }
}

// ---- squelch FSM on the threshold bits
struct FsmState { int mode; unsigned timer; };
__device__ __forceinline__ bool fsm_same(const FsmState &a, const FsmState &b)
{
    return a.mode == b.mode && (a.mode != SQ_SIGNALLO || a.timer == b.timer);
}

// run the FSM over bits [i0, i1) of a lane; when gate != nullptr write one gate bit (mode == SIGNALHI) per sample
__device__ __forceinline__ void fsm_run(const BackendParams &p, const unsigned *bits, unsigned *gate, FsmState &s,
                                        int i0, int i1)
{
    int i = i0;
    while (i < i1) {
        const int w = i >> 5, lo = i & 31, cnt = min(32 - lo, i1 - i);
        const unsigned full = (cnt == 32) ? 0xffffffffu : ((1u << cnt) - 1u);
        const unsigned word = (bits[w] >> lo) & full;
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

__global__ void __launch_bounds__(128) k_backend_fsm(const BackendParams p)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.nseg) return;
    const int lane = (int)(t / p.nseg), seg = (int)(t - (long long)lane * p.nseg);
    const unsigned *bits = p.exbits + (long long)lane * p.nwords;
    unsigned *gate = p.gatebits + (long long)lane * p.nwords;
    const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
    FsmState s;
    int w0 = b0 - p.FW * p.L;
    if (w0 <= 0) { w0 = 0; s.mode = p.lane[lane].mode; s.timer = p.lane[lane].timer; }   // exact
    else { s.mode = SQ_ENABLED; s.timer = 0; }
    fsm_run(p, bits, nullptr, s, w0, b0);
    p.fsm_start[t] = s;
    fsm_run(p, bits, gate, s, b0, b1);
    p.fsm_end[t] = s;
}

__global__ void k_backend_fsm_verify(const BackendParams p)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.nseg) return;
    const int lane = (int)(t / p.nseg), seg = (int)(t - (long long)lane * p.nseg);
    if (seg > 0 && !fsm_same(p.fsm_start[t], p.fsm_end[t - 1])) atomicMin(&p.first_bad[2 * lane + 1], (unsigned)seg);
}

__global__ void k_backend_fsm_fix(const BackendParams p)
{
    const int lane = blockIdx.x;
    __shared__ unsigned s_next;
    __shared__ int s_cur;
    FsmState *E = p.fsm_end + (long long)lane * p.nseg;
    const FsmState *S0 = p.fsm_start + (long long)lane * p.nseg;
    const unsigned *bits = p.exbits + (long long)lane * p.nwords;
    unsigned *gate = p.gatebits + (long long)lane * p.nwords;
    bool first = true;
    if (threadIdx.x == 0) s_cur = 1;
    __syncthreads();
    while (true) {
        if (threadIdx.x == 0) s_next = first ? p.first_bad[2 * lane + 1] : 0xffffffffu;
        __syncthreads();
        if (!first) {
            const int cur = s_cur;
            for (int j = cur + threadIdx.x; j < p.nseg; j += blockDim.x)
                if (!fsm_same(S0[j], E[j - 1])) { atomicMin(&s_next, (unsigned)j); break; }
            __syncthreads();
        }
        first = false;
        const unsigned nxt = s_next;
        if (nxt == 0xffffffffu) break;
        if (threadIdx.x == 0) {
            int seg = (int)nxt;
            unsigned long long redone = 0;
            while (seg < p.nseg) {
                FsmState s = E[seg - 1];
                if (fsm_same(S0[seg], s)) break;
                const int b0 = seg * p.L, b1 = min(b0 + p.L, p.n);
                fsm_run(p, bits, gate, s, b0, b1);
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
    if (threadIdx.x == 0) {
        // gate of the last sample of the previous chunk, needed by the discriminator's first sample
        p.prev_gate[lane] = (p.lane[lane].mode == SQ_SIGNALHI) ? 1u : 0u;
        const FsmState e = E[p.nseg - 1];
        p.lane[lane].mode = e.mode; p.lane[lane].timer = e.timer;
        p.first_bad[2 * lane + 1] = 0xffffffffu;
    }
}

// apply the gate.  One thread per 32 samples.  cf32 output: closed samples become 0+0j (Liquid.chs:704).
// Discriminator output: m[i] = arg(conj(r'[i-1]) r'[i]) with r' the GATED samples; when either neighbour is closed the
// product is a signed zero whose argument is 0 or +-pi exactly as in the sequential code, so those values are
// recomputed here from the sign bits of the ungated samples.
__global__ void k_backend_gate(const BackendParams p)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.nwords) return;
    const int lane = (int)(t / p.nwords), w = (int)(t - (long long)lane * p.nwords);
    const unsigned *gate = p.gatebits + (long long)lane * p.nwords;
    const int i0 = w * 32, cnt = min(32, p.n - i0);
    const unsigned full = (cnt == 32) ? 0xffffffffu : ((1u << cnt) - 1u);
    const unsigned g = gate[w] & full;
    if (p.demod != 1) {
        if (g == full) return;
        float2 *oc = (float2 *)p.out + (long long)lane * p.out_lane_stride + i0;
        for (int b = 0; b < cnt; b++)
            if (!((g >> b) & 1u)) oc[b] = cf(0.f, 0.f);
        return;
    }
    const unsigned gprev = (g << 1) | (w ? (gate[w - 1] >> 31) : p.prev_gate[lane]);
    if ((g & gprev & full) == full) return;              // every sample and its predecessor are open
    const unsigned *sr = p.sgnr + (long long)lane * p.nwords, *si = p.sgni + (long long)lane * p.nwords;
    const unsigned r = sr[w], im = si[w];
    const unsigned rprev = (r << 1) | (w ? (sr[w - 1] >> 31) : (p.prev_sign[lane] & 1u));
    const unsigned iprev = (im << 1) | (w ? (si[w - 1] >> 31) : ((p.prev_sign[lane] >> 1) & 1u));
    float *of = (float *)p.out + (long long)lane * p.out_lane_stride + i0;
    for (int b = 0; b < cnt; b++) {
        const bool open = (g >> b) & 1u, popen = (gprev >> b) & 1u;
        if (open && popen) continue;
        // unit-magnitude stand-ins carry the signs; a closed sample is +0+0j
        const float yr = open ? (((r >> b) & 1u) ? -1.f : 1.f) : 0.f;
        const float yi = open ? (((im >> b) & 1u) ? -1.f : 1.f) : 0.f;
        const float fr = popen ? (((rprev >> b) & 1u) ? -1.f : 1.f) : 0.f;
        const float fi = popen ? (((iprev >> b) & 1u) ? -1.f : 1.f) : 0.f;
        const float re = __fadd_rn(__fmul_rn(fr, yr), __fmul_rn(fi, yi));
        const float ii = __fsub_rn(__fmul_rn(fr, yi), __fmul_rn(fi, yr));
        of[b] = atan2f(ii, re) * p.fm_ref;
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
        for (int c = 1; c < nlanes; c++) acc = __fadd_rn(acc, in[(long long)c * lane_stride + i]);
        out[i] = acc;
    }
}

// ------------------------------------------------------------------------------------------ launch sequences
// `launch(kernel, grid, block, smem_bytes, args...)` is supplied by the caller (CUDA stream launcher in
// csdr_b200.cu, thread emulator in the CPU-only tests) so that both run the same sequence with the same grids.
template <class Launch>
inline void be_launch_dc(Launch &launch, const DcParams &d, bool apply)
{
    launch(k_dc_local, dim3(d.nblk, d.nlanes), dim3(kDcGB), 0, d);
    launch(k_dc_carry, dim3(d.nlanes), dim3(256), 0, d);
    if (apply) {
        const long long items = (long long)d.nlanes * d.ngrp;
        launch(k_dc_apply, dim3((unsigned)((items + 127) / 128)), dim3(128), 0, d);
    }
}

template <int CFG, class Launch>
inline void be_launch_gain(Launch &launch, const BackendParams &b, unsigned gb)
{
    launch(k_backend_spec<CFG>, dim3(gb), dim3(128), 0, b);
    if (CFG & BE_AGC) {
        // two rounds of verify + parallel refine (a run of consecutive misses needs one round per level of
        // inaccuracy handed down the run; an empty list costs a few microseconds), then the in-order repair
        launch(k_backend_verify, dim3(gb), dim3(128), 0, b, 0);
        launch.debug_after_verify(b);
        launch(k_backend_refine<CFG>, dim3(64), dim3(128), 0, b);
        launch(k_backend_list_reset, dim3(1), dim3(1), 0, b);
        launch(k_backend_verify, dim3(gb), dim3(128), 0, b, 0);
        launch(k_backend_refine<CFG>, dim3(64), dim3(128), 0, b);
        launch(k_backend_verify, dim3(gb), dim3(128), 0, b, 1);
    }
    launch(k_backend_fixup<CFG>, dim3(b.nlanes), dim3(128), 0, b);
}
template <int N, class Launch>
inline void be_dispatch_gain(int cfg, Launch &launch, const BackendParams &b, unsigned gb)
{
    if constexpr (N >= 0) {
        if (cfg == N) be_launch_gain<N>(launch, b, gb);
        else be_dispatch_gain<N - 1>(cfg, launch, b, gb);
    }
}

template <class Launch>
inline void be_launch(Launch &launch, const BackendParams &b)
{
    const long long segs = (long long)b.nlanes * b.nseg;
    const unsigned gb = (unsigned)((segs + 127) / 128);
    const int cfg = (b.has_dc ? BE_DC : 0) | (b.has_agc ? BE_AGC : 0) | (b.demod == 1 ? BE_FM : 0) | (b.exact_math ? BE_EXACT : 0);
    be_dispatch_gain<BE_NCFG - 1>(cfg, launch, b, gb);
    if (b.has_agc) {
        launch(k_backend_fsm, dim3(gb), dim3(128), 0, b);
        launch(k_backend_fsm_verify, dim3(gb), dim3(128), 0, b);
        launch(k_backend_fsm_fix, dim3(b.nlanes), dim3(128), 0, b);
        if (b.gate) {
            const long long words = (long long)b.nlanes * b.nwords;
            launch(k_backend_gate, dim3((unsigned)((words + 127) / 128)), dim3(128), 0, b);
        }
    }
}

}  // namespace csdr
