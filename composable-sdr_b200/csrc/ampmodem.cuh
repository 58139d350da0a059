// ampmodem.cuh -- ampmodem_create(0.8, LIQUID_AMPMODEM_DSB, 0) demodulator (Liquid.chs:452-469, liquid
// src/modem/src/ampmodem.c, 2019 redesign).  Two candidate DSB/carrier-present demodulators exist in liquid's
// history (confidence LOW); both are implemented, CSDR_OPT_AMPMODEM_PLL selects:
//   peak detector : |x| -> 51-tap real dc-blocking FIR -> / mod_index              (fully time-parallel)
//   carrier PLL   : 51-tap low-pass isolates the carrier, an NCO/PLL tracks it, the 25-sample-delayed signal is
//                   mixed down, Re()/mod_index, then the same dc-blocking FIR      (PLL loop: one thread per lane)
#pragma once
#include "platform.cuh"
#ifndef CSDR_EMU
#include "design.hpp"
#include <vector>
#endif

namespace csdr {

constexpr int kAmM = 25;                 // filter semi-length (ampmodem.c: q->m = 25)
constexpr int kAmTaps = 2 * kAmM + 1;    // 51
constexpr int kAmHist = kAmTaps - 1;     // 50 samples of history per lane

struct AmParams {
    int nlanes, n;
    const float2 *x; long long x_stride;          // input chunk
    float2 *xh; long long xh_stride;              // [hist 50 | n] copy of the input
    float2 *x0;                                   // [lanes][n] low-passed (PLL)
    float *mh; long long mh_stride;               // [hist 50 | n] pre-dc-block real signal
    float *y; long long y_stride;                 // output
    const float *h_lp, *h_dc;                     // 51 taps each; h[k] multiplies sample (i - k)
    const float *sintab;                          // 1024-entry NCO table (pll)
    unsigned *pll;                                // [lanes][2] theta, d_theta
    float inv_mod, pll_alpha, pll_beta, out_scale;
    int use_pll;
    // time-parallel carrier loop (k_am_pll_spec / k_am_pll_fix): segments of pll_L samples after a pll_W-sample pull-in
    int pll_L, pll_W, pll_nseg;
    unsigned *seg_start, *seg_end;                // [lanes][nseg][2] loop state at the segment boundaries
    unsigned tol_theta, tol_dtheta;               // two runs of the loop count as the same within these (phase words)
    unsigned long long *pll_redone;               // segments re-run in stream order (diagnostic)
};

// xh[50+i] = x[i]; peak variant also writes mh[50+i] = |x[i]|
__global__ void k_am_stage_in(const AmParams p)
{
    const int lane = blockIdx.y;
    const float2 *x = p.x + (long long)lane * p.x_stride;
    float2 *xh = p.xh + (long long)lane * p.xh_stride + kAmHist;
    float *mh = p.mh + (long long)lane * p.mh_stride + kAmHist;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        float2 v = x[i];
        if (p.use_pll) xh[i] = v;
        else mh[i] = hypotf(v.x, v.y);
    }
}

// The two 51-tap FIRs: kAmR consecutive outputs per thread, the taps in registers, every staged sample read once per
// thread and used by up to kAmR accumulators.  Output r takes its taps oldest sample first (liquid's dotprod order):
// sample j of the thread's window is tap r + 50 - j of output r.  Tile element e sits at e + (e >> 2): consecutive
// threads (4 elements apart) then start 5 elements apart, which is conflict-free for 4- and 8-byte accesses.
constexpr int kAmR = 4, kAmTile = 256 * kAmR;
__device__ __forceinline__ int am_pad(int e) { return e + (e >> 2); }

// x0[i] = sum_k h_lp[k] * xh[50 + i - k]
__global__ void __launch_bounds__(256) k_am_lowpass(const AmParams p)
{
    __shared__ float2 tile[kAmTile + kAmHist + (kAmTile + kAmHist) / 4 + 1];
    const int lane = blockIdx.y;
    const float2 *xh = p.xh + (long long)lane * p.xh_stride;
    float h[kAmTaps];
#pragma unroll
    for (int k = 0; k < kAmTaps; k++) h[k] = p.h_lp[k];
    for (int base = blockIdx.x * kAmTile; base < p.n; base += gridDim.x * kAmTile) {
        __syncthreads();
        for (int j = threadIdx.x; j < kAmTile + kAmHist; j += blockDim.x)
            tile[am_pad(j)] = (base + j < p.n + kAmHist) ? xh[base + j] : cf(0.f, 0.f);
        __syncthreads();
        const int o = threadIdx.x * kAmR;
        if (base + o < p.n) {
            float ar[kAmR], ai[kAmR];
#pragma unroll
            for (int r = 0; r < kAmR; r++) { ar[r] = 0.f; ai[r] = 0.f; }
#pragma unroll
            for (int j = 0; j < kAmHist + kAmR; j++) {
                const float2 v = tile[am_pad(o + j)];
#pragma unroll
                for (int r = 0; r < kAmR; r++) {
                    const int k = r + kAmHist - j;
                    if (k >= 0 && k < kAmTaps) { ar[r] += h[k] * v.x; ai[r] += h[k] * v.y; }
                }
            }
            float2 *dst = p.x0 + (long long)lane * p.n + base + o;
#pragma unroll
            for (int r = 0; r < kAmR; r++) if (base + o + r < p.n) dst[r] = cf(ar[r], ai[r]);
        }
    }
}

__device__ __forceinline__ unsigned am_constrain(float theta)
{
    float pp = (float)((double)theta * 0.159154943091895);
    float frac = pp - truncf(pp);                      // == pp - (float)((long long)pp): the integer part is exact in float
    if (frac < 0.0f) frac = __fadd_rn(frac, 1.0f);     // == (float)((double)frac + 1.0): one rounding either way
    float scaled = frac * 4294967296.0f;
    return scaled >= 4294967296.0f ? 0u : (unsigned)scaled;
}

// carrier PLL, sequential per lane (ampmodem_demod_dsb_pll_carrier).  The loop is a dependent chain
// theta -> table -> mix -> constrain -> theta; the sine table sits in shared memory and the lane's samples are fetched
// eight ahead of the chain so that only the chain's own latency is paid per sample.
__global__ void __launch_bounds__(32) k_am_pll(const AmParams p)
{
    __shared__ float stab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) stab[i] = p.sintab[i];
    __syncthreads();
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= p.nlanes) return;
    const float2 *xh = p.xh + (long long)lane * p.xh_stride + (kAmHist - kAmM);
    const float2 *x0 = p.x0 + (long long)lane * p.n;
    float *mh = p.mh + (long long)lane * p.mh_stride + kAmHist;
    unsigned theta = p.pll[2 * lane], dtheta = p.pll[2 * lane + 1];
    const float inv_mod = p.inv_mod, ka = p.pll_alpha, kb = p.pll_beta;
    constexpr int B = 8;
    float2 na[B], nb[B];
#pragma unroll
    for (int k = 0; k < B; k++) { na[k] = (k < p.n) ? x0[k] : cf(0.f, 0.f); nb[k] = (k < p.n) ? xh[k] : cf(0.f, 0.f); }
    for (int i = 0; i < p.n; i += B) {
        float2 ca[B], cb[B];
#pragma unroll
        for (int k = 0; k < B; k++) { ca[k] = na[k]; cb[k] = nb[k]; }
#pragma unroll
        for (int k = 0; k < B; k++) {
            const int j = i + B + k;
            if (j < p.n) { na[k] = x0[j]; nb[k] = xh[j]; }
        }
#pragma unroll
        for (int k = 0; k < B; k++) {
            if (i + k < p.n) {
                const unsigned idx = ((theta + (1u << 21)) >> 22) & 0x3ffu;
                const float s = stab[idx], c = stab[(idx + 256) & 0x3ffu];
                const float v0i = __fsub_rn(__fmul_rn(ca[k].y, c), __fmul_rn(ca[k].x, s));
                const float v1r = __fadd_rn(__fmul_rn(cb[k].x, c), __fmul_rn(cb[k].y, s));
                dtheta += am_constrain(__fmul_rn(v0i, ka));
                theta += am_constrain(__fmul_rn(v0i, kb));
                theta += dtheta;
                mh[i + k] = v1r * inv_mod;
            }
        }
    }
    p.pll[2 * lane] = theta; p.pll[2 * lane + 1] = dtheta;
}

// ---- the carrier loop over time segments ------------------------------------------------------------------------
// The loop is a second-order PLL (phase gain sqrt(1e-3), frequency gain 1e-3) around a 1024-LEVEL phasor table: a phase
// offset smaller than what flips a table index changes nothing the loop sees, so it is never corrected -- two runs of
// liquid's own loop that differ by 4e-7 rad keep differing by exactly 4e-7 rad (measured on a restatement of the loop), and
// runs that start a few 1e-3 rad apart dither within one table level (6e-3 rad) of each other for ever.  That dead zone
// is why a 1e-7 perturbation of the demodulator's INPUT moves liquid's output by -60 dB (tests/test_gpu_chain.py, config
// 5), and it is the natural meaning of "the same trajectory": behind a 512-sample pull-in from a good guess -- the
// frequency word the lane carried into the chunk, the phase of the low-passed carrier at the window's first sample -- a
// run is within 3e-3 rad rms of the sequential one.  Segments are accepted when their start state continues the
// predecessor's end state within 1.5e-2 rad / 5e-4 rad per sample; the others (the loop is still pulling in: the first
// ~15 k samples of a stream) are re-run in stream order by k_am_pll_fix.  CSDR_OPT_AM_PLL_SEQUENTIAL = 1 runs the one
// sequential loop instead.
__device__ __forceinline__ void am_pll_run(const AmParams &p, const float *stab, const float2 *__restrict__ x0, const float2 *__restrict__ xh,
                                           float *__restrict__ mh, unsigned &theta, unsigned &dtheta, int i0, int i1, bool emit)
{
    const float inv_mod = p.inv_mod, ka = p.pll_alpha, kb = p.pll_beta;
    constexpr int B = 8;
    float2 na[B], nb[B];
#pragma unroll
    for (int k = 0; k < B; k++) { na[k] = (i0 + k < i1) ? x0[i0 + k] : cf(0.f, 0.f); nb[k] = (i0 + k < i1) ? xh[i0 + k] : cf(0.f, 0.f); }
    for (int i = i0; i < i1; i += B) {
        float2 ca[B], cb[B];
#pragma unroll
        for (int k = 0; k < B; k++) { ca[k] = na[k]; cb[k] = nb[k]; }
#pragma unroll
        for (int k = 0; k < B; k++) {
            const int j = i + B + k;
            if (j < i1) { na[k] = x0[j]; nb[k] = xh[j]; }
        }
#pragma unroll
        for (int k = 0; k < B; k++) {
            if (i + k < i1) {
                const unsigned idx = ((theta + (1u << 21)) >> 22) & 0x3ffu;
                const float s = stab[idx], c = stab[(idx + 256) & 0x3ffu];
                const float v0i = __fsub_rn(__fmul_rn(ca[k].y, c), __fmul_rn(ca[k].x, s));
                const float v1r = __fadd_rn(__fmul_rn(cb[k].x, c), __fmul_rn(cb[k].y, s));
                dtheta += am_constrain(__fmul_rn(v0i, ka));
                theta += am_constrain(__fmul_rn(v0i, kb));
                theta += dtheta;
                if (emit) mh[i + k] = v1r * inv_mod;
            }
        }
    }
}
__device__ __forceinline__ bool am_pll_same(const AmParams &p, unsigned th_a, unsigned dth_a, unsigned th_b, unsigned dth_b)
{
    const int dt = (int)(th_a - th_b), dd = (int)(dth_a - dth_b);
    return (unsigned)abs(dt) <= p.tol_theta && (unsigned)abs(dd) <= p.tol_dtheta;
}

__global__ void __launch_bounds__(64) k_am_pll_spec(const AmParams p)
{
    __shared__ float stab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) stab[i] = p.sintab[i];
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)p.nlanes * p.pll_nseg) return;
    const int lane = (int)(t / p.pll_nseg), seg = (int)(t - (long long)lane * p.pll_nseg);
    const float2 *xh = p.xh + (long long)lane * p.xh_stride + (kAmHist - kAmM);
    const float2 *x0 = p.x0 + (long long)lane * p.n;
    float *mh = p.mh + (long long)lane * p.mh_stride + kAmHist;
    const int s0 = seg * p.pll_L, s1 = min(s0 + p.pll_L, p.n);
    int w0 = s0 - p.pll_W;
    unsigned theta, dtheta = p.pll[2 * lane + 1];
    if (w0 <= 0) { w0 = 0; theta = p.pll[2 * lane]; }                    // continues the previous call: exact
    else {
        // phase of the low-passed carrier: the loop's error Im(x0 e^{-j theta}) starts at zero
        const float2 c0 = x0[w0];
        theta = am_constrain(atan2f(c0.y, c0.x));
    }
    am_pll_run(p, stab, x0, xh, mh, theta, dtheta, w0, s0, false);
    p.seg_start[2 * t] = theta; p.seg_start[2 * t + 1] = dtheta;
    am_pll_run(p, stab, x0, xh, mh, theta, dtheta, s0, s1, true);
    p.seg_end[2 * t] = theta; p.seg_end[2 * t + 1] = dtheta;
}

// one thread per lane: segments whose start state does not continue their predecessor's end state are re-run in stream
// order; the lane's loop state for the next call is the last segment's end state
__global__ void __launch_bounds__(32) k_am_pll_fix(const AmParams p)
{
    __shared__ float stab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) stab[i] = p.sintab[i];
    __syncthreads();
    const int lane = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= p.nlanes) return;
    const float2 *xh = p.xh + (long long)lane * p.xh_stride + (kAmHist - kAmM);
    const float2 *x0 = p.x0 + (long long)lane * p.n;
    float *mh = p.mh + (long long)lane * p.mh_stride + kAmHist;
    unsigned *S = p.seg_start + 2 * (long long)lane * p.pll_nseg, *E = p.seg_end + 2 * (long long)lane * p.pll_nseg;
    unsigned long long redone = 0;
    for (int seg = 1; seg < p.pll_nseg; seg++) {
        unsigned theta = E[2 * (seg - 1)], dtheta = E[2 * (seg - 1) + 1];
        if (am_pll_same(p, S[2 * seg], S[2 * seg + 1], theta, dtheta)) continue;
        const int s0 = seg * p.pll_L, s1 = min(s0 + p.pll_L, p.n);
        am_pll_run(p, stab, x0, xh, mh, theta, dtheta, s0, s1, true);
        E[2 * seg] = theta; E[2 * seg + 1] = dtheta;
        redone++;
    }
    p.pll[2 * lane] = E[2 * (p.pll_nseg - 1)]; p.pll[2 * lane + 1] = E[2 * (p.pll_nseg - 1) + 1];
    if (redone) atomicAdd(p.pll_redone, redone);
}

// y[i] = out_scale * sum_k h_dc[k] * mh[50 + i - k]
__global__ void __launch_bounds__(256) k_am_dcfir(const AmParams p)
{
    __shared__ float tile[kAmTile + kAmHist + (kAmTile + kAmHist) / 4 + 1];
    const int lane = blockIdx.y;
    const float *mh = p.mh + (long long)lane * p.mh_stride;
    float *y = p.y + (long long)lane * p.y_stride;
    float h[kAmTaps];
#pragma unroll
    for (int k = 0; k < kAmTaps; k++) h[k] = p.h_dc[k];
    for (int base = blockIdx.x * kAmTile; base < p.n; base += gridDim.x * kAmTile) {
        __syncthreads();
        for (int j = threadIdx.x; j < kAmTile + kAmHist; j += blockDim.x)
            tile[am_pad(j)] = (base + j < p.n + kAmHist) ? mh[base + j] : 0.f;
        __syncthreads();
        const int o = threadIdx.x * kAmR;
        if (base + o < p.n) {
            float acc[kAmR];
#pragma unroll
            for (int r = 0; r < kAmR; r++) acc[r] = 0.f;
#pragma unroll
            for (int j = 0; j < kAmHist + kAmR; j++) {
                const float v = tile[am_pad(o + j)];
#pragma unroll
                for (int r = 0; r < kAmR; r++) {
                    const int k = r + kAmHist - j;
                    if (k >= 0 && k < kAmTaps) acc[r] += h[k] * v;
                }
            }
#pragma unroll
            for (int r = 0; r < kAmR; r++) if (base + o + r < p.n) y[base + o + r] = acc[r] * p.out_scale;
        }
    }
}

// keep the last 50 samples of xh / mh as the next call's history (n >= 50: plain shift; n < 50: via temp)
__global__ void k_am_tail(const AmParams p, float2 *xtmp, float *mtmp, int phase)
{
    const int lane = blockIdx.x;
    float2 *xh = p.xh + (long long)lane * p.xh_stride;
    float *mh = p.mh + (long long)lane * p.mh_stride;
    const int j = threadIdx.x;
    if (j >= kAmHist) return;
    if (phase == 0) { xtmp[lane * kAmHist + j] = xh[p.n + j]; mtmp[lane * kAmHist + j] = mh[p.n + j]; }
    else            { xh[j] = xtmp[lane * kAmHist + j]; mh[j] = mtmp[lane * kAmHist + j]; }
}

#ifndef CSDR_EMU
// host-side owner of the AM demodulator state (device resident)
struct AmDemod {
    int nlanes = 0; float mod_index = 0.8f; bool use_pll = true;
    void *d_hlp = nullptr, *d_hdc = nullptr, *d_sintab = nullptr, *d_pll = nullptr;
    void *d_xh = nullptr, *d_mh = nullptr, *d_x0 = nullptr, *d_xtmp = nullptr, *d_mtmp = nullptr;
    void *d_seg = nullptr, *d_redone = nullptr; size_t seg_cap = 0;      // time-parallel carrier loop
    bool spec = true;                                                   // false: the sequential loop (k_am_pll), as a cross-check
    size_t cap_n = 0;
    unsigned long long launches = 0;
    unsigned long long take_launches() { unsigned long long l = launches; launches = 0; return l; }

    static void ck(cudaError_t e, const char *what) { if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e)); }
    ~AmDemod() { for (void *p : {d_hlp, d_hdc, d_sintab, d_pll, d_xh, d_mh, d_x0, d_xtmp, d_mtmp, d_seg, d_redone}) if (p) cudaFree(p); }

    void init(cudaStream_t st, int lanes, float mod, bool pll)
    {
        nlanes = lanes; mod_index = mod; use_pll = pll;
        // firfilt_rrrf_create_dc_blocker(25, 20 dB): h = delta[m] - w/sum(w), Kaiser w
        std::vector<float> hdc(kAmTaps), hlp;
        double beta = design::kaiser_beta(20.0f), sum = 0.0;
        std::vector<double> w(kAmTaps);
        for (int i = 0; i < kAmTaps; i++) { w[i] = design::kaiser(i, kAmTaps, beta); sum += w[i]; }
        for (int i = 0; i < kAmTaps; i++) hdc[i] = (float)(-w[i] / sum);
        hdc[kAmM] += 1.0f;
        // firfilt_crcf_create_kaiser(51, 0.01, 40 dB, 0)
        hlp = design::firdes_kaiser(kAmTaps, 0.01f, 40.0f, 0.0f);
        std::vector<float> tab(1024);
        for (int i = 0; i < 1024; i++) tab[i] = sinf((float)(2.0f * design::kPi * (float)i / 1024.0f));
        ck(cudaMalloc(&d_hlp, kAmTaps * 4), "cudaMalloc"); ck(cudaMalloc(&d_hdc, kAmTaps * 4), "cudaMalloc");
        ck(cudaMalloc(&d_sintab, 1024 * 4), "cudaMalloc"); ck(cudaMalloc(&d_pll, (size_t)lanes * 8), "cudaMalloc");
        ck(cudaMalloc(&d_xtmp, (size_t)lanes * kAmHist * 8), "cudaMalloc"); ck(cudaMalloc(&d_mtmp, (size_t)lanes * kAmHist * 4), "cudaMalloc");
        ck(cudaMemcpyAsync(d_hlp, hlp.data(), kAmTaps * 4, cudaMemcpyHostToDevice, st), "copy");
        ck(cudaMemcpyAsync(d_hdc, hdc.data(), kAmTaps * 4, cudaMemcpyHostToDevice, st), "copy");
        ck(cudaMemcpyAsync(d_sintab, tab.data(), 1024 * 4, cudaMemcpyHostToDevice, st), "copy");
        ck(cudaMemsetAsync(d_pll, 0, (size_t)lanes * 8, st), "memset");
        ck(cudaMalloc(&d_redone, 8), "cudaMalloc"); ck(cudaMemsetAsync(d_redone, 0, 8, st), "memset");
        ck(cudaStreamSynchronize(st), "sync");
        ensure(st, 1024);
    }
    void ensure(cudaStream_t st, size_t n)
    {
        if (n <= cap_n) return;
        size_t want = n + n / 4;
        void *nx = nullptr, *nm = nullptr, *n0 = nullptr;
        ck(cudaMalloc(&nx, (size_t)nlanes * (want + kAmHist) * 8), "cudaMalloc");
        ck(cudaMalloc(&nm, (size_t)nlanes * (want + kAmHist) * 4), "cudaMalloc");
        ck(cudaMalloc(&n0, (size_t)nlanes * want * 8), "cudaMalloc");
        ck(cudaMemsetAsync(nx, 0, (size_t)nlanes * (want + kAmHist) * 8, st), "memset");
        ck(cudaMemsetAsync(nm, 0, (size_t)nlanes * (want + kAmHist) * 4, st), "memset");
        if (d_xh) {
            // carry the histories over (they sit at the head of each lane's buffer)
            ck(cudaMemcpy2DAsync(nx, (want + kAmHist) * 8, d_xh, (cap_n + kAmHist) * 8, kAmHist * 8, nlanes, cudaMemcpyDeviceToDevice, st), "copy");
            ck(cudaMemcpy2DAsync(nm, (want + kAmHist) * 4, d_mh, (cap_n + kAmHist) * 4, kAmHist * 4, nlanes, cudaMemcpyDeviceToDevice, st), "copy");
            ck(cudaStreamSynchronize(st), "sync");
            cudaFree(d_xh); cudaFree(d_mh); cudaFree(d_x0);
        }
        d_xh = nx; d_mh = nm; d_x0 = n0; cap_n = want;
    }
    void run(cudaStream_t st, const float2 *x, long long x_stride, float *y, long long y_stride, int n)
    {
        if (n <= 0) return;
        ensure(st, (size_t)n);
        AmParams p{};
        p.nlanes = nlanes; p.n = n; p.x = x; p.x_stride = x_stride;
        p.xh = (float2 *)d_xh; p.xh_stride = (long long)(cap_n + kAmHist);
        p.x0 = (float2 *)d_x0; p.mh = (float *)d_mh; p.mh_stride = (long long)(cap_n + kAmHist);
        p.y = y; p.y_stride = y_stride; p.h_lp = (const float *)d_hlp; p.h_dc = (const float *)d_hdc;
        p.sintab = (const float *)d_sintab; p.pll = (unsigned *)d_pll;
        p.inv_mod = 1.0f / mod_index; p.pll_alpha = 0.001f; p.pll_beta = sqrtf(0.001f);
        p.use_pll = use_pll ? 1 : 0;
        p.out_scale = use_pll ? 1.0f : 1.0f / mod_index;
        int gx = std::max(1, std::min((n + 255) / 256, 2048));
        const int gf = std::max(1, std::min((n + kAmTile - 1) / kAmTile, 2048));      // the FIR kernels: kAmTile outputs per CTA
        k_am_stage_in<<<dim3(gx, nlanes), 256, 0, st>>>(p); launches++;
        if (use_pll) {
            k_am_lowpass<<<dim3(gf, nlanes), 256, 0, st>>>(p); launches++;
            // segments of 2048 samples behind a 512-sample pull-in when there are enough of them to matter
#ifndef CSDR_AM_PLL_L
#define CSDR_AM_PLL_L 2048
#endif
            p.pll_L = CSDR_AM_PLL_L; p.pll_W = 512; p.pll_nseg = (n + p.pll_L - 1) / p.pll_L;
            if (spec && p.pll_nseg >= 4) {
                const size_t need = (size_t)nlanes * p.pll_nseg * 2 * sizeof(unsigned) * 2;
                if (need > seg_cap) { if (d_seg) cudaFree(d_seg); ck(cudaMalloc(&d_seg, need + need / 4), "cudaMalloc"); seg_cap = need + need / 4; }
                p.seg_start = (unsigned *)d_seg; p.seg_end = p.seg_start + (size_t)nlanes * p.pll_nseg * 2;
                p.tol_theta = (unsigned)(1.5e-2 / 6.283185307179586 * 4294967296.0); p.tol_dtheta = (unsigned)(5e-4 / 6.283185307179586 * 4294967296.0);
                p.pll_redone = (unsigned long long *)d_redone;
                const long long nt = (long long)nlanes * p.pll_nseg;
                k_am_pll_spec<<<(unsigned)((nt + 63) / 64), 64, 0, st>>>(p); launches++;
                k_am_pll_fix<<<(nlanes + 31) / 32, 32, 0, st>>>(p); launches++;
            } else {
                k_am_pll<<<(nlanes + 31) / 32, 32, 0, st>>>(p); launches++;
            }
        }
        k_am_dcfir<<<dim3(gf, nlanes), 256, 0, st>>>(p); launches++;
        k_am_tail<<<nlanes, 64, 0, st>>>(p, (float2 *)d_xtmp, (float *)d_mtmp, 0); launches++;
        k_am_tail<<<nlanes, 64, 0, st>>>(p, (float2 *)d_xtmp, (float *)d_mtmp, 1); launches++;
        ck(cudaGetLastError(), "ampmodem launch");
    }
};
#endif

}  // namespace csdr
