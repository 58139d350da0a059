// wbfm.cuh -- the real-valued tail of the wide-band FM demodulator:
//   iirfilt_rrrf second-order sections (de-emphasis; reference iirFilter, Liquid.chs:629-650) and
//   firdecim_rrrf (output decimator; reference firDecimator, Liquid.chs:487-503),
// as wbFMDemodulator composes them after freqdem (Liquid.chs:652-656).
//
// The recursion of a section is  s' = A s + (x, 0),  s = (v1, v2),  A = [[-a1, -a2], [1, 0]]  -- an affine map per
// sample, so a chunk is filtered in three passes of independent work:
//   k_iir2_seg    a warp per segment of 256 samples: every lane runs its 8 samples from the zero state (fp64), a
//                 warp scan composes the lanes' maps (the multipliers are the constants A^(8 2^s))
//   k_iir2_carry  a warp per stream lane: the segments' maps are composed into the state on entry of every segment
//                 (every lane a run of segments, one warp scan between two sequential sweeps)
//   k_iir2_apply  a warp per segment again: entry state of every lane = A^(8 l) (segment entry) + scanned prefix,
//                 then the samples are filtered in float32 with the arithmetic of liquid's direct form II
// so the float32 rounding of every output sample is the sequential filter's; only the carried state comes from the
// fp64 scan (difference ~1e-7 relative, decaying with the filter's memory).
#pragma once
#include "platform.cuh"
#include <algorithm>
#include <cmath>

namespace csdr {

constexpr int kIirRL = 8;                  // samples per lane
constexpr int kIirSeg = 32 * kIirRL;       // samples per warp segment
constexpr int kIirWarps = 4;               // segments per CTA

struct Iir2Params {
    const float *x; long long x_stride;    // [lanes][n]
    float *y; long long y_stride;
    int n, nseg;
    float b0, b1, b2, a1, a2;              // section coefficients, a0 = 1
    double P[5][4];                        // A^(kIirRL 2^s), s = 0..4
    const double *Q;                       // device table [32][4]: A^(kIirRL l)
    double AS[4];                          // A^kIirSeg
    int K;                                 // segments per lane in k_iir2_carry
    double PK[5][4];                       // A^(kIirSeg K 2^s)
    double2 *seg_z;                        // [lanes][nseg] zero-state response at the end of a segment
    double2 *seg_in;                       // [lanes][nseg] state on entry of a segment
    float2 *state;                         // [lanes] (v1, v2) carried between calls
};

__device__ __forceinline__ double2 iir2_mv(const double (&m)[4], double2 z)
{
    return make_double2(m[0] * z.x + m[1] * z.y, m[2] * z.x + m[3] * z.y);
}

// zero-state response of this lane's samples, then the inclusive warp scan: lane l ends up with the zero-state
// response at the end of lane l's samples counted from the start of the segment
__device__ __forceinline__ double2 iir2_lane_scan(const Iir2Params &p, const float (&v)[kIirRL])
{
    const int l = threadIdx.x & 31;
    double2 z = make_double2(0.0, 0.0);
    const double a1 = (double)p.a1, a2 = (double)p.a2;
#pragma unroll
    for (int q = 0; q < kIirRL; q++) z = make_double2((double)v[q] - a1 * z.x - a2 * z.y, z.x);
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int d = 1 << s;
        const double px = __shfl_up_sync(0xffffffffu, z.x, d), py = __shfl_up_sync(0xffffffffu, z.y, d);
        if (l >= d) {
            const double2 t = iir2_mv(p.P[s], make_double2(px, py));
            z.x += t.x; z.y += t.y;
        }
    }
    return z;
}

__device__ __forceinline__ void iir2_load(const float *__restrict__ x, int n, int i0, float (&v)[kIirRL])
{
#pragma unroll
    for (int q = 0; q < kIirRL; q++) v[q] = (i0 + q < n) ? x[i0 + q] : 0.0f;
}

__global__ void __launch_bounds__(32 * kIirWarps) k_iir2_seg(const Iir2Params p)
{
    const int lane = blockIdx.y, sg = blockIdx.x * kIirWarps + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (sg >= p.nseg) return;
    float v[kIirRL];
    iir2_load(p.x + (long long)lane * p.x_stride, p.n, sg * kIirSeg + l * kIirRL, v);
    const double2 z = iir2_lane_scan(p, v);
    if (l == 31) p.seg_z[(long long)lane * p.nseg + sg] = z;
}

__global__ void __launch_bounds__(32) k_iir2_carry(const Iir2Params p)
{
    const int lane = blockIdx.x, l = threadIdx.x;
    const double2 *sz = p.seg_z + (long long)lane * p.nseg;
    double2 *si = p.seg_in + (long long)lane * p.nseg;
    const int k0 = l * p.K, k1 = min(p.nseg, k0 + p.K);
    // lane 0 starts from the carried state, so the scan hands every lane the true state at the start of its run
    const float2 st = p.state[lane];
    double2 z = (l == 0) ? make_double2((double)st.x, (double)st.y) : make_double2(0.0, 0.0);
    for (int k = k0; k < k0 + p.K; k++) {
        const double2 t = iir2_mv(p.AS, z);
        const double2 r = (k < k1) ? sz[k] : make_double2(0.0, 0.0);
        z = make_double2(t.x + r.x, t.y + r.y);
    }
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int d = 1 << s;
        const double px = __shfl_up_sync(0xffffffffu, z.x, d), py = __shfl_up_sync(0xffffffffu, z.y, d);
        if (l >= d) {
            const double2 t = iir2_mv(p.PK[s], make_double2(px, py));
            z.x += t.x; z.y += t.y;
        }
    }
    double ex = __shfl_up_sync(0xffffffffu, z.x, 1), ey = __shfl_up_sync(0xffffffffu, z.y, 1);
    if (l == 0) { ex = (double)st.x; ey = (double)st.y; }
    double2 s = make_double2(ex, ey);
    for (int k = k0; k < k1; k++) {
        si[k] = s;
        const double2 t = iir2_mv(p.AS, s), r = sz[k];
        s = make_double2(t.x + r.x, t.y + r.y);
    }
}

__global__ void __launch_bounds__(32 * kIirWarps) k_iir2_apply(const Iir2Params p)
{
    const int lane = blockIdx.y, sg = blockIdx.x * kIirWarps + (threadIdx.x >> 5), l = threadIdx.x & 31;
    if (sg >= p.nseg) return;
    const int i0 = sg * kIirSeg + l * kIirRL;
    float v[kIirRL];
    iir2_load(p.x + (long long)lane * p.x_stride, p.n, i0, v);
    const double2 z = iir2_lane_scan(p, v);
    double ex = __shfl_up_sync(0xffffffffu, z.x, 1), ey = __shfl_up_sync(0xffffffffu, z.y, 1);
    if (l == 0) { ex = 0.0; ey = 0.0; }
    double q4[4];
#pragma unroll
    for (int k = 0; k < 4; k++) q4[k] = p.Q[l * 4 + k];
    const double2 e = iir2_mv(q4, p.seg_in[(long long)lane * p.nseg + sg]);
    // liquid iirfiltsos_execute_df2 in float32, products and sums rounded one by one
    float v1 = (float)(e.x + ex), v2 = (float)(e.y + ey);
    float *y = p.y + (long long)lane * p.y_stride;
#pragma unroll
    for (int q = 0; q < kIirRL; q++) {
        if (i0 + q < p.n) {
            const float v0 = __fsub_rn(__fsub_rn(v[q], __fmul_rn(p.a1, v1)), __fmul_rn(p.a2, v2));
            y[i0 + q] = __fadd_rn(__fadd_rn(__fmul_rn(p.b0, v0), __fmul_rn(p.b1, v1)), __fmul_rn(p.b2, v2));
            v2 = v1; v1 = v0;
            if (i0 + q == p.n - 1) p.state[lane] = make_float2(v1, v2);
        }
    }
}

// firdecim_rrrf: y[j] = sum_i h[i] z[j M + i]  (h = reversed prototype, z = history | pending | new samples), summed
// in liquid's order with separately rounded products
struct FirDecimParams {
    const float *z; long long z_stride;
    float *y; long long y_stride;
    const float *h;
    int nout, M, Lh;
};
__global__ void __launch_bounds__(256) k_firdecim(const FirDecimParams p)
{
    CSDR_DYN_SMEM(smem_raw);
    float *hs = reinterpret_cast<float *>(smem_raw);
    for (int i = threadIdx.x; i < p.Lh; i += blockDim.x) hs[i] = p.h[i];
    __syncthreads();
    const int lane = blockIdx.y;
    const float *z = p.z + (long long)lane * p.z_stride;
    float *y = p.y + (long long)lane * p.y_stride;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.nout; j += gridDim.x * blockDim.x) {
        const float *w = z + (long long)j * p.M;
        float r = 0.0f;
        for (int i = 0; i < p.Lh; i++) r = __fadd_rn(r, __fmul_rn(hs[i], w[i]));
        y[j] = r;
    }
}

// dst[lane][doff + i] = src[lane][soff + i], i < count
__global__ void k_rows_copy(const float *__restrict__ src, long long sstride, long long soff, float *__restrict__ dst,
                            long long dstride, long long doff, int count)
{
    const int lane = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
        dst[(long long)lane * dstride + doff + i] = src[(long long)lane * sstride + soff + i];
}

// ---- host side: constants of the scans, launch sequences (shared with the test-only CPU emulation) ------------
inline void iir2_mat_mul(const double (&a)[4], const double (&b)[4], double (&c)[4])
{
    const double r[4] = {a[0] * b[0] + a[1] * b[2], a[0] * b[1] + a[1] * b[3], a[2] * b[0] + a[3] * b[2], a[2] * b[1] + a[3] * b[3]};
    for (int i = 0; i < 4; i++) c[i] = r[i];
}
inline void iir2_mat_pow(const double (&a)[4], unsigned long long e, double (&out)[4])
{
    double r[4] = {1.0, 0.0, 0.0, 1.0}, b[4] = {a[0], a[1], a[2], a[3]};
    while (e) {
        if (e & 1) iir2_mat_mul(r, b, r);
        iir2_mat_mul(b, b, b);
        e >>= 1;
    }
    for (int i = 0; i < 4; i++) out[i] = r[i];
}
// q[32][4] = A^(kIirRL l): goes into a device table once per section
inline void iir2_q_table(float a1, float a2, double *q)
{
    const double A[4] = {-(double)a1, -(double)a2, 1.0, 0.0};
    for (int l = 0; l < 32; l++) {
        double m[4];
        iir2_mat_pow(A, (unsigned long long)kIirRL * l, m);
        for (int k = 0; k < 4; k++) q[l * 4 + k] = m[k];
    }
}
// everything of Iir2Params that depends on the coefficients and on the chunk length
inline void iir2_plan(Iir2Params &p, const float (&b)[3], const float (&a)[3], int n)
{
    p.b0 = b[0]; p.b1 = b[1]; p.b2 = b[2]; p.a1 = a[1]; p.a2 = a[2];
    p.n = n; p.nseg = (n + kIirSeg - 1) / kIirSeg;
    p.K = std::max(1, (p.nseg + 31) / 32);
    const double A[4] = {-(double)a[1], -(double)a[2], 1.0, 0.0};
    for (int s = 0; s < 5; s++) {
        iir2_mat_pow(A, (unsigned long long)kIirRL << s, p.P[s]);
        iir2_mat_pow(A, ((unsigned long long)kIirSeg * p.K) << s, p.PK[s]);
    }
    iir2_mat_pow(A, kIirSeg, p.AS);
}
template <class Launch>
inline void iir2_launch(Launch &&launch, const Iir2Params &p, int lanes)
{
    if (p.n <= 0) return;
    const dim3 g((p.nseg + kIirWarps - 1) / kIirWarps, lanes);
    launch(k_iir2_seg, g, dim3(32 * kIirWarps), 0, p);
    launch(k_iir2_carry, dim3(lanes), dim3(32), 0, p);
    launch(k_iir2_apply, g, dim3(32 * kIirWarps), 0, p);
}
template <class Launch>
inline void firdecim_launch(Launch &&launch, const FirDecimParams &p, int lanes, int max_ctas)
{
    if (p.nout <= 0) return;
    const int g = std::max(1, std::min((p.nout + 255) / 256, max_ctas));
    launch(k_firdecim, dim3(g, lanes), dim3(256), (size_t)p.Lh * sizeof(float), p);
}

}  // namespace csdr
