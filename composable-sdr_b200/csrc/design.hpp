// design.hpp -- host-side parameter derivation for the B200 blocks: what liquid-dsp v1.3.2 computes inside
// its *_create() functions (filter prototypes, stage plan, phase words), restated here because the kernels
// need the same numbers.  Pure C++ (no CUDA), shared by csdr_b200.cu and the CPU-only tests.
//
// Sources restated (liquid-dsp v1.3.2, not vendored by the reference): src/filter/src/firdes.c,
// src/math/src/windows.c, src/filter/src/{msresamp,msresamp2,resamp2,resamp.fixed,firpfb}.c,
// src/multichannel/src/firpfbch.c, src/nco/src/nco.c, src/agc/src/agc.c.  Reference call sites:
// src/ComposableSDR/Liquid.chs:100-117 (resampler), :782-789 (nco), :811-821 (channelizer), :707-717 (agc).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <complex>
#include <algorithm>

namespace csdr { namespace design {

constexpr double kPi = 3.14159265358979323846;

inline float kaiser_beta(float As)
{
    As = std::fabs(As);
    if (As > 50.0f) return 0.1102f * (As - 8.7f);
    if (As > 21.0f) return 0.5842f * std::pow(As - 21.0f, 0.4f) + 0.07886f * (As - 21.0f);
    return 0.0f;
}

// I0 by its power series (double; liquid uses a 32-term float series: ~1e-6 relative difference)
inline double bessel_i0(double z)
{
    double sum = 1.0, term = 1.0;
    for (int k = 1; k < 200; k++) {
        term *= (z / 2.0) / k;
        double add = term * term;
        sum += add;
        if (add < sum * 1e-20) break;
    }
    return sum;
}

// Kaiser window sample i of N, with r = 2t/(N-1) (the v1.3.2 definition; pinned by the firpfbch tap print)
inline double kaiser(unsigned i, unsigned N, double beta, double mu = 0.0)
{
    double t = (double)i - (double)(N - 1) / 2.0 + mu;
    double r = 2.0 * t / (double)(N - 1);
    double a = std::max(0.0, 1.0 - r * r);
    return bessel_i0(beta * std::sqrt(a)) / bessel_i0(beta);
}

inline double sinc(double x) { return std::fabs(x) < 1e-9 ? 1.0 : std::sin(kPi * x) / (kPi * x); }

// liquid_firdes_kaiser(n, fc, As, mu): sinc(2 fc t) * kaiser, un-normalised
inline std::vector<float> firdes_kaiser(unsigned n, float fc, float As, float mu = 0.0f)
{
    std::vector<float> h(n);
    double beta = kaiser_beta(As);
    for (unsigned i = 0; i < n; i++) {
        double t = (double)i - (double)(n - 1) / 2.0 + mu;
        h[i] = (float)(sinc(2.0 * fc * t) * kaiser(i, n, beta, mu));
    }
    return h;
}

inline unsigned estimate_req_filter_len(float df, float As) { return (unsigned)((As - 7.95f) / (14.26f * df)); }

// NCO(_constrain): radians -> uint32 turn fraction, evaluated in float32 like liquid
inline uint32_t nco_constrain(float theta)
{
    float p = (float)(theta * 0.159154943091895);
    float frac = p - (float)((long)p);
    if (frac < 0.0f) frac = (float)(frac + 1.0);
    float scaled = frac * 4294967296.0f;
    return scaled >= 4294967296.0f ? 0u : (uint32_t)scaled;
}

// ---- msresamp_crcf plan ------------------------------------------------------------------------------
struct HalfbandStage { unsigned m; std::vector<float> h1; };   // h1[u] multiplies odd sample O[q+u], 2m taps
struct MsresampPlan {
    float rate = 0, As = 0;
    bool interp = false;
    unsigned S = 0;                    // half-band stages
    float rate_arb = 0;                // arbitrary stage rate in [0.5, 1) (decim) or (1, 2] (interp)
    std::vector<HalfbandStage> st;     // st[0] = lowest-rate stage (liquid's stage index)
    unsigned npfb = 256, bits = 8, m_arb = 7;
    uint32_t step = 0;                 // round(2^24 / rate_arb)
    std::vector<float> bank;           // [npfb][2*m_arb]; bank[i][j] multiplies the j-th newest sample
};

inline HalfbandStage design_halfband(unsigned m, float As)
{
    // RESAMP2(_create): h[i] = sinc(t/2) kaiser(i, 4m+1, beta(As)), branch taps = odd-indexed, reversed
    HalfbandStage s; s.m = m;
    unsigned n = 4 * m + 1;
    double beta = kaiser_beta(As);
    std::vector<double> h(n);
    for (unsigned i = 0; i < n; i++) h[i] = sinc(((double)i - (double)(n - 1) / 2.0) / 2.0) * kaiser(i, n, beta);
    for (unsigned i = 1; i < n; i += 2) s.h1.push_back((float)h[n - i - 1]);
    return s;
}

inline MsresampPlan plan_msresamp(float r, float As, bool fc_old = false)
{
    MsresampPlan p; p.rate = r; p.As = As;
    p.interp = r > 1.0f;
    p.rate_arb = r;
    if (p.interp) while (p.rate_arb > 2.0f) { p.S++; p.rate_arb *= 0.5f; }
    else          while (p.rate_arb < 0.5f) { p.S++; p.rate_arb *= 2.0f; }
    // MSRESAMP2(_create)(type, S, fc = 0.4, f0 = 0, As): per-stage transition band -> length -> m
    float fc = 0.4f, As_stage = As + 5.0f;
    for (unsigned i = 0; i < p.S; i++) {
        fc = (i == 1) ? (float)((0.5 - fc) / 2.0f) : 0.5f * fc;
        float ft = 2 * (0.25f - fc);
        unsigned h_len = estimate_req_filter_len(ft, As_stage);
        unsigned m = (unsigned)std::ceil((float)(h_len - 1) / 4.0f);
        p.st.push_back(design_halfband(std::max(m, 3u), As_stage));
    }
    // RESAMP(_create)(rate_arb, 7, fc, As, npfb)
    float fca; unsigned npfb;
    if (fc_old) { fca = 0.4f; npfb = 64; } else { fca = std::min(0.515f * p.rate_arb, 0.49f); npfb = 256; }
    p.bits = 0; while ((1u << p.bits) < npfb) p.bits++;
    p.npfb = 1u << p.bits;
    p.step = (uint32_t)std::round((float)(1 << 24) / p.rate_arb);
    unsigned n = 2 * p.m_arb * p.npfb + 1;
    std::vector<float> hf = firdes_kaiser(n, fca / (float)p.npfb, As);
    float gain = 0.0f;
    for (unsigned i = 0; i < n; i++) gain += hf[i];
    gain = (float)p.npfb / gain;
    unsigned hs = 2 * p.m_arb;
    p.bank.resize((size_t)p.npfb * hs);
    for (unsigned i = 0; i < p.npfb; i++)
        for (unsigned j = 0; j < hs; j++) p.bank[(size_t)i * hs + j] = hf[i + j * p.npfb] * gain;
    return p;
}

// ---- firpfbch_crcf_create_kaiser(ANALYZER, M, m, As): prototype, first 2*M*m taps used ---------------
inline std::vector<float> design_firpfbch(unsigned M, unsigned m, float As)
{
    std::vector<float> h = firdes_kaiser(2 * M * m + 1, 0.5f / (float)M, std::fabs(As));
    h.resize((size_t)2 * M * m);
    return h;
}
// firpfbch2_crcf_create_kaiser (analyzer): 2 M m + 1 taps at fc = 1/M, scaled to sum M, the first 2 M m used
inline std::vector<float> design_firpfbch2(unsigned M, unsigned m, float As)
{
    std::vector<float> h = firdes_kaiser(2 * M * m + 1, 1.0f / (float)M, As);
    float sum = 0.0f;
    for (float v : h) sum += v;
    for (float &v : h) v = v * (float)M / sum;
    h.resize((size_t)2 * M * m);
    return h;
}
// reference pre-rotation frequency, evaluated in float32 like the Haskell expression (Liquid.chs:817)
inline float firpfbch_rotation(unsigned C)
{
    float off = 0.5f * ((float)C - 1.0f);
    off = off / (float)C;
    off = off * 2.0f;
    off = off * (float)kPi;
    return -off;
}

// ---- iirfilt_rrrf_create_prototype: Butterworth low-pass as second-order sections ---------------------
// liquid_iirdes for (LIQUID_IIRDES_BUTTER, LOWPASS, SOS), all in float32 like liquid: analog poles on the unit
// circle, pre-warp m = tan(pi fc), bilinear map z = (1 + m s) / (1 - m s) with all zeros at z = -1, one section
// per conjugate pole pair (a real pole last), the gain spread evenly over the sections' numerators.
// Reference: iirfiltCreate (Liquid.chs:629-633); wbFMDemodulator uses order 2, fc = 5000 / quadRate.
struct Sos { float b[3], a[3]; };
inline std::vector<Sos> butter_lowpass_sos(unsigned order, float fc)
{
    typedef std::complex<float> cf32;
    const unsigned r = order % 2, L = (order - r) / 2;
    std::vector<cf32> pd;
    const float m = std::tan((float)kPi * fc);
    cf32 G(1.0f, 0.0f);
    auto map_pole = [&](cf32 pa) {
        const cf32 pm = pa * m, one(1.0f, 0.0f);
        const cf32 z = (one + pm) / (one - pm);
        pd.push_back(z);
        G *= (one - z) / cf32(2.0f, 0.0f);          // (1 - p) / (1 - zero), zero = -1
    };
    for (unsigned i = 0; i < L; i++) {
        const float th = (float)(2 * (i + 1) + order - 1) * (float)kPi / (float)(2 * order);
        map_pole(cf32(std::cos(th), std::sin(th)));
        map_pole(cf32(std::cos(th), -std::sin(th)));
    }
    if (r) map_pole(cf32(-1.0f, 0.0f));
    const float k = std::pow(G.real(), 1.0f / (float)(L + r));
    std::vector<Sos> out;
    for (unsigned i = 0; i < L; i++) {
        const cf32 p0 = -pd[2 * i], p1 = -pd[2 * i + 1];
        Sos s;
        s.a[0] = 1.0f; s.a[1] = (p0 + p1).real(); s.a[2] = (p0 * p1).real();
        s.b[0] = k; s.b[1] = 2.0f * k; s.b[2] = k;
        out.push_back(s);
    }
    if (r) {
        Sos s;
        s.a[0] = 1.0f; s.a[1] = -pd[order - 1].real(); s.a[2] = 0.0f;
        s.b[0] = k; s.b[1] = k; s.b[2] = 0.0f;
        out.push_back(s);
    }
    return out;
}

// ---- firdecim_rrrf_create_kaiser: 2 M m + 1 taps, fc = 0.5 / M, stored reversed (firdecim_create) ----------
inline std::vector<float> design_firdecim(unsigned M, unsigned m, float As)
{
    std::vector<float> h = firdes_kaiser(2 * M * m + 1, 0.5f / (float)M, As, 0.0f);
    std::reverse(h.begin(), h.end());
    return h;
}

// ---- agc: smallest gain for which rssi = -20 log10(g) is NOT above the threshold ---------------------
inline float agc_gain_threshold(float thr_db)
{
    auto exceeded = [&](float g) { return (float)(-20 * std::log10((double)g)) > thr_db; };
    // bisect on the bit pattern of positive floats (monotone)
    uint32_t lo = 0x00800000u, hi = 0x7f7fffffu;      // exceeded(lo) expected true, exceeded(hi) false
    auto f = [](uint32_t b) { float x; std::memcpy(&x, &b, 4); return x; };
    if (!exceeded(f(lo))) return f(lo);
    if (exceeded(f(hi))) return INFINITY;
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (exceeded(f(mid))) lo = mid; else hi = mid;
    }
    return f(hi);
}

}}  // namespace csdr::design
