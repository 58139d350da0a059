// frontend_plan.hpp -- host-side tile geometry of k_frontend (see frontend.cuh).  Pure host code, shared by
// csdr_b200.cu and the CPU-only emulation test.
#pragma once
#include "frontend.cuh"
#include "frontend_geom.hpp"
#include "design.hpp"
#include <string>

namespace csdr {

struct FrontendGeometry {
    FrontendParams base;     // geometry + taps filled in; per-call fields zero
    FeGeom geom;
    bool std_kernel = false; // k_frontend_std<S, V> applies (compile-time geometry)
    int variant = 0;         // 0: k_frontend_std (register prefetch + loader pass), 1: k_frontend_direct (TMA tile read in place),
                             // 2: k_frontend_ws (the same, warp-specialised: two groups of warps on different tiles)
    int hcap = 0;            // raw-sample history the kernel may reach back over
    size_t smem_bytes = 0;
    std::string error;
};

inline int fe_roundup(int v, int m) { return (v + m - 1) / m * m; }

// Tc: c-samples (half-band cascade outputs) per tile for the generic kernel; multiple of 8.  allow_std: use the
// compile-time geometry (and its own tile size) when the half-band plan is the standard one.
inline FrontendGeometry plan_frontend(const design::MsresampPlan &ms, int Tc, bool allow_std = true, int variant = 0)
{
    FrontendGeometry g{};
    FrontendParams &p = g.base;
    if (ms.interp) { g.error = "msresamp: interpolation (rate > 1) is not implemented on the GPU path"; return g; }
    if (ms.S > (unsigned)kMaxStages) { g.error = "msresamp: too many half-band stages"; return g; }
    if (2 * ms.m_arb != (unsigned)kHsub) { g.error = "msresamp: unexpected arbitrary-stage length"; return g; }
    const int S = (int)ms.S;
    int m[kMaxStages] = {};
    FeStdM stdm{};
    bool is_std = allow_std && S >= 1 && S <= kFeStdMaxS;
    for (int s = 0; s < S; s++) {
        m[s] = (int)ms.st[s].m;
        if (m[s] > kMaxHbM) { g.error = "msresamp: half-band stage too long"; return g; }
        if (m[s] != stdm.v[s]) is_std = false;
    }
    g.std_kernel = is_std;
    g.variant = (is_std && variant == 2 && S >= 2) ? 2 : (is_std && variant != 0) ? 1 : 0;
    g.geom = is_std ? fe_make_geom_std(S, g.variant) : fe_make_geom(S, Tc, m, 0, 0);
    const FeGeom &G = g.geom;
    p.S = S; p.Tc = G.Tc;
    p.zeta = 1.0f / (float)(1u << ms.S);
    p.step = ms.step; p.bits = (int)ms.bits;
    for (int s = 0; s < S; s++) {
        p.m[s] = G.m[s]; p.R[s] = G.R[s];
        for (int u = 0; u < 2 * p.m[s]; u++) p.taps[s][u] = ms.st[s].h1[u];
    }
    for (int L = 0; L <= S; L++) { p.n[L] = G.n[L]; p.d[L] = G.d[L]; p.stride[L] = G.stride[L]; p.off[L] = G.off[L]; }
    p.off_bank = 2 * G.total_f2;
    g.smem_bytes = (size_t)G.total_f2 * 8 + (size_t)(1u << ms.bits) * (kHsub + 1) * 4;
    p.smem_bytes = (int)g.smem_bytes;
    g.hcap = G.hcap;
    return g;
}

}  // namespace csdr

namespace csdr {

// Position of one stream in the front end: everything the sequential liquid objects would hold, in closed form.
struct FrontendCursor {
    unsigned long long n_abs = 0;     // input samples consumed so far (NCO phase and half-band block alignment)
    unsigned long long phase = 0;     // resamp_crcf q->phase: timing phase before the next push, in [0, step)
};

// Fill the per-call fields of p (n0, nx, K0, K1, ntiles, ph0) for a chunk of nx samples, advance the cursor and
// return the number of output samples the call produces (exactly what msresamp_crcf_execute would write).
inline long long fe_prepare_call(const FrontendGeometry &g, FrontendCursor &cur, long long nx, FrontendParams &p)
{
    const int S = g.base.S;
    p.n0 = (long long)cur.n_abs;
    p.nx = nx;
    p.K0 = (long long)(cur.n_abs >> S);
    p.K1 = (long long)((cur.n_abs + (unsigned long long)nx) >> S);
    const unsigned long long pushes = (unsigned long long)(p.K1 - p.K0);
    p.ntiles = (int)((pushes + (unsigned long long)g.base.Tc - 1) / (unsigned long long)g.base.Tc);
    p.ph0 = cur.phase;
    const unsigned long long span = pushes << 24, st = g.base.step;
    const unsigned long long ny = span > cur.phase ? (span - cur.phase + st - 1) / st : 0;
    cur.phase = cur.phase + ny * st - span;
    cur.n_abs += (unsigned long long)nx;
    return (long long)ny;
}

// cursor for a stream whose first n_prior samples were consumed elsewhere (time-segment sharding)
inline FrontendCursor fe_seek(const FrontendGeometry &g, unsigned long long n_prior)
{
    FrontendCursor c; c.n_abs = n_prior;
    const unsigned __int128 span = (unsigned __int128)(n_prior >> g.base.S) << 24;
    const unsigned __int128 st = g.base.step;
    const unsigned __int128 outs = (span + st - 1) / st;
    c.phase = (unsigned long long)(outs * st - span);
    return c;
}

}  // namespace csdr
