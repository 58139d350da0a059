// frontend_plan.hpp -- host-side tile geometry of k_frontend (see frontend.cuh).  Pure host code, shared by
// csdr_b200.cu and the CPU-only emulation test.
#pragma once
#include "frontend.cuh"
#include "design.hpp"
#include <string>

namespace csdr {

struct FrontendGeometry {
    FrontendParams base;     // geometry + taps filled in; per-call fields zero
    int hcap = 0;            // raw-sample history the kernel may reach back over
    size_t smem_bytes = 0;
    std::string error;
};

inline int fe_roundup(int v, int m) { return (v + m - 1) / m * m; }

// Tc: c-samples (half-band cascade outputs) per tile; multiple of 8
inline FrontendGeometry plan_frontend(const design::MsresampPlan &ms, int Tc)
{
    FrontendGeometry g{};
    FrontendParams &p = g.base;
    if (ms.interp) { g.error = "msresamp: interpolation (rate > 1) is not implemented on the GPU path"; return g; }
    if (ms.S > (unsigned)kMaxStages) { g.error = "msresamp: too many half-band stages"; return g; }
    if (2 * ms.m_arb != (unsigned)kHsub) { g.error = "msresamp: unexpected arbitrary-stage length"; return g; }
    p.S = (int)ms.S;
    p.zeta = 1.0f / (float)(1u << ms.S);
    p.step = ms.step; p.bits = (int)ms.bits;
    for (int s = 0; s < p.S; s++) {
        p.m[s] = (int)ms.st[s].m;
        if (p.m[s] > kMaxHbM) { g.error = "msresamp: half-band stage too long"; return g; }
        for (int u = 0; u < 2 * p.m[s]; u++) p.taps[s][u] = ms.st[s].h1[u];
        p.R[s] = (s == 0 && p.S > 1) ? 4 : 8;
    }
    p.Tc = Tc;
    p.n[0] = Tc + kHcPad; p.d[0] = 0;
    for (int L = 0; L < p.S; L++) {
        int need = 2 * p.n[L] + 4 * p.m[L] - 2;
        int mult = 2 * p.R[L];
        if (L + 1 < p.S) mult = std::max(mult, p.R[L + 1]);
        p.n[L + 1] = fe_roundup(need, mult);
        p.d[L + 1] = 2 * p.d[L] + 1 - 4 * p.m[L];
    }
    // shared-memory carve-up: level 0 (c buffer, padded), then two ping-pong regions for levels >= 1
    int size0 = p.n[0] + (p.n[0] >> 3) + 8;
    int sizeA = 0, sizeB = 0;
    for (int L = 1; L <= p.S; L++) {
        int D = p.R[L - 1];
        int st = p.n[L] / (2 * D) + 1;
        if (L == p.S) { while ((st & 15) != 2) st++; }     // loader writes conflict-free (see frontend.cuh)
        p.stride[L] = st;
        int sz = 2 * D * st;
        if (((p.S - L) & 1) == 0) sizeA = std::max(sizeA, sz); else sizeB = std::max(sizeB, sz);
    }
    p.stride[0] = 0;
    p.off[0] = 0;
    size0 = fe_roundup(size0, 2);
    for (int L = 1; L <= p.S; L++) p.off[L] = size0 + ((((p.S - L) & 1) == 0) ? 0 : sizeA);
    if (p.S == 0) { /* loader writes level 0 directly */ }
    int total_f2 = size0 + sizeA + sizeB;
    p.off_bank = 2 * total_f2;
    g.smem_bytes = (size_t)total_f2 * 8 + (size_t)(1u << ms.bits) * (kHsub + 1) * 4;
    p.smem_bytes = (int)g.smem_bytes;
    // history: n0 - lo_S(first tile) <= (2^S - 1) + (kHcPad << S) - d[S]
    g.hcap = fe_roundup(((1 << p.S) - 1) + (kHcPad << p.S) - p.d[p.S] + 1, 64);
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
