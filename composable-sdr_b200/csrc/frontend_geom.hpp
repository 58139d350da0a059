// frontend_geom.hpp -- tile geometry of the fused front end, as ONE constexpr function used both at compile time
// (k_frontend_std<S>: every shared-memory address is an immediate) and at run time (generic kernel, host planning).
//
// Levels: level 0 = output of the half-band cascade ("c" samples, input of the arbitrary resampler);
// level L >= 1 = input of half-band stage L-1; level S = the (mixed) raw input samples.
// Tile t produces c indices [kA, kA+Tc); level L holds n[L] samples starting at absolute index
// lo_L = (kA - kHcPad) * 2^L + d[L].  Half-band stage s (liquid index; s = 0 is the lowest rate) computes
//   out[q] = in[2(q+m)+sh] + sum_{u<2m} h1[u] * in[2(q+u)+1+sh]            sh = shift of its input level
// which is y[k] = sum_i h[i] x[2k+1-i] of resamp2.c re-indexed to the tile.
#pragma once
#include "platform.cuh"

namespace csdr {

#ifndef CSDR_FE_NT
#define CSDR_FE_NT 384
#endif
constexpr int kFeNT = CSDR_FE_NT;   // threads per CTA of the compile-time-geometry kernels

struct FeGeom {
    int S = 0, Tc = 0, shift = 0;
    int m[kMaxStages] = {}, R[kMaxStages] = {};
    int d[kMaxStages + 1] = {}, n[kMaxStages + 1] = {}, stride[kMaxStages + 1] = {}, off[kMaxStages + 1] = {};
    int total_f2 = 0;     // float2 elements of all level buffers
    int off_alt = 0;      // raw = 2: second copy of level S-1 (producer / consumer double buffer)
    int hcap = 0;         // raw-sample history the first tile of a chunk may reach back over
};

__host__ __device__ constexpr int ce_max(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int ce_roundup(int v, int m) { return (v + m - 1) / m * m; }

// shift = 1: the top level starts one sample early so that pairs (2p, 2p+1) are 16-byte aligned when the chunk is.
// Sub-array strides are == 2 (mod 16): 64-bit shared accesses are served per half-warp, and with this stride both the
// loader (pair p -> sub-array p & 7, index p >> 3) and the R = 8 producers (t -> sub-array 4 (t & 1) + c, index t >> 1)
// touch 16 distinct 8-byte banks per half-warp.
__host__ __device__ constexpr FeGeom fe_make_geom(int S, int Tc, const int *m, int shift, int raw)
{
    FeGeom g{};
    g.S = S; g.Tc = Tc; g.shift = S > 0 ? shift : 0;
    // outputs per thread slot: 8 everywhere but the last (longest, lowest-rate) stage.  Fewer outputs per slot keep
    // more threads busy but re-read shared memory more often, and shared-memory bandwidth is what binds this kernel
    // (measured: slots 8/4/2 -> 236 us, 8/8/4 -> 227 us per 2^26 samples).
#ifndef CSDR_FE_R0          // slot-width experiments (scripts/exp_build.sh): last stage / middle stages
#define CSDR_FE_R0 4
#endif
#ifndef CSDR_FE_R1
#define CSDR_FE_R1 8
#endif
    for (int s = 0; s < S; s++) { g.m[s] = m[s]; g.R[s] = (s == 0 && S > 1) ? CSDR_FE_R0 : (s == S - 1) ? 8 : CSDR_FE_R1; }
    // raw = 1: the top level is the raw tile itself as the TMA tensor copy delivers it: whole rows of 16 samples
    // (128 bytes), 128-byte swizzle, at the start of shared memory (1024-byte aligned).  The first half-band stage
    // reads (even, odd) pairs from it with 16-byte loads, 8 outputs per thread slot = one row per slot.
    // raw = 2 (k_frontend_ws): the same, and level S-1 -- the hand-over between the two warp groups -- is double-buffered
    // in a region of its own
    const int direct = ((raw == 1 || raw == 2) && S > 0) ? 1 : 0;
    const int ws = (raw == 2 && S > 1) ? 1 : 0;
    g.n[0] = Tc + kHcPad; g.d[0] = 0;
    for (int L = 0; L < S; L++) {
        const int sh = (L + 1 == S) ? g.shift : 0;
        const int need = 2 * g.n[L] + 4 * g.m[L] - 2 + sh;
        int mult = 2 * g.R[L];
        if (L + 1 < S) mult = ce_max(mult, g.R[L + 1]);
        if (direct && L + 1 == S) mult = (need > 256 * kFeRawRow) ? 2 * kFeRawRow : kFeRawRow;   // one or two TMA boxes of <= 256 rows
        g.n[L + 1] = ce_roundup(need, mult);
        g.d[L + 1] = 2 * g.d[L] + 1 - 4 * g.m[L] - sh;
    }
    // the staging buffer is private (the next tile is copied into it while the lower stages of the current tile run);
    // the de-interleaved levels ping-pong between two regions
    const int priv = direct;
    int sizeA = 0, sizeB = 0, sizeT = 0, sizeP = 0;
    for (int L = 1; L <= S; L++) {
        if (direct && L == S) {
            sizeT = ce_roundup(g.n[S], 128);                        // whole 1024-byte swizzle atoms
            continue;
        }
        const int D = g.R[L - 1];
        int st = g.n[L] / (2 * D) + 1;
        while ((st & 15) != 2) st++;
        g.stride[L] = st;
        const int sz = ce_roundup(2 * D * st + kFePlanePad, 2);
        if (priv && L == S) sizeT = sz;
        else if (ws && L == S - 1) sizeP = sz;
        else if (((S - L) & 1) == 0) sizeA = ce_max(sizeA, sz); else sizeB = ce_max(sizeB, sz);
    }
    const int size0 = ce_roundup(g.n[0] + 2, 2);
    g.off[0] = priv ? sizeT : 0;                                   // the swizzled tile comes first (alignment)
    for (int L = 1; L <= S; L++) g.off[L] = size0 + sizeT + 2 * sizeP + ((((S - L) & 1) == 0) ? 0 : sizeA);
    if (priv) g.off[S] = 0;
    if (ws) { g.off[S - 1] = size0 + sizeT; g.off_alt = size0 + sizeT + sizeP; }
    g.total_f2 = size0 + sizeT + 2 * sizeP + sizeA + sizeB;
    g.hcap = ce_roundup(((1 << S) - 1) + (kHcPad << S) - g.d[S] + 1, 64);
    return g;
}

// the half-band plan msresamp_crcf_create(r, 60 dB) always produces (As + 5 = 65 dB per stage): m = 10, 5, 3, 3, ...
struct FeStdM { int v[kMaxStages] = {10, 5, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3}; };
__host__ __device__ constexpr int fe_std_tc(int S)
{
    return S == 1 ? 1792 : S == 2 ? 896 : S == 3 ? 400 : S == 4 ? 208 : S == 5 ? 96 : 48;
}
// variant 1 (k_frontend_direct): TMA tensor copy into a swizzled tile that IS the top level; the first stage reads
// pairs from it and mixes in registers.  Tiles of ~6100 raw samples (two CTAs of 384 threads per SM, 80 registers):
// the tile is as large as the first stage can finish in ONE round of thread slots (n[S-1] / 8 <= 384) and the
// resampler in one round of push pairs (Tc / 2 <= 384), so no warp runs two slots back to back while others idle.
// Measured per 2^27 samples: 3440-sample tiles, three CTAs of 256: 328 us; 6528-sample tiles, two CTAs of 256
// (two rounds in the first stage): 315; 6144-sample tiles, two CTAs of 384: 300; 416 threads with 6528: 307;
// 352 with 5632: 313; narrower slots in the lower stages (more threads busy, more shared-memory reads): 313-322.
__host__ __device__ constexpr int fe_std_tc_direct(int S)
{
#ifdef CSDR_FE_TC3      // tile-size experiments: output samples per tile for S = 3 (scaled for the other S)
    return S == 1 ? 4 * CSDR_FE_TC3 : S == 2 ? 2 * CSDR_FE_TC3 : S == 3 ? CSDR_FE_TC3 : S == 4 ? CSDR_FE_TC3 / 2 : S == 5 ? 64 : 32;
#else
    return S == 1 ? 2880 : S == 2 ? 1440 : S == 3 ? 720 : S == 4 ? 336 : S == 5 ? 144 : 32;
#endif
}
// variant 2 (k_frontend_ws): smaller tiles so that three CTAs with the double-buffered hand-over level fit an SM
__host__ __device__ constexpr int fe_std_tc_ws(int S)
{
    return S == 2 ? 640 : S == 3 ? 320 : S == 4 ? 160 : S == 5 ? 64 : 32;
}
__host__ __device__ constexpr FeGeom fe_make_geom_std(int S, int variant = 0)
{
    FeStdM mm{};
    return variant == 2 ? fe_make_geom(S, fe_std_tc_ws(S), mm.v, 1, 2)
         : variant ? fe_make_geom(S, fe_std_tc_direct(S), mm.v, 1, 1) : fe_make_geom(S, fe_std_tc(S), mm.v, 1, 0);
}
constexpr int kFeStdMaxS = 6;      // both kernels are instantiated for S = 1..6

}  // namespace csdr
