// emu_kernels.cpp -- TEST-ONLY: runs the real kernel sources (composable-sdr_b200/csrc/*.cuh) under the CPU
// thread emulator (cuda_emu.h) so tile geometry / state carry can be checked against the oracle without a GPU.
// Built by tests/emu/build.py with g++ -DCSDR_EMU; never part of libcsdr_b200.so.
#include "cuda_emu.h"
#include "frontend.cuh"
#include "frontend_std.cuh"
#include "frontend_plan.hpp"
#include "interp.cuh"
#include "backend.cuh"
#include "pfb.cuh"
#include "wbfm.cuh"
#include <vector>
#include <cstdio>

using namespace csdr;

struct EmuLaunch {
    template <class K, class... A> void operator()(K k, dim3 g, dim3 b, size_t smem, A... a) const { csdr_emu::launch(g, b, smem, k, a...); }
    void debug_after_verify(const BackendParams &) const {}
    // cooperative kernel: CTAs run one after the other here, so the phases (separated by grid barriers on the GPU) are
    // launched one by one; every phase decides for itself whether it has anything to do
    template <class K> void coop(K k, dim3 b, const BackendParams &p, int lo, int hi) const
    {
        for (int ph = lo; ph <= hi; ph++) csdr_emu::launch(dim3(3), b, 0, k, p, ph, ph);
    }
};

extern "C" {

// mix (mode 0/1/2, freq in radians/sample) + msresamp(rate, As), fed in the given chunk sizes.
// Returns total outputs written to y (capacity cap), or -1.
long long emu_frontend(float rate, float As, int mix_mode, float freq, int quantize, int Tc, int nthreads,
                       const float2 *x, long long n, const long long *chunks, int nchunks, float2 *y, long long cap,
                       unsigned long long seek, int allow_std, long long misalign)
{
    design::MsresampPlan ms = design::plan_msresamp(rate, As);
    if (ms.interp) {
        // rate > 1: arbitrary stage + half-band interpolators (interp.cuh), same launch sequence as the product
        InterpPlan ip;
        ip.S = (int)ms.S; ip.step = ms.step; ip.bits = (int)ms.bits;
        for (int st = 0; st < ip.S; st++) { ip.m[st] = (int)ms.st[st].m; for (int u = 0; u < 2 * ip.m[st]; u++) ip.h1[st][u] = ms.st[st].h1[u]; }
        std::vector<float2> hist[2], xm;
        hist[0].assign(kInterpHcap, make_float2(0, 0)); hist[1] = hist[0];
        std::vector<float2> scratch[2];
        int cur_h = 0;
        unsigned long long n_abs = seek;
        long long pos = 0, total = 0;
        EmuLaunch launch;
        const unsigned dth = design::nco_constrain(freq);
        for (int c = 0; c < nchunks; c++) {
            const long long nx = chunks[c];
            if (pos + nx > n) return -1;
            if (nx == 0) continue;
            const float2 *xc = x + pos;
            if (mix_mode) {
                xm.resize((size_t)nx);
                launch(k_nco_mix, dim3(4), dim3(64), 0, xc, xm.data(), nx, (unsigned)n_abs * dth, dth, quantize, mix_mode == 2 ? 1 : 0);
                xc = xm.data();
            }
            if (total + interp_max_out(ip, nx) > cap) return -1;
            auto buf = [&](int slot, size_t bytes) -> void * { scratch[slot].resize(bytes / sizeof(float2) + 1); return scratch[slot].data(); };
            const long long ny = interp_launch(launch, buf, ip, ms.bank.data(), 1, xc, 0, hist[cur_h].data(), n_abs, nx, y + total, 0);
            launch(k_hist_update, dim3((kInterpHcap + 127) / 128), dim3(128), 0, (const float2 *)hist[cur_h].data(),
                   hist[cur_h ^ 1].data(), xc, 0LL, nx, kInterpHcap);
            cur_h ^= 1;
            n_abs += (unsigned long long)nx; pos += nx; total += ny;
        }
        return total;
    }
    FrontendGeometry g = plan_frontend(ms, Tc, allow_std != 0, allow_std >= 2 ? allow_std - 1 : 0);
    void (*kernel)(FrontendParams) = k_frontend;
    void (*kernel_direct)(FrontendParams, FeTmap) = nullptr;
    if (g.std_kernel) {
        nthreads = g.variant == 2 ? 2 * kFeWsGroup : kFeNT;
        if (g.variant == 0) {
            switch (ms.S) {
            case 1: kernel = k_frontend_std<1>; break;
            case 2: kernel = k_frontend_std<2>; break;
            case 3: kernel = k_frontend_std<3>; break;
            case 4: kernel = k_frontend_std<4>; break;
            case 5: kernel = k_frontend_std<5>; break;
            default: kernel = k_frontend_std<6>; break;
            }
        } else if (g.variant == 2) {
            switch (ms.S) {
            case 2: kernel_direct = k_frontend_ws<2>; break;
            case 3: kernel_direct = k_frontend_ws<3>; break;
            case 4: kernel_direct = k_frontend_ws<4>; break;
            case 5: kernel_direct = k_frontend_ws<5>; break;
            default: kernel_direct = k_frontend_ws<6>; break;
            }
        } else {
            switch (ms.S) {
            case 1: kernel_direct = k_frontend_direct<1>; break;
            case 2: kernel_direct = k_frontend_direct<2>; break;
            case 3: kernel_direct = k_frontend_direct<3>; break;
            case 4: kernel_direct = k_frontend_direct<4>; break;
            case 5: kernel_direct = k_frontend_direct<5>; break;
            default: kernel_direct = k_frontend_direct<6>; break;
            }
        }
    }
    // the caller's buffer is copied to a 16-byte aligned (or deliberately misaligned) one: exercises both loaders
    std::vector<float2> xbuf((size_t)n + 4);
    float2 *xa = (float2 *)(((uintptr_t)xbuf.data() + 15) & ~(uintptr_t)15) + misalign;
    memcpy(xa, x, (size_t)n * sizeof(float2));
    x = xa;
    if (!g.error.empty()) { fprintf(stderr, "%s\n", g.error.c_str()); return -1; }
    std::vector<float2> hist[2];
    hist[0].assign(g.hcap, make_float2(0, 0)); hist[1] = hist[0];
    int cur_h = 0;
    FrontendCursor cur = seek ? fe_seek(g, seek) : FrontendCursor();
    long long pos = 0, total = 0;
    for (int c = 0; c < nchunks; c++) {
        long long nx = chunks[c];
        if (pos + nx > n) return -1;
        FrontendParams p = g.base;
        long long ny = fe_prepare_call(g, cur, nx, p);
        if (total + ny > cap) return -1;
        p.x = x + pos; p.hist = hist[cur_h].data(); p.y = y + total; p.hcap = g.hcap;
        p.x_stride = 0; p.y_stride = 0;
        p.mix_mode = mix_mode; p.theta0 = 0; p.dtheta = design::nco_constrain(freq); p.quantize = quantize;
        p.bank = ms.bank.data();
        if (p.ntiles > 0 && kernel_direct) {
            // what Frontend::make_tensor_map does: rows of 16 samples from sample r on, 16-byte aligned base
            FeTmap tm{};
            const FeGeom &G = g.geom;
            const long long lo0 = (p.K0 - kHcPad) * (1LL << G.S) + G.d[G.S];
            const long long r = (((lo0 - p.n0) % kFeRawRow) + kFeRawRow) % kFeRawRow;
            p.tma_ok = 0; p.tma_r = 0; p.tma_rows = 0;
            if (nx >= kFeRawRow && (nx - r) / kFeRawRow >= 1 && ((uintptr_t)(p.x + r) & 15) == 0) {
                tm.base = p.x + r; tm.nrows = (nx - r) / kFeRawRow; tm.stream_stride = 0;
                p.tma_ok = 1; p.tma_r = (int)r; p.tma_rows = tm.nrows;
            }
            csdr_emu::launch(dim3(std::min(p.ntiles, 3)), dim3(nthreads), g.smem_bytes, kernel_direct, p, tm);
        } else if (p.ntiles > 0)
            csdr_emu::launch(dim3(std::min(p.ntiles, 3)), dim3(nthreads), g.smem_bytes, kernel, p);
        csdr_emu::launch(dim3((g.hcap + 127) / 128), dim3(128), 0, k_hist_update, (const float2 *)hist[cur_h].data(),
                         hist[cur_h ^ 1].data(), x + pos, 0LL, nx, g.hcap);
        cur_h ^= 1;
        pos += nx; total += ny;
    }
    return total;
}

// dc blocker -> agc(+gate) -> fm for nlanes lanes of n samples, fed in chunks; out is float (demod=1) or cf32.
long long emu_backend(int nlanes, long long lane_stride, int has_dc, float dc_alpha, int has_agc, float thr_db,
                      int demod, float kf, int L, int W, int G, const float2 *x, long long n, const long long *chunks,
                      int nchunks, void *out, unsigned long long *fixups_out, float *dbg /* optional [nseg][4] of the last chunk */)
{
    std::vector<LaneState> lane(nlanes);
    for (auto &l : lane) { l.dc_re = l.dc_im = 0; l.g = 1000.0f; l.y2p = 1.0f; l.mode = SQ_ENABLED; l.timer = 0; l.fm_re = l.fm_im = 0; }
    unsigned long long fixups2[3] = {0, 0, 0};
    std::vector<float2> dcst[2];
    dcst[0].assign(nlanes, make_float2(0, 0)); dcst[1] = dcst[0];
    int dc_cur = 0; unsigned epoch = 0;
    long long pos = 0;
    for (int c = 0; c < nchunks; c++) {
        int nx = (int)chunks[c];
        if (nx == 0) continue;
        int ngrp = (nx + G - 1) / G, nseg = (nx + L - 1) / L, nblk = (ngrp + kDcGB - 1) / kDcGB;
        const long long pws = (nx + 3) / 4 * 4;
        std::vector<SelfValid16> agg((size_t)nlanes * nblk);
        memset(agg.data(), 0xff, agg.size() * sizeof(SelfValid16));
        unsigned ticket[2] = {0, 0};
        std::vector<double> powA(kDcGB + 1), powAB;
        std::vector<SegState> ss((size_t)nlanes * nseg), se((size_t)nlanes * nseg);
        std::vector<FsmState> fs((size_t)nlanes * nseg), fe((size_t)nlanes * nseg);
        std::vector<float2> ydc((size_t)nlanes * pws), yfirst(nlanes);
        std::vector<float> pw((size_t)nlanes * pws), gfirst(nlanes);
        std::vector<float2> yend(nlanes), sylast((size_t)nlanes * nseg);
        unsigned barrier[4] = {0, 0, 0, 0};
        int nwords = (nx + 31) / 32;
        std::vector<unsigned> exb((size_t)nlanes * nwords), gb((size_t)nlanes * nwords), pg(nlanes, 0), ps(nlanes, 0);
        std::vector<unsigned> sgr((size_t)nlanes * nwords), sgi((size_t)nlanes * nwords), fb(2 * nlanes, 0xffffffffu), blist(2 * 4096);
        unsigned bcount[4] = {0, 0, 0, 0};
        DcParams d{};
        d.in = x + pos; d.in_lane_stride = lane_stride; d.n = nx; d.nlanes = nlanes; d.G = G; d.ngrp = ngrp; d.nblk = nblk;
        d.has_dc = has_dc; d.out = has_dc ? ydc.data() : nullptr; d.out_lane_stride = pws;
        d.pw = has_agc ? pw.data() : nullptr; d.pw_stride = pws;
        d.a1 = -1.0f + dc_alpha; d.c = -(double)d.a1; d.agg = agg.data(); d.ticket = ticket;
        d.dc_in = dcst[dc_cur].data(); d.dc_out = dcst[dc_cur ^ 1].data();
        { double A = 1.0; for (int i = 0; i < G; i++) A *= d.c; powA[0] = 1.0; for (int k = 1; k <= kDcGB; k++) powA[k] = powA[k - 1] * A; }
        {
            const double AB = powA[kDcGB];
            d.depth = std::max(1, (int)std::ceil(std::log(1e-13) / std::log(AB)));
            powAB.assign((size_t)d.depth + 1, 1.0);
            for (int k = 1; k <= d.depth; k++) powAB[k] = powAB[k - 1] * AB;
            d.powAB = powAB.data();
        }
        { double cs = 1.0; for (int i = 0; i < G / 32; i++) cs *= d.c; for (int k = 0; k < 5; k++) { d.cS[k] = cs; cs *= cs; } }
        d.powA = powA.data();
        EmuLaunch launch;
        if (has_dc) { be_launch_dc(launch, d, 3); dc_cur ^= 1; }
        else if (has_agc) be_launch_prep(launch, d);
        BackendParams b{};
        b.in = x + pos; b.in_lane_stride = lane_stride;
        b.out = demod ? (void *)((float *)out + pos) : (void *)((float2 *)out + pos); b.out_lane_stride = lane_stride;
        b.n = nx; b.nlanes = nlanes; b.L = L; b.W = W; b.G = G; b.nseg = nseg; b.ngrp = ngrp;
        b.has_dc = has_dc; b.has_agc = has_agc; b.demod = demod;
        b.alpha = 0.1f; b.one_minus_alpha_f = (float)(1.0 - (double)b.alpha); b.neg_half_alpha = -0.5f * b.alpha;
        b.g_thr = design::agc_gain_threshold(thr_db); b.timeout = 1000; b.fm_ref = (float)(1.0f / (2 * design::kPi * kf));
        b.squelch_enabled = 1; b.gate = 1;
        b.lane = lane.data(); b.seg_start = ss.data(); b.seg_end = se.data();
        b.ydc = has_dc ? ydc.data() : x + pos; b.ydc_stride = has_dc ? pws : lane_stride;
        b.pw = pw.data(); b.pw_stride = pws; b.g_first = gfirst.data(); b.y_first = yfirst.data();
        b.y_end = yend.data(); b.seg_ylast = sylast.data(); b.barrier = barrier;
        b.nwords = nwords; b.FW = (1000 + 8 + L - 1) / L;
        b.exbits = exb.data(); b.gatebits = gb.data(); b.fsm_start = fs.data(); b.fsm_end = fe.data();
        b.prev_gate = pg.data(); b.prev_sign = ps.data(); b.sgnr = sgr.data(); b.sgni = sgi.data();
        b.first_bad = fb.data(); b.fixups = fixups2; b.bad_list = blist.data(); b.bad_count = bcount; b.bad_cap = 4096;
        be_launch(launch, b);
        if (dbg) for (int i = 0; i < nseg; i++) { float *d4 = dbg + 4 * i; d4[0] = ss[i].g; d4[1] = ss[i].y2p; d4[2] = se[i].g; d4[3] = se[i].y2p; }
        pos += nx;
    }
    if (fixups_out) { fixups_out[0] = fixups2[0]; fixups_out[1] = fixups2[1]; fixups_out[2] = fixups2[2]; }
    return pos;
}

}  // extern "C"

// firpfbchChannelizer M (pre-rotation NCO + analyzer), fed in chunks of whole frames; y is [M][nf_total] channel-major.
// kind: 0 generic k_pfb, 1 k_pfb_tile (M = 2..32), 2 k_pfb_ring (M = 128..1024)
extern "C" long long emu_pfb(int M, int kind, const float2 *x, long long nf_total, const long long *chunk_frames, int nchunks,
                             float2 *y)
{
    const int m = 7, P = 2 * m;
    std::vector<float> h = design::design_firpfbch((unsigned)M, (unsigned)m, 80.0f);
    std::vector<float2> tw(M);
    for (int i = 0; i < M; i++) { tw[i].x = (float)std::cos(-2.0 * design::kPi * i / M); tw[i].y = (float)std::sin(-2.0 * design::kPi * i / M); }
    int log2M = -1;
    if (M > 1 && (M & (M - 1)) == 0) { log2M = 0; while ((1 << log2M) < M) log2M++; }
    const unsigned dth = design::nco_constrain(design::firpfbch_rotation((unsigned)M));
    unsigned theta = 0;
    const size_t H = (size_t)(P - 1) * M;
    std::vector<float2> xr(H, make_float2(0, 0));
    EmuLaunch launch;
    long long pos = 0;
    for (int c = 0; c < nchunks; c++) {
        const long long nf = chunk_frames[c];
        if (pos + nf > nf_total) return -1;
        if (nf == 0) continue;
        xr.resize(H + (size_t)nf * M);
        launch(k_nco_mix, dim3(4), dim3(64), 0, x + pos * M, xr.data() + H, nf * M, theta, dth, 1, 0);
        theta += (unsigned)(nf * M) * dth;
        if (kind == 1 || kind == 4 || kind == 5) {
            PfbTileParams tp{};
            tp.xr = xr.data(); tp.y = y + pos; tp.y_stride = nf_total; tp.nf = (int)nf; tp.ocs = 1;
            for (int i = 0; i < M / 2; i++) tp.tw[i] = tw[i];
            for (int k = 0; k < P; k++) for (int n = 0; n < M; n++) tp.h[k * M + n] = h[(M - 1 - n) + k * M];
            const size_t smem = (size_t)(kPfbTileF + kPfbTileP - 1) * (M + 2) * sizeof(float2);
            const dim3 g((unsigned)((nf + kPfbTileF - 1) / kPfbTileF)), b(kPfbTileF);
            if (kind == 5) {                // even M that is not a power of two
                if (M == 20) launch(k_pfb_tile_any<20>, g, b, smem, tp);
                else if (M == 12) launch(k_pfb_tile_any<12>, g, b, smem, tp);
                else if (M == 6) launch(k_pfb_tile_any<6>, g, b, smem, tp);
                else return -1;
            } else if (kind == 4) {                // two frames per thread (M = 8, 16)
                const dim3 b2(kPfbTileF / 2);
                if (log2M == 3) launch(k_pfb_tile2<3>, g, b2, smem, tp);
                else if (log2M == 4) launch(k_pfb_tile2<4>, g, b2, smem, tp);
                else return -1;
            } else
            switch (log2M) {
            case 1: launch(k_pfb_tile<1>, g, b, smem, tp); break;
            case 2: launch(k_pfb_tile<2>, g, b, smem, tp); break;
            case 3: launch(k_pfb_tile<3>, g, b, smem, tp); break;
            case 4: launch(k_pfb_tile<4>, g, b, smem, tp); break;
            case 5: launch(k_pfb_tile<5>, g, b, smem, tp); break;
            default: return -1;
            }
        } else if (kind == 3) {
            PfbStreamParams sp{};
            std::vector<unsigned short> pm(M);
            pfb_stream_perm(M, pm.data());
            sp.xr = xr.data(); sp.y = y + pos; sp.y_stride = nf_total; sp.nf = (int)nf; sp.M = M; sp.log2M = log2M;
            sp.h = h.data(); sp.tw = tw.data(); sp.perm = pm.data(); sp.ocs = 1;
            sp.T = 3 * kPfbStTF;
            const dim3 gs((unsigned)((nf + sp.T - 1) / sp.T)), bs(M);
            switch (log2M) {
            case 7: launch(k_pfb_stream<7>, gs, bs, pfb_stream_smem(M), sp); break;
            case 8: launch(k_pfb_stream<8>, gs, bs, pfb_stream_smem(M), sp); break;
            case 9: launch(k_pfb_stream<9>, gs, bs, pfb_stream_smem(M), sp); break;
            case 10: launch(k_pfb_stream<10>, gs, bs, pfb_stream_smem(M), sp); break;
            default: return -1;
            }
        } else if (kind == 2) {
            PfbRingParams rp{};
            rp.xr = xr.data(); rp.y = y + pos; rp.y_stride = nf_total; rp.nf = (int)nf; rp.M = M; rp.log2M = log2M;
            rp.h = h.data(); rp.tw = tw.data();
            rp.T = 3 * kPfbRingTF;
            if (pfb_ring_lfz(log2M) == 4) launch(k_pfb_ring<4>, dim3((unsigned)((nf + rp.T - 1) / rp.T)), dim3(2 * M / kPfbRingCPT), pfb_ring_smem(M, log2M), rp);
            else launch(k_pfb_ring<5>, dim3((unsigned)((nf + rp.T - 1) / rp.T)), dim3(2 * M / kPfbRingCPT), pfb_ring_smem(M, log2M), rp);
        } else {
            PfbParams p{};
            p.xr = xr.data(); p.y = y + pos; p.y_stride = nf_total; p.M = M; p.P = P; p.nf = (int)nf; p.F = 8; p.log2M = log2M;
            p.h = h.data(); p.tw = tw.data(); p.hop = M;
            const size_t smem = (size_t)p.F * M * sizeof(float2) * (log2M >= 0 ? 1 : 2);
            launch(k_pfb, dim3((unsigned)((nf + p.F - 1) / p.F)), dim3(64), smem, p);
        }
        // carry the last (P-1) rows
        std::vector<float2> tail(xr.end() - (long)H, xr.end());
        xr.assign(tail.begin(), tail.end());
        pos += nf;
    }
    return pos;
}

// wbFMDemodulator's tail (de-emphasis sections + firdecim), `lanes` independent sequences x[lane][n_total] fed in
// chunks; y is [lanes][n_total / M].  order/fc: the Butterworth prototype.  Returns outputs per lane.
extern "C" long long emu_wbfm_tail(int lanes, unsigned order, float fc, unsigned M, const float *x, long long n_total,
                                   const long long *chunks, int nchunks, float *y, float *b_out, float *a_out)
{
    std::vector<design::Sos> sos = design::butter_lowpass_sos(order, fc);
    for (size_t i = 0; i < sos.size(); i++) for (int j = 0; j < 3; j++) { if (b_out) b_out[3 * i + j] = sos[i].b[j]; if (a_out) a_out[3 * i + j] = sos[i].a[j]; }
    std::vector<float> h = design::design_firdecim(M, 10, 60.0f);
    const int Lh = (int)h.size();
    const long long hs = Lh - 1 + M, ycap = n_total / M;
    std::vector<float> hist((size_t)hs * lanes, 0.0f);
    std::vector<float2> state(sos.size() * lanes, make_float2(0, 0));
    std::vector<double> qtab(sos.size() * 128);
    for (size_t i = 0; i < sos.size(); i++) iir2_q_table(sos[i].a[1], sos[i].a[2], qtab.data() + i * 128);
    EmuLaunch launch;
    long long pos = 0, produced = 0; size_t fill = 0;
    for (int c = 0; c < nchunks; c++) {
        const int n = (int)chunks[c];
        if (pos + n > n_total) return -1;
        if (n == 0) continue;
        const int head = Lh - 1 + (int)fill;
        const long long zs = ((long long)head + n + 3) & ~3LL;
        std::vector<float> z((size_t)zs * lanes, 0.0f);
        launch(k_rows_copy, dim3((head + 63) / 64, lanes), dim3(64), 0, (const float *)hist.data(), hs, 0LL, z.data(), zs, 0LL, head);
        for (size_t i = 0; i < sos.size(); i++) {
            Iir2Params p{};
            p.x = i ? z.data() + head : x + pos; p.x_stride = i ? zs : n_total; p.y = z.data() + head; p.y_stride = zs;
            iir2_plan(p, sos[i].b, sos[i].a, n);
            std::vector<double2> sz((size_t)p.nseg * lanes), si((size_t)p.nseg * lanes);
            p.seg_z = sz.data(); p.seg_in = si.data(); p.Q = qtab.data() + i * 128; p.state = state.data() + i * lanes;
            iir2_launch(launch, p, lanes);
        }
        const size_t tot = fill + (size_t)n;
        const int nb = (int)(tot / M);
        FirDecimParams fp{};
        fp.z = z.data(); fp.z_stride = zs; fp.y = y + produced; fp.y_stride = ycap; fp.h = h.data(); fp.nout = nb; fp.M = (int)M; fp.Lh = Lh;
        firdecim_launch(launch, fp, lanes, 3);
        fill = tot - (size_t)nb * M;
        const int keep = Lh - 1 + (int)fill;
        launch(k_rows_copy, dim3((keep + 63) / 64, lanes), dim3(64), 0, (const float *)z.data(), zs, (long long)nb * M, hist.data(), hs, 0LL, keep);
        pos += n; produced += nb;
    }
    return produced;
}

// firpfbch2 analyzer (k_pfb with hop M/2): x = nf_total * M/2 samples fed in chunks of whole frames; y [M][nf_total]
extern "C" long long emu_pfb2(int M, const float2 *x, long long nf_total, const long long *chunk_frames, int nchunks, float2 *y,
                              float *taps)
{
    const int m = 7, P = 2 * m, M2 = M / 2;
    std::vector<float> h = design::design_firpfbch2((unsigned)M, (unsigned)m, 80.0f);
    if (taps) memcpy(taps, h.data(), h.size() * sizeof(float));
    std::vector<float2> tw(M);
    for (int i = 0; i < M; i++) { tw[i].x = (float)std::cos(-2.0 * design::kPi * i / M); tw[i].y = (float)std::sin(-2.0 * design::kPi * i / M); }
    int log2M = -1;
    if (M > 1 && (M & (M - 1)) == 0) { log2M = 0; while ((1 << log2M) < M) log2M++; }
    const size_t H = (size_t)(P - 1) * M + M2;
    std::vector<float2> xr(H, make_float2(0, 0));
    EmuLaunch launch;
    long long pos = 0;
    for (int c = 0; c < nchunks; c++) {
        const long long nf = chunk_frames[c];
        if (pos + nf > nf_total) return -1;
        if (nf == 0) continue;
        xr.resize(H + (size_t)nf * M2);
        memcpy(xr.data() + H, x + pos * M2, (size_t)nf * M2 * sizeof(float2));
        PfbParams p{};
        p.xr = xr.data(); p.y = y + pos; p.y_stride = nf_total; p.M = M; p.P = P; p.nf = (int)nf; p.F = 8; p.log2M = log2M;
        p.h = h.data(); p.tw = tw.data(); p.hop = M2; p.over2 = 1; p.parity0 = (int)(pos & 1); p.scale = 1.0f / (float)M;
        const size_t smem = (size_t)p.F * M * sizeof(float2) * (log2M >= 0 ? 1 : 2);
        launch(k_pfb, dim3((unsigned)((nf + p.F - 1) / p.F)), dim3(64), smem, p);
        std::vector<float2> tail(xr.end() - (long)H, xr.end());
        xr.assign(tail.begin(), tail.end());
        pos += nf;
    }
    return pos;
}

// ---- product-side filter design (design.hpp), exported so the CPU suite can compare it with the oracle ----
extern "C" {
int emu_design_msresamp(float rate, float As, unsigned *S, unsigned *m /*[12]*/, float *h1 /*[12][32]*/, float *rate_arb,
                        unsigned *step, unsigned *npfb, float *bank /* npfb*14 */)
{
    design::MsresampPlan p = design::plan_msresamp(rate, As);
    *S = p.S; *rate_arb = p.rate_arb; *step = p.step; *npfb = p.npfb;
    for (unsigned s = 0; s < p.S && s < 12; s++) {
        m[s] = p.st[s].m;
        for (size_t u = 0; u < p.st[s].h1.size() && u < 32; u++) h1[s * 32 + u] = p.st[s].h1[u];
    }
    if (bank) memcpy(bank, p.bank.data(), p.bank.size() * sizeof(float));
    return 0;
}
int emu_design_firpfbch(unsigned M, unsigned m, float As, float *h) { auto v = design::design_firpfbch(M, m, As); memcpy(h, v.data(), v.size() * 4); return (int)v.size(); }
unsigned emu_design_nco_constrain(float f) { return design::nco_constrain(f); }
float emu_design_rotation(unsigned C) { return design::firpfbch_rotation(C); }
float emu_design_agc_threshold(float thr) { return design::agc_gain_threshold(thr); }
int emu_frontend_geometry(float rate, float As, int Tc, int *n /*[13]*/, int *d /*[13]*/, int *hcap, int *smem)
{
    design::MsresampPlan ms = design::plan_msresamp(rate, As);
    FrontendGeometry g = plan_frontend(ms, Tc);
    if (!g.error.empty()) return -1;
    for (int i = 0; i <= g.base.S; i++) { n[i] = g.base.n[i]; d[i] = g.base.d[i]; }
    *hcap = g.hcap; *smem = (int)g.smem_bytes;
    return g.base.S;
}
}
