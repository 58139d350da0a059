// chain.inl -- the fused receive chain (csdr_chain_*), i.e. sdrProcess of apps/SoapySDR.hs:181-283:
//   offset (mixDown/mixUp)  ->  resampler (msresamp r 60)  ->  dcBlocker  ->  compact  ->
//   C == 1 : demod                     where demod = (fmDemodulator kf | amDemodulator | id) . agc
//   C  > 1 : firpfbchChannelizer C -> mux (replicate C demod) [-> mix]
// `compact` only re-chunks (Trans.hs:58-85): here whole frames of C samples are consumed as they arrive and the
// < C left-over samples wait in the handle.  Included at the end of csdr_b200.cu.

struct csdr_chain_s : Tagged {
    Ctx ctx;
    csdr_chain_cfg cfg;
    unsigned C = 1, nstreams = 1, nout = 1, hop = 1;        // hop: input samples per channelizer frame (C, or C/2 oversampled)
    bool over2 = false;
    size_t esz = 8;
    bool has_resamp = false, has_mix = false, has_agc = false;
    int mix_mode = 0; uint32_t mix_theta = 0, mix_dtheta = 0; int quantize = 1;
    Frontend fe;
    Backend be;          // C == 1: dc + agc + fm per stream;  C > 1: agc + fm per channel
    Backend dcb;         // C > 1: the single wide-band dc blocker
    Channelizer ch; uint32_t rot_theta = 0, rot_dtheta = 0;
    AmDemod am;
    WbfmTail wb; bool has_wb = false;       // DeWBFM: de-emphasis + decimator behind the discriminator
    DevBuf wbout;
    DevBuf xin, r, left, chan, dem, amout, outstage;
    size_t nleft = 0;
    std::vector<void *> out_ptrs;
    unsigned long long fixups_seen = 0, fixups_last = 0;
    // second stream: the back end of part i runs while the front end filters part i+1 (and host copies overlap)
    cudaStream_t copy_stream = nullptr; cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    cudaStream_t be_stream = nullptr, d2h_stream = nullptr;
    std::vector<cudaEvent_t> ev_pool;
    DevBuf xpipe[2];
    // per-channel output pointers for the back end (it writes the caller's buffers directly): device table, filled through a
    // small ring of pinned staging slots
    DevBuf out_tab; void **tab_host = nullptr; cudaEvent_t tab_ev[4] = {nullptr, nullptr, nullptr, nullptr}; unsigned tab_next = 0;
    void *const *output_table(size_t count)
    {
        if (!tab_host) {
            CK(cudaHostAlloc((void **)&tab_host, 4 * 4096 * sizeof(void *), cudaHostAllocDefault));
            for (auto &e : tab_ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        if (count > 4096) return nullptr;
        const unsigned slot = tab_next++ & 3;
        CK(cudaEventSynchronize(tab_ev[slot]));                      // the copy that last used this slot has been executed
        void **h = tab_host + (size_t)slot * 4096;
        for (size_t i = 0; i < count; i++) h[i] = out_ptrs[i];
        out_tab.ensure(4096 * sizeof(void *));
        CK(cudaMemcpyAsync(out_tab.p, h, count * sizeof(void *), cudaMemcpyHostToDevice, ctx.stream));
        CK(cudaEventRecord(tab_ev[slot], ctx.stream));
        return (void *const *)out_tab.p;
    }
    cudaEvent_t event(size_t i)
    {
        while (ev_pool.size() <= i) {
            cudaEvent_t e = nullptr;
            CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ev_pool.push_back(e);
        }
        return ev_pool[i];
    }

    csdr_chain_s(const csdr_chain_cfg &c) : Tagged(TAG_CHAIN), ctx(c.device), cfg(c) {}
    ~csdr_chain_s()
    {
        cudaSetDevice(ctx.device);
        if (ctx.stream) cudaStreamSynchronize(ctx.stream);
        if (copy_stream) { cudaStreamSynchronize(copy_stream); cudaStreamDestroy(copy_stream); }
        if (be_stream) { cudaStreamSynchronize(be_stream); cudaStreamDestroy(be_stream); }
        if (d2h_stream) { cudaStreamSynchronize(d2h_stream); cudaStreamDestroy(d2h_stream); }
        for (auto e : ev_pool) if (e) cudaEventDestroy(e);
        for (auto e : ev_copy) if (e) cudaEventDestroy(e);
        for (auto e : ev_done) if (e) cudaEventDestroy(e);
        for (auto e : tab_ev) if (e) cudaEventDestroy(e);
        if (tab_host) cudaFreeHost(tab_host);
    }
};

namespace {

void chain_init(csdr_chain_s *q)
{
    const csdr_chain_cfg &c = q->cfg;
    if (!(c.samplerate > 0)) throw CudaError{"chain: samplerate must be positive"};
    q->C = c.channels > 1 ? c.channels : 1;
    q->over2 = q->C > 1 && c.channelizer == CSDR_CHANNELIZER_FIRPFBCH2;
    if (c.channelizer != CSDR_CHANNELIZER_FIRPFBCH && c.channelizer != CSDR_CHANNELIZER_FIRPFBCH2) throw CudaError{"chain: unknown channelizer"};
    if (q->over2 && (q->C & 1)) throw CudaError{"chain: firpfbch2 needs an even channel count"};
    q->hop = q->over2 ? q->C / 2 : q->C;
    q->nstreams = c.nstreams > 1 ? c.nstreams : 1;
    if (q->C > 1 && q->nstreams > 1) throw CudaError{"chain: channelizer with nstreams > 1 is not implemented"};
    if (c.demod < 0 || c.demod > 3) throw CudaError{"chain: unknown demodulator"};
    q->has_wb = c.demod == CSDR_DEMOD_WBFM;
    if (q->has_wb && c.decim > 4096) throw CudaError{"chain: DeWBFM decimation must be in [1, 4096]"};
    if (c.demod == CSDR_DEMOD_NBFM && !(c.kf > 0.0f)) throw CudaError{"chain: DeNBFM needs kf > 0"};
    q->nout = (q->C > 1 && !c.mix) ? q->C : 1;
    q->esz = c.demod ? sizeof(float) : sizeof(float2);
    q->has_agc = c.agc_thresh_db != 0.0f;
    q->quantize = g_options[CSDR_OPT_VCO_DIRECT] ? 0 : 1;
    // f = 2*pi*offset/samplerate :: Float, f > 0 -> mixDown f, f < 0 -> mixUp (-f)   (SoapySDR.hs:200-205)
    float f = (float)(2.0f * (float)design::kPi * (float)c.offset_hz / (float)c.samplerate);
    if (f != 0.0f) {
        q->has_mix = true;
        q->mix_mode = f > 0.0f ? 1 : 2;
        q->mix_dtheta = design::nco_constrain(f > 0.0f ? f : -f);
    }
    if (c.bandwidth_hz != 0.0) {
        q->has_resamp = true;
        q->fe.init(q->ctx, (float)(c.bandwidth_hz / c.samplerate), 60.0f, (int)q->nstreams);
        q->fe.mix_mode = q->mix_mode; q->fe.theta0 = 0; q->fe.dtheta = q->mix_dtheta; q->fe.quantize = q->quantize;
    }
    const int am_lanes = (int)(q->C > 1 ? q->C : q->nstreams);
    if (q->C == 1) {
        q->be.has_dc = true; q->be.dc_alpha = 0.0005f;
        q->be.has_agc = q->has_agc; q->be.agc_thr = c.agc_thresh_db;
        q->be.demod = (c.demod == CSDR_DEMOD_NBFM || q->has_wb) ? 1 : 0; q->be.kf = q->has_wb ? 0.6f : c.kf > 0 ? c.kf : 0.3f;
        q->be.init(q->ctx, (int)q->nstreams);
    } else {
        q->dcb.has_dc = true; q->dcb.dc_alpha = 0.0005f;
        q->dcb.init(q->ctx, 1);
        q->ch.init(q->ctx, q->C, 7, 80.0f, q->over2);
        q->rot_dtheta = q->over2 ? 0u : design::nco_constrain(design::firpfbch_rotation(q->C));
        q->be.has_dc = false;
        q->be.has_agc = q->has_agc; q->be.agc_thr = c.agc_thresh_db;
        q->be.demod = (c.demod == CSDR_DEMOD_NBFM || q->has_wb) ? 1 : 0; q->be.kf = q->has_wb ? 0.6f : c.kf > 0 ? c.kf : 0.3f;
        q->be.init(q->ctx, (int)q->C);
        q->left.ensure(sizeof(float2) * 2 * q->C);          // < C rotated left-over samples (+ a partial chunk appended)
    }
    if (c.demod == CSDR_DEMOD_AM) { q->am.init(q->ctx.stream, am_lanes, 0.8f, g_options[CSDR_OPT_AMPMODEM_PLL] != 0); q->am.spec = g_options[CSDR_OPT_AM_PLL_SEQUENTIAL] == 0; }
    // wbFMDemodulator outBW decim (SoapySDR.hs:257): the quadrature rate is the resampler's output rate
    if (q->has_wb) q->wb.init(q->ctx, c.bandwidth_hz != 0.0 ? c.bandwidth_hz : c.samplerate, c.decim > 1 ? c.decim : 1, am_lanes);
    q->out_ptrs.resize((size_t)q->nstreams * q->nout);
    CK(cudaStreamCreateWithFlags(&q->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&q->be_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
        CK(cudaEventCreateWithFlags(&q->ev_copy[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&q->ev_done[i], cudaEventDisableTiming));
    }
}

size_t chain_max_out(const csdr_chain_s *q, size_t nx)
{
    size_t n = q->has_resamp ? (size_t)q->fe.max_out((long long)nx) : nx;
    if (q->C > 1) n = (n + q->nleft) / q->hop + 1;
    if (q->has_wb) n = q->wb.max_out(n) + 1;
    return n;
}

// Run the chain on device-resident input xd ([nstreams][nx] at x_stride).  Results go to q->out_ptrs[] (device
// pointers, one per stream/output, capacity out_cap samples).  Returns samples per output.
size_t chain_run_device(csdr_chain_s *q, const float2 *xd, size_t nx, size_t x_stride, size_t out_cap)
{
    const Ctx &c = q->ctx;
    const unsigned S = q->nstreams;
    // ---- optional (CSDR_OPT_OVERLAP): software pipeline over parts of the chunk, back end of part i on a second
    // stream while the front end filters part i+1.  Measured on B200 it does not pay (0.87 ms vs 0.79 ms per 2^27
    // samples): the persistent front-end grid leaves room for one back-end CTA per SM, and the back end needs the
    // whole machine to hide the latency of its recurrences.  Off by default.
    if (q->C == 1 && S == 1 && q->has_resamp && q->cfg.demod != CSDR_DEMOD_AM && !q->has_wb && nx >= ((size_t)1 << 23) &&
        g_options[CSDR_OPT_OVERLAP] != 0) {
        // few, large parts: the back end of a part is latency bound (~0.15 ms almost regardless of its size)
        size_t nparts = nx >= ((size_t)1 << 28) ? 4 : 2;
        size_t part = ((nx + nparts - 1) / nparts + 255) & ~(size_t)255;
        nparts = (nx + part - 1) / part;
        const size_t rcap = ((size_t)q->fe.max_out((long long)part) + 3) & ~(size_t)3;
        q->r.ensure(sizeof(float2) * rcap * nparts);
        size_t produced = 0;
        for (size_t i = 0; i < nparts; i++) {
            const size_t off = i * part, n_i = std::min(part, nx - off);
            float2 *ri = q->r.as<float2>() + i * rcap;
            const long long nr_i = q->fe.run(c, xd + off, (long long)n_i, 0, ri, 0);
            if (produced + (size_t)nr_i > out_cap) throw CudaError{"chain: output capacity too small"};
            CK(cudaEventRecord(q->event(i), c.stream));
            CK(cudaStreamWaitEvent(q->be_stream, q->event(i), 0));
            q->be.run_on(q->be_stream, ri, 0, (char *)q->out_ptrs[0] + produced * q->esz, 0, (int)nr_i);
            produced += (size_t)nr_i;
        }
        CK(cudaEventRecord(q->event(nparts), q->be_stream));
        CK(cudaStreamWaitEvent(c.stream, q->event(nparts), 0));
        return produced;
    }
    // ---- offset mix + resampler
    const float2 *r = xd; long long r_stride = (long long)x_stride; long long nr = (long long)nx;
    if (q->has_resamp) {
        r_stride = q->fe.max_out((long long)nx);
        q->r.ensure(sizeof(float2) * (size_t)r_stride * S);
        nr = q->fe.run(c, xd, (long long)nx, (long long)x_stride, q->r.as<float2>(), r_stride);
        r = q->r.as<float2>();
    } else if (q->has_mix && nx) {
        r_stride = (long long)nx;
        q->r.ensure(sizeof(float2) * nx * S);
        for (unsigned s = 0; s < S; s++)
            launch(k_nco_mix, dim3(grid_for((long long)nx, 256, c.sms)), dim3(256), 0, c.stream, xd + (size_t)s * x_stride,
                   q->r.as<float2>() + (size_t)s * nx, (long long)nx, q->mix_theta, q->mix_dtheta, q->quantize,
                   q->mix_mode == 2 ? 1 : 0);
        q->mix_theta += (uint32_t)nx * q->mix_dtheta;
        r = q->r.as<float2>();
    }
    if (nr > 0x7fffffffLL) throw CudaError{"chain: chunk too large"};

    if (q->C == 1) {
        if ((q->has_wb ? q->wb.max_out((size_t)nr) : (size_t)nr) > out_cap) throw CudaError{"chain: output capacity too small"};
        if (nr == 0) return 0;
        // out_ptrs are per-stream device buffers; the back end wants one base + stride
        const bool contiguous = (S == 1);
        if (q->cfg.demod == CSDR_DEMOD_AM) {
            // agc (cf32) -> ampmodem
            q->dem.ensure(sizeof(float2) * (size_t)nr * S);
            q->be.run(c, r, r_stride, q->dem.p, nr, (int)nr);
            float *dst = contiguous ? (float *)q->out_ptrs[0] : (q->amout.ensure(sizeof(float) * (size_t)nr * S), q->amout.as<float>());
            q->am.run(c.stream, q->dem.as<float2>(), nr, dst, nr, (int)nr);
            g_launches.fetch_add(q->am.take_launches());
            if (!contiguous)
                for (unsigned s = 0; s < S; s++)
                    CK(cudaMemcpyAsync(q->out_ptrs[s], dst + (size_t)s * nr, sizeof(float) * nr, cudaMemcpyDeviceToDevice, c.stream));
        } else if (q->has_wb) {
            // agc -> freqdem (kf 0.6) -> de-emphasis -> decimator
            q->dem.ensure(sizeof(float) * (size_t)nr * S);
            q->be.run(c, r, r_stride, q->dem.p, nr, (int)nr);
            const size_t nb = q->wb.max_out((size_t)nr);
            if (nb > out_cap) throw CudaError{"chain: output capacity too small"};
            float *dst = contiguous ? (float *)q->out_ptrs[0] : (q->wbout.ensure(sizeof(float) * (nb + 1) * S), q->wbout.as<float>());
            q->wb.run(c, q->dem.as<float>(), nr, (int)nr, dst, (long long)nb);
            if (!contiguous && nb)
                for (unsigned s = 0; s < S; s++)
                    CK(cudaMemcpyAsync(q->out_ptrs[s], dst + (size_t)s * nb, sizeof(float) * nb, cudaMemcpyDeviceToDevice, c.stream));
            return nb;
        } else {
            void *dst = q->out_ptrs[0];
            void *const *tab = contiguous ? nullptr : q->output_table(S);
            if (tab) { q->be.run(c, r, r_stride, nullptr, 0, (int)nr, tab); return (size_t)nr; }
            if (!contiguous) { q->dem.ensure(q->esz * (size_t)nr * S); dst = q->dem.p; }
            q->be.run(c, r, r_stride, dst, nr, (int)nr);
            if (!contiguous)
                for (unsigned s = 0; s < S; s++)
                    CK(cudaMemcpyAsync(q->out_ptrs[s], (char *)dst + (size_t)s * nr * q->esz, q->esz * nr, cudaMemcpyDeviceToDevice, c.stream));
        }
        return (size_t)nr;
    }

    // ---- channelizer path: wide-band dc blocker; its output pass also applies the channelizer's pre-rotation
    // (Liquid.chs:847) and writes straight into the channelizer's input slot, behind the (already rotated) left-over
    const unsigned C = q->C, hop = q->hop;
    const size_t tot = q->nleft + (size_t)nr, nf = tot / hop, used_in = nf * hop, used = nf * C;
    if ((q->has_wb ? q->wb.max_out(nf) : nf) > out_cap) throw CudaError{"chain: output capacity too small"};
    float2 *dst = nullptr;
    if (nf) {
        float2 *slot = q->ch.input_slot(c, tot);
        if (q->nleft) CK(cudaMemcpyAsync(slot, q->left.p, sizeof(float2) * q->nleft, cudaMemcpyDeviceToDevice, c.stream));
        dst = slot + q->nleft;
    } else if (nr) {
        // not even one frame yet (tot < C): append to the left-over
        dst = q->left.as<float2>() + q->nleft;
    }
    if (nr) {
        q->dcb.run_dc_only(c, r, 0, dst, 0, (int)nr, !q->over2, q->rot_theta, q->rot_dtheta, q->quantize);
        q->rot_theta += (uint32_t)nr * q->rot_dtheta;
    }
    if (nf) {
        float2 *slot = dst - q->nleft;
        q->chan.ensure(sizeof(float2) * used);
        long long pw_stride = 0;
        float *pw = q->has_agc ? q->be.pw_target((int)nf, &pw_stride) : nullptr;     // the channelizer writes |y|^2 as well
        q->ch.run(c, (int)nf, q->chan.as<float2>(), (long long)nf, pw, pw_stride);
        // remaining < C (rotated) samples wait for the next call
        const size_t rem = tot - used_in;
        if (rem) CK(cudaMemcpyAsync(q->left.p, slot + used_in, sizeof(float2) * rem, cudaMemcpyDeviceToDevice, c.stream));
        q->nleft = rem;
        // per-channel agc -> demod
        void *dem = nullptr;
        if (q->cfg.demod == CSDR_DEMOD_AM) {
            q->dem.ensure(sizeof(float2) * used);
            q->be.run(c, q->chan.as<float2>(), (long long)nf, q->dem.p, (long long)nf, (int)nf);
            q->amout.ensure(sizeof(float) * used);
            q->am.run(c.stream, q->dem.as<float2>(), (long long)nf, q->amout.as<float>(), (long long)nf, (int)nf);
            g_launches.fetch_add(q->am.take_launches());
            dem = q->amout.p;
        } else if (!q->cfg.mix && !q->has_wb) {
            // every channel's result goes straight into the caller's buffer for that channel
            void *const *tab = q->output_table(C);
            if (tab) { q->be.run(c, q->chan.as<float2>(), (long long)nf, nullptr, 0, (int)nf, tab); return nf; }
            q->dem.ensure(q->esz * used);
            q->be.run(c, q->chan.as<float2>(), (long long)nf, q->dem.p, (long long)nf, (int)nf);
            dem = q->dem.p;
        } else {
            q->dem.ensure(q->esz * used);
            q->be.run(c, q->chan.as<float2>(), (long long)nf, q->dem.p, (long long)nf, (int)nf);
            dem = q->dem.p;
        }
        size_t nfo = nf;                               // samples per channel behind the demodulator
        if (q->has_wb) {
            nfo = q->wb.max_out(nf);
            q->wbout.ensure(sizeof(float) * (nfo + 1) * C);
            q->wb.run(c, (const float *)dem, (long long)nf, (int)nf, q->wbout.as<float>(), (long long)nfo);
            dem = q->wbout.p;
        }
        if (nfo == 0) return 0;
        if (q->cfg.mix) {
            // mix = foldl1 (zipWith (+)) over channels 1..C (Trans.hs:119-122); cf32 is summed as 2 floats
            const long long nfl = (long long)nfo * (long long)(q->esz / sizeof(float));
            // summands straight from the gated back end: closed words are exact zeros and are not read
            const bool gated = dem == q->dem.p && q->be.has_agc && q->be.gate && q->cfg.demod != CSDR_DEMOD_AM && !q->has_wb && q->be.last_nwords > 0;
            if (gated)
                launch(k_lane_sum_gated, dim3(grid_for((long long)q->be.last_nwords * 32, 256, c.sms)), dim3(256), 0, c.stream, (const float *)dem, nfl,
                       (int)C, (float *)q->out_ptrs[0], (long long)nfo, (const unsigned *)q->be.gatebits.as<unsigned>(), q->be.last_nwords,
                       (const unsigned *)q->be.prev_gate.as<unsigned>(), (int)(q->esz / sizeof(float)), q->be.demod == 1 ? 1 : 0);
            else
                launch(k_lane_sum, dim3(grid_for(nfl, 256, c.sms)), dim3(256), 0, c.stream, (const float *)dem, nfl, (int)C,
                       (float *)q->out_ptrs[0], nfl);
        } else {
            for (unsigned ch = 0; ch < C; ch++)
                CK(cudaMemcpyAsync(q->out_ptrs[ch], (char *)dem + (size_t)ch * nfo * q->esz, q->esz * nfo, cudaMemcpyDeviceToDevice, c.stream));
        }
        return nfo;
    } else if (nr) {
        q->nleft = tot;
    }
    return 0;
}

}  // namespace

extern "C" {

csdr_chain csdr_chain_create(const csdr_chain_cfg *cfg)
{
    API_BEGIN
    if (!cfg) throw CudaError{"chain: null configuration"};
    std::unique_ptr<csdr_chain_s> q(new csdr_chain_s(*cfg));
    chain_init(q.get());
    q->ctx.sync();
    return q.release();
    API_END(nullptr)
}
int csdr_chain_destroy(csdr_chain q) { if (!q) return 0; REQUIRE(q, TAG_CHAIN, -1); delete q; return 0; }
unsigned csdr_chain_num_outputs(csdr_chain q) { REQUIRE(q, TAG_CHAIN, 0); return q->nout; }
size_t csdr_chain_out_elem_size(csdr_chain q) { REQUIRE(q, TAG_CHAIN, 0); return q->esz; }
size_t csdr_chain_max_output(csdr_chain q, size_t nx) { REQUIRE(q, TAG_CHAIN, 0); return chain_max_out(q, nx); }
void * csdr_chain_cuda_stream(csdr_chain q) { REQUIRE(q, TAG_CHAIN, nullptr); return (void *)q->ctx.stream; }
void csdr_chain_print(csdr_chain q)
{
    REQUIRE(q, TAG_CHAIN, );
    const csdr_chain_cfg &c = q->cfg;
    printf("csdr_b200 chain: sr %.1f offset %.1f bw %.1f demod %d kf %.3f agc %.1f dB channels %u mix %d streams %u\n",
           c.samplerate, c.offset_hz, c.bandwidth_hz, c.demod, c.kf, c.agc_thresh_db, q->C, c.mix, q->nstreams);
    if (q->has_mix) printf("  offset nco [phase: 0x%.8x rad, freq: 0x%.8x rad/sample] %s\n", q->mix_theta, q->mix_dtheta, q->mix_mode == 1 ? "down" : "up");
    if (q->has_resamp) {
        const auto &ms = q->fe.ms;
        printf("  msresamp rate %.10f: %u half-band stages (m =", ms.rate, ms.S);
        for (unsigned s = 0; s < ms.S; s++) printf(" %u", ms.st[s].m);
        printf("), arbitrary %.10f step 0x%08x, tile %d c-samples, smem %zu B\n", ms.rate_arb, ms.step, q->fe.geo.base.Tc, q->fe.geo.smem_bytes);
    }
    if (q->C > 1) printf("  firpfbch %u channels, rotation freq word 0x%.8x\n", q->C, q->rot_dtheta);
}

int csdr_chain_process(csdr_chain q, const csdr_cf32 *x, size_t nx, size_t x_stride, void *const *outs, size_t out_cap,
                       size_t *n_out)
{
    if (n_out) *n_out = 0;
    REQUIRE(q, TAG_CHAIN, -1);
    API_BEGIN
    q->ctx.use();
    const Ctx &c = q->ctx;
    const unsigned S = q->nstreams;
    const size_t nptr = (size_t)S * q->nout;
    if (S == 1) x_stride = nx;
    // classify outputs; host outputs are produced into device staging and copied back at the end
    bool any_host_out = false;
    for (size_t i = 0; i < nptr; i++) if (!is_device_ptr(outs[i])) any_host_out = true;
    const size_t cap_each = std::min(out_cap, chain_max_out(q, nx));
    if (any_host_out) {
        q->outstage.ensure(cap_each * q->esz * nptr);
        for (size_t i = 0; i < nptr; i++) q->out_ptrs[i] = (char *)q->outstage.p + i * cap_each * q->esz;
    } else {
        for (size_t i = 0; i < nptr; i++) q->out_ptrs[i] = outs[i];
    }
    size_t n = 0;
    constexpr size_t kPart = (size_t)1 << 23;                      // 64 MiB of CF32 per pipelined part
    if (nx == 0 || is_device_ptr(x)) {
        n = chain_run_device(q, (const float2 *)x, nx, x_stride, any_host_out ? cap_each : out_cap);
    } else if (S == 1 && nx >= 2 * kPart) {
        // large host chunk: the stream is fed in parts, the host->device copy of part i+1 (copy stream) and the
        // device->host copy of part i-1's results (third stream) overlap the kernels of part i (PCIe is full duplex)
        if (!q->d2h_stream) CK(cudaStreamCreateWithFlags(&q->d2h_stream, cudaStreamNonBlocking));
        q->xin.ensure(sizeof(float2) * nx);
        const size_t nparts = (nx + kPart - 1) / kPart, cap = any_host_out ? cap_each : out_cap;
        constexpr size_t kEv = 128;                                   // event-pool indices of this path
        CK(cudaEventRecord(q->event(kEv - 1), c.stream));              // earlier work on the handle may still read xin
        CK(cudaStreamWaitEvent(q->copy_stream, q->event(kEv - 1), 0));
        for (size_t i = 0; i < nparts; i++) {
            const size_t off = i * kPart, cnt = std::min(kPart, nx - off);
            CK(cudaMemcpyAsync(q->xin.as<float2>() + off, (const float2 *)x + off, cnt * sizeof(float2), cudaMemcpyHostToDevice, q->copy_stream));
            CK(cudaEventRecord(q->event(kEv + 2 * i), q->copy_stream));
        }
        const std::vector<void *> base = q->out_ptrs;
        for (size_t i = 0; i < nparts; i++) {
            const size_t off = i * kPart, cnt = std::min(kPart, nx - off);
            CK(cudaStreamWaitEvent(c.stream, q->event(kEv + 2 * i), 0));
            for (size_t k = 0; k < nptr; k++) q->out_ptrs[k] = (char *)base[k] + n * q->esz;
            const size_t ni = chain_run_device(q, q->xin.as<float2>() + off, cnt, cnt, cap - n);
            if (any_host_out && ni) {
                CK(cudaEventRecord(q->event(kEv + 2 * i + 1), c.stream));
                CK(cudaStreamWaitEvent(q->d2h_stream, q->event(kEv + 2 * i + 1), 0));
                for (size_t k = 0; k < nptr; k++)
                    CK(cudaMemcpyAsync((char *)outs[k] + n * q->esz, q->out_ptrs[k], ni * q->esz,
                                       is_device_ptr(outs[k]) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, q->d2h_stream));
            }
            n += ni;
        }
        q->out_ptrs = base;
        if (any_host_out) { CK(cudaStreamSynchronize(q->d2h_stream)); c.sync(); }
        else CK(cudaEventSynchronize(q->event(kEv + 2 * (nparts - 1))));   // the caller may reuse x: its last part has been copied
        if (n_out) *n_out = n;
        return 0;
    } else {
        // host input: one asynchronous copy per stream into contiguous device staging
        q->xin.ensure(sizeof(float2) * nx * S);
        CK(cudaMemcpy2DAsync(q->xin.p, nx * sizeof(float2), x, x_stride * sizeof(float2), nx * sizeof(float2), S,
                             cudaMemcpyHostToDevice, c.stream));
        CK(cudaEventRecord(q->event(127), c.stream));
        n = chain_run_device(q, q->xin.as<float2>(), nx, nx, any_host_out ? cap_each : out_cap);
        if (!any_host_out) CK(cudaEventSynchronize(q->event(127)));       // pinned x: the copy is asynchronous, the caller may reuse x
    }
    if (any_host_out) {
        for (size_t i = 0; i < nptr; i++)
            if (n) CK(cudaMemcpyAsync(outs[i], q->out_ptrs[i], n * q->esz, is_device_ptr(outs[i]) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c.stream));
        c.sync();
    }
    if (n_out) *n_out = n;
    return 0;
    API_END(-1)
}

// ---- CF32 file in, raw sample files out: readFromFile (Source.chs:259-271: raw interleaved
// little-endian float32 I/Q pairs, read in arrays of `chunk` samples), takeNArr numsamples behind the resampler
// (SoapySDR.hs:207), fileSink = FS.writeChunks (Sink.hs:29-34) with the app's names: <name>.cf32, or <name>_ch<K>.cf32
// (K from 1) behind the channelizer without --mix (SoapySDR.hs:222-240).  Demodulated outputs are written as raw
// float32 (.f32); the reference wraps them in AU/WAV containers through libsndfile, which is outside this path.
// The next chunk is read and the previous results are written by helper threads while the chain runs.
int csdr_chain_run_file(csdr_chain q, const char *in_path, const char *out_name, uint64_t numsamples, size_t chunk,
                        uint64_t *n_in, uint64_t *n_out)
{
    REQUIRE(q, TAG_CHAIN, -1);
    API_BEGIN
    if (!in_path || !out_name) throw CudaError{"chain_run_file: null path"};
    if (q->nstreams != 1) throw CudaError{"chain_run_file: one stream per file"};
    q->ctx.use();
    if (chunk == 0) chunk = (size_t)1 << 24;
    struct File { FILE *f = nullptr; ~File() { if (f) fclose(f); } };
    struct Pinned { void *p = nullptr; ~Pinned() { if (p) cudaFreeHost(p); } };
    File fin; fin.f = fopen(in_path, "rb");
    if (!fin.f) throw CudaError{std::string("chain_run_file: cannot open ") + in_path};
    const unsigned nout = q->nout;
    const char *ext = q->esz == sizeof(float2) ? ".cf32" : ".f32";
    std::vector<File> fout(nout);
    for (unsigned k = 0; k < nout; k++) {
        const std::string name = nout == 1 ? std::string(out_name) + ext : std::string(out_name) + "_ch" + std::to_string(k + 1) + ext;
        fout[k].f = fopen(name.c_str(), "wb");
        if (!fout[k].f) throw CudaError{"chain_run_file: cannot create " + name};
    }
    // takeNArr counts samples behind the resampler: numsamples / C frames per channel, / decim behind the decimator
    uint64_t limit = ~0ULL;
    if (numsamples) {
        limit = numsamples / q->C;
        if (q->has_wb) limit /= q->wb.dec.M;
    }
    const size_t cap = chain_max_out(q, chunk) + 8;
    Pinned xin[2], yout[2];
    for (int i = 0; i < 2; i++) {
        CK(cudaHostAlloc(&xin[i].p, chunk * sizeof(float2), cudaHostAllocDefault));
        CK(cudaHostAlloc(&yout[i].p, (size_t)nout * cap * q->esz, cudaHostAllocDefault));
    }
    auto read_chunk = [&](int i) { return fread(xin[i].p, sizeof(float2), chunk, fin.f); };
    uint64_t produced = 0, consumed = 0;
    std::future<size_t> rd = std::async(std::launch::async, read_chunk, 0);
    std::future<bool> wr;
    for (int i = 0; produced < limit; i ^= 1) {
        const size_t nx = rd.get();
        if (nx == 0) break;
        rd = std::async(std::launch::async, read_chunk, i ^ 1);
        std::vector<void *> outs(nout);
        for (unsigned k = 0; k < nout; k++) outs[k] = (char *)yout[i].p + (size_t)k * cap * q->esz;
        size_t n = 0;
        if (csdr_chain_process(q, (const csdr_cf32 *)xin[i].p, nx, nx, outs.data(), cap, &n) != 0) {
            if (wr.valid()) wr.get();
            rd.get();
            return -1;                                      // csdr_last_error() holds the reason
        }
        consumed += nx;
        const size_t take = (size_t)std::min<uint64_t>(n, limit - produced);
        if (wr.valid() && !wr.get()) { rd.get(); throw CudaError{"chain_run_file: write failed"}; }
        wr = std::async(std::launch::async, [&fout, outs, take, nout, esz = q->esz]() {
            for (unsigned k = 0; k < nout; k++)
                if (take && fwrite(outs[k], esz, take, fout[k].f) != take) return false;
            return true;
        });
        produced += take;
    }
    if (rd.valid()) rd.get();
    if (wr.valid() && !wr.get()) throw CudaError{"chain_run_file: write failed"};
    if (n_in) *n_in = consumed;
    if (n_out) *n_out = produced;
    return 0;
    API_END(-1)
}

int csdr_chain_seek(csdr_chain q, uint64_t n_prior)
{
    REQUIRE(q, TAG_CHAIN, -1);
    API_BEGIN
    q->ctx.use();
    // samples behind the resampler that precede the new position (closed form: fe_seek)
    unsigned long long o_prior = n_prior;
    if (q->has_resamp) {
        if (q->fe.interp) {
            if (q->C > 1) throw CudaError{"chain: seek with a channelizer behind an interpolating resampler is not implemented"};
            q->fe.cursor = FrontendCursor(); q->fe.cursor.n_abs = n_prior;
        } else {
            q->fe.cursor = fe_seek(q->fe.geo, n_prior);
            const unsigned __int128 span = (unsigned __int128)(n_prior >> q->fe.geo.base.S) << 24, st = q->fe.geo.base.step;
            o_prior = (unsigned long long)((span + st - 1) / st);
        }
    } else q->mix_theta = (uint32_t)n_prior * q->mix_dtheta;
    if (q->C > 1) {
        // channelizer: the pre-rotation NCO (Liquid.chs:817-821, 847) has advanced by o_prior samples; frames stay on the
        // stream's own grid (sample index = 0 mod C), so the o_prior mod C samples of the frame the new position falls into
        // count as already waiting (zeros: that frame and the 13 that still see the empty history belong to the warm-up)
        // (firpfbch2: frames of C/2 samples, no pre-rotation; the sign (-1)^(c t) of its per-channel factor follows the absolute
        // frame index, so the filterbank's frame counter restarts at the number of whole frames in front of the position)
        q->rot_theta = (uint32_t)o_prior * q->rot_dtheta;
        q->nleft = (size_t)(o_prior % q->hop);
        q->ch.frames_done = o_prior / q->hop;
        CK(cudaMemsetAsync(q->left.p, 0, q->left.cap, q->ctx.stream));
        for (auto &b : q->ch.xr) if (b.p) CK(cudaMemsetAsync(b.p, 0, q->ch.hist_samples() * sizeof(float2), q->ctx.stream));
        q->ctx.sync();
    }
    if (q->has_wb) {
        // wide-band FM tail: the output decimator consumes whole blocks of `decim` demodulated samples on the stream's own grid,
        // so the samples of the block the new position falls into count as pending (zeros; they belong to the warm-up, like
        // the de-emphasis filter's and the decimator's empty histories: 2 decim m + 1 taps, m = 10)
        const unsigned long long idx = q->C > 1 ? o_prior / q->hop : o_prior;       // demodulated samples per lane in front
        q->wb.dec.fill = (size_t)(idx % q->wb.dec.M);
        CK(cudaMemsetAsync(q->wb.dec.hist.p, 0, q->wb.dec.hist.cap, q->ctx.stream));
        CK(cudaMemsetAsync(q->wb.deemph.state.p, 0, q->wb.deemph.state.cap, q->ctx.stream));
        q->ctx.sync();
    }
    return 0;
    API_END(-1)
}
size_t csdr_chain_warmup_len(csdr_chain q)
{
    REQUIRE(q, TAG_CHAIN, 0);
    // front-end FIR history + dc blocker settling (0.9995^k < 1e-9 after ~41.5k post-resample samples) + AGC; behind a
    // channelizer the per-channel loops run at 1/C of that rate: 13 frames of filterbank history, the gain loop's
    // settling (~400 samples) and the squelch FSM's memory (timeout + 8 samples) per channel
    double r = q->has_resamp ? (double)q->fe.ms.rate : 1.0;
    double post = 45000.0;
    if (q->C > 1) post += (double)q->hop * (28.0 + 512.0 + (double)q->be.agc_timeout + 8.0);
    if (q->has_wb) post += (double)q->hop * (2.0 * 10.0 * (double)q->wb.dec.M + 64.0);   // decimator taps + de-emphasis settling, per lane
    size_t w = (size_t)(q->has_resamp ? q->fe.geo.hcap : 0) + (size_t)std::ceil(post / r);
    return w;
}
int csdr_chain_profile(csdr_chain q, int enable)
{
    REQUIRE(q, TAG_CHAIN, -1);
    API_BEGIN
    q->fe.collect(q->ctx);
    q->fe.profile = enable != 0;
    q->fe.prof_ms = 0.0; q->fe.prof_launches = 0;
    return 0;
    API_END(-1)
}
double csdr_chain_frontend_ms(csdr_chain q, uint64_t *launches)
{
    REQUIRE(q, TAG_CHAIN, -1.0);
    API_BEGIN
    q->fe.collect(q->ctx);
    if (launches) *launches = q->fe.prof_launches;
    return q->fe.prof_ms;
    API_END(-1.0)
}
uint64_t csdr_chain_agc_fixups(csdr_chain q)
{
    REQUIRE(q, TAG_CHAIN, 0);
    API_BEGIN
    unsigned long long v = q->be.read_fixups(q->ctx);
    unsigned long long d = v - q->fixups_seen;
    q->fixups_seen = v;
    return d;
    API_END(0)
}
int csdr_chain_agc_counters(csdr_chain q, uint64_t out[3])
{
    REQUIRE(q, TAG_CHAIN, -1);
    API_BEGIN
    unsigned long long v[3] = {0, 0, 0};
    CK(cudaMemcpyAsync(v, q->be.fixups.p, sizeof(v), cudaMemcpyDeviceToHost, q->ctx.stream));
    q->ctx.sync();
    out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
    return 0;
    API_END(-1)
}
int csdr_chain_agc_plan(csdr_chain q, int out[2])
{
    REQUIRE(q, TAG_CHAIN, -1);
    API_BEGIN
    out[0] = q->be.last_L; out[1] = q->be.last_W;
    return 0;
    API_END(-1)
}

}  // extern "C"
