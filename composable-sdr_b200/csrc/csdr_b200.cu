// csdr_b200.cu -- host side of libcsdr_b200.so: handles, device-resident stream state, kernel launches and the
// C ABI declared in include/csdr_b200.h.  CUDA runtime only (no torch, no other library on the data path).
#include "../../include/csdr_b200.h"

#include "platform.cuh"
#include "design.hpp"
#include "frontend.cuh"
#include "frontend_std.cuh"
#include "frontend_plan.hpp"
#include "interp.cuh"
#include <cuda.h>
#include "backend.cuh"
#include "pfb.cuh"
#include <stdexcept>
#include <string>
#include "ampmodem.cuh"
#include "wbfm.cuh"

#include <atomic>
#include <map>
#include <future>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

using namespace csdr;

// ------------------------------------------------------------------------------------------ library state
namespace {

thread_local std::string t_err;
std::atomic<unsigned long long> g_launches{0};
int g_options[16] = {0, 1, 0, 512, 384, 0, 0, 0, 0, 1};

void set_err(const std::string &s) { t_err = s; }
void clear_err() { t_err.clear(); }

struct CudaError { std::string msg; };
#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            throw CudaError{std::string(#call) + ": " + cudaGetErrorString(e_)};                      \
    } while (0)

template <class... Args, class... Act>
void launch(void (*k)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act &&...a)
{
    k<<<grid, block, smem, st>>>(std::forward<Act>(a)...);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CK(cudaGetLastError());
}

// launch as thread-block clusters of `cluster_x` CTAs (distributed shared memory between the CTAs of a cluster)
template <class... Args, class... Act>
void launch_cluster(void (*k)(Args...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, unsigned cluster_x, Act &&...a)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k, Args(std::forward<Act>(a))...));
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

// cooperative launch (all CTAs co-resident: the kernel synchronises the grid itself): as many CTAs as fit, at most `want`
template <class... Args, class... Act>
void launch_coop(void (*k)(Args...), int want, dim3 block, cudaStream_t st, int sms, Act &&...a)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, int> per_sm;
    int dev = 0, occ = 0;
    CK(cudaGetDevice(&dev));
    {
        std::lock_guard<std::mutex> lk(mu);
        int &v = per_sm[{reinterpret_cast<const void *>(k), dev}];
        if (v == 0) { CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, k, (int)block.x, 0)); if (v < 1) throw CudaError{"cooperative kernel does not fit on an SM"}; }
        occ = v;
    }
    const int grid = std::max(1, std::min(want, occ * sms));
    std::tuple<std::decay_t<Args>...> vals(std::forward<Act>(a)...);
    void *argv[sizeof...(Args)];
    size_t i = 0;
    std::apply([&](auto &...v) { ((argv[i++] = (void *)&v), ...); }, vals);
    CK(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(k), dim3(grid), block, argv, 0, st));
    g_launches.fetch_add(1, std::memory_order_relaxed);
}

struct DevBuf {
    void *p = nullptr; size_t cap = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    void ensure(size_t bytes)
    {
        if (bytes <= cap) return;
        if (p) { CK(cudaFree(p)); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void ensure_zero(size_t bytes, cudaStream_t st)
    {
        if (bytes <= cap) return;
        ensure(bytes);
        CK(cudaMemsetAsync(p, 0, cap, st));
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

bool is_device_ptr(const void *p)
{
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Every handle owns a stream on one device.
struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sms = 148;
    Ctx(int dev)
    {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) throw CudaError{"no CUDA device available (libcsdr_b200 has no CPU fallback)"};
        if (dev < 0) CK(cudaGetDevice(&dev));
        device = dev;
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    }
    ~Ctx() { if (stream) { cudaSetDevice(device); cudaStreamSynchronize(stream); cudaStreamDestroy(stream); } }
    void use() const { CK(cudaSetDevice(device)); }
    void sync() const { CK(cudaStreamSynchronize(stream)); }
};

// Input/output staging for the liquid-compatible block calls: host pointers are copied through device scratch.
struct Staging {
    DevBuf in, out;
    const void *to_dev(const Ctx &c, const void *p, size_t bytes)
    {
        if (bytes == 0 || is_device_ptr(p)) return p;
        in.ensure(bytes);
        CK(cudaMemcpyAsync(in.p, p, bytes, cudaMemcpyHostToDevice, c.stream));
        return in.p;
    }
    void *out_dev(void *p, size_t bytes)
    {
        if (bytes == 0 || is_device_ptr(p)) return p;
        out.ensure(bytes);
        return out.p;
    }
    void finish(const Ctx &c, void *user, void *dev, size_t bytes)
    {
        if (bytes && user != dev) CK(cudaMemcpyAsync(user, dev, bytes, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
    }
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-function (and per-device) limit shared by every handle that uses
// the kernel: it is only ever RAISED, so that a later handle with a smaller tile cannot starve an earlier one.
template <class K>
void raise_dyn_smem(K kernel, size_t bytes)
{
    static std::mutex mu;
    static std::map<std::pair<const void *, int>, size_t> seen;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = seen[{reinterpret_cast<const void *>(kernel), dev}];
    if (bytes <= cur) return;
    CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
}

int grid_for(long long n, int block, int sms, int per_sm = 8)
{
    long long g = (n + block - 1) / block;
    long long cap = (long long)sms * per_sm;
    return (int)std::max<long long>(1, std::min(g, cap));
}

// launch-sequence helpers shared with the CPU-only emulation take a callable `launch(kernel, grid, block, smem, args...)`
struct StreamLauncher {
    cudaStream_t st;
    template <class... A, class... B>
    void operator()(void (*k)(A...), dim3 g, dim3 b, size_t sm, B &&...a) const { launch(k, g, b, sm, st, std::forward<B>(a)...); }
};

// ---------------------------------------------------------------------------------- front end (mix + msresamp)
struct Frontend {
    design::MsresampPlan ms;
    FrontendGeometry geo;
    int nstreams = 1;
    DevBuf hist[2]; int cur = 0;
    DevBuf bank;
    FrontendCursor cursor;
    bool interp = false; InterpPlan ip; DevBuf ibuf[2], xmix;      // rate > 1: arbitrary stage first, then half-band interpolators
    int mix_mode = 0; uint32_t theta0 = 0, dtheta = 0; int quantize = 1;
    void (*kernel)(FrontendParams) = k_frontend;
    void (*kernel_direct)(FrontendParams, FeTmap) = nullptr;     // k_frontend_direct<S>: needs a tensor map per call
    int ctas_per_sm = 2, fe_threads = 256;
    // optional event timing of k_frontend (bench roofline): pairs recorded on the launching stream
    bool profile = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
    double prof_ms = 0.0; unsigned long long prof_launches = 0;
    ~Frontend()
    {
        for (auto &e : ev_pending) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
        for (auto &e : ev_free) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    }
    void collect(const Ctx &c)
    {
        if (ev_pending.empty()) return;
        c.sync();
        for (auto &e : ev_pending) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e.first, e.second));
            prof_ms += ms; prof_launches++;
            ev_free.push_back(e);
        }
        ev_pending.clear();
    }

    void init(const Ctx &c, float rate, float As, int streams)
    {
        ms = design::plan_msresamp(rate, As, g_options[CSDR_OPT_RESAMP_FC_OLD] != 0);
        nstreams = streams;
        interp = ms.interp;
        if (interp) {
            if (ms.S > (unsigned)kMaxStages) throw CudaError{"msresamp: too many half-band stages"};
            if (2 * ms.m_arb != (unsigned)kHsub) throw CudaError{"msresamp: unexpected arbitrary-stage length"};
            ip = InterpPlan{};
            ip.S = (int)ms.S; ip.step = ms.step; ip.bits = (int)ms.bits;
            for (int st = 0; st < ip.S; st++) {
                ip.m[st] = (int)ms.st[st].m;
                if (ip.m[st] > kMaxHbM) throw CudaError{"msresamp: half-band stage too long"};
                for (int u = 0; u < 2 * ip.m[st]; u++) ip.h1[st][u] = ms.st[st].h1[u];
            }
            geo = FrontendGeometry{};
            geo.hcap = kInterpHcap; geo.base.S = 0; geo.base.step = ms.step;
            size_t hb = (size_t)geo.hcap * sizeof(float2) * nstreams;
            for (auto &h : hist) { h.ensure(hb); CK(cudaMemsetAsync(h.p, 0, h.cap, c.stream)); }
            bank.ensure(ms.bank.size() * sizeof(float));
            CK(cudaMemcpyAsync(bank.p, ms.bank.data(), ms.bank.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
            c.sync();
            return;
        }
        int Tc = 464;
        if (ms.S != 3) {
            // aim for ~4096 input samples per tile, c-count multiple of 8
            int t = 4096 >> ms.S; t = std::max(32, std::min(2048, t));
            Tc = (t + 7) / 8 * 8;
        }
        geo = plan_frontend(ms, Tc, g_options[CSDR_OPT_GENERIC_FRONTEND] == 0, g_options[CSDR_OPT_FRONTEND_VARIANT]);
        if (!geo.error.empty()) throw CudaError{geo.error};
        kernel = k_frontend;
        if (geo.std_kernel && geo.variant == 0) {
            switch (ms.S) {
            case 1: kernel = k_frontend_std<1>; break;
            case 2: kernel = k_frontend_std<2>; break;
            case 3: kernel = k_frontend_std<3>; break;
            case 4: kernel = k_frontend_std<4>; break;
            case 5: kernel = k_frontend_std<5>; break;
            case 6: kernel = k_frontend_std<6>; break;
            default: throw CudaError{"frontend: no specialised kernel for this stage count"};
            }
        } else if (geo.std_kernel && geo.variant == 2) {
            switch (ms.S) {
            case 2: kernel_direct = k_frontend_ws<2>; break;
            case 3: kernel_direct = k_frontend_ws<3>; break;
            case 4: kernel_direct = k_frontend_ws<4>; break;
            case 5: kernel_direct = k_frontend_ws<5>; break;
            case 6: kernel_direct = k_frontend_ws<6>; break;
            default: throw CudaError{"frontend: no warp-specialised kernel for this stage count"};
            }
        } else if (geo.std_kernel) {
            switch (ms.S) {
            case 1: kernel_direct = k_frontend_direct<1>; break;
            case 2: kernel_direct = k_frontend_direct<2>; break;
            case 3: kernel_direct = k_frontend_direct<3>; break;
            case 4: kernel_direct = k_frontend_direct<4>; break;
            case 5: kernel_direct = k_frontend_direct<5>; break;
            case 6: kernel_direct = k_frontend_direct<6>; break;
            default: throw CudaError{"frontend: no direct-read kernel for this stage count"};
            }
        }
        size_t hb = (size_t)geo.hcap * sizeof(float2) * nstreams;
        for (auto &h : hist) { h.ensure(hb); CK(cudaMemsetAsync(h.p, 0, h.cap, c.stream)); }
        bank.ensure(ms.bank.size() * sizeof(float));
        CK(cudaMemcpyAsync(bank.p, ms.bank.data(), ms.bank.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
        fe_threads = !geo.std_kernel ? 256 : geo.variant == 2 ? 2 * kFeWsGroup : kFeNT;
        if (kernel_direct) {
            raise_dyn_smem(kernel_direct, geo.smem_bytes);
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel_direct, fe_threads, geo.smem_bytes));
        } else {
            raise_dyn_smem(kernel, geo.smem_bytes);
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, fe_threads, geo.smem_bytes));
        }
        if (ctas_per_sm < 1) throw CudaError{"frontend: tile does not fit in shared memory"};
        c.sync();
    }
    // Tensor map over the chunk for k_frontend_direct: x viewed as [nstreams][rows][32 floats], rows of 16 samples,
    // starting at sample tma_r so that every tile starts on a row boundary (tiles are a multiple of 16 samples apart).
    // Not possible (tma_ok = 0, tiles are then filled by ordinary loads) when that start is not 16-byte aligned.
    void make_tensor_map(const float2 *x, long long nx, long long x_stride, FrontendParams &p, FeTmap &tm)
    {
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                     const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        static EncodeFn encode = nullptr;
        static bool looked = false;
        if (!looked) {
            looked = true;
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess)
                encode = (EncodeFn)fn;
        }
        p.tma_ok = 0; p.tma_r = 0; p.tma_rows = 0;
        if (!encode || nx < kFeRawRow) return;
        const FeGeom &G = geo.geom;
        const long long lo0 = (p.K0 - kHcPad) * (1LL << G.S) + G.d[G.S];          // first tile's first raw sample
        const long long r = (((lo0 - p.n0) % kFeRawRow) + kFeRawRow) % kFeRawRow;
        const float2 *base = x + r;
        const long long rows = (nx - r) / kFeRawRow;
        if (rows < 1 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return;
        if (nstreams > 1 && ((x_stride * (long long)sizeof(float2)) & 15) != 0) return;
        constexpr int kBoxMax = 256;
        const int tile_rows = G.n[G.S] / kFeRawRow, nb = (tile_rows + kBoxMax - 1) / kBoxMax, box = tile_rows / nb;
        cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)nstreams};
        cuuint64_t strides[2] = {128, (cuuint64_t)(nstreams > 1 ? x_stride * (long long)sizeof(float2) : rows * 128)};
        cuuint32_t boxd[3] = {32, (cuuint32_t)box, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        static_assert(sizeof(FeTmap) == sizeof(CUtensorMap), "FeTmap mirrors CUtensorMap");
        const CUresult rc = encode(reinterpret_cast<CUtensorMap *>(&tm), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)base, dims, strides, boxd,
                                   estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return;
        p.tma_ok = 1; p.tma_r = (int)r; p.tma_rows = rows;
    }
    long long run_interp(const Ctx &c, const float2 *x, long long nx, long long x_stride, float2 *y, long long y_stride)
    {
        if (nx <= 0) return 0;
        if (mix_mode) {
            // mix into a scratch copy first (stream s at s * nx); the carried history then holds mixed samples
            xmix.ensure(sizeof(float2) * (size_t)nstreams * nx);
            for (int s = 0; s < nstreams; s++)
                launch(k_nco_mix, dim3(grid_for(nx, 256, c.sms)), dim3(256), 0, c.stream, x + (long long)s * x_stride,
                       xmix.as<float2>() + (long long)s * nx, nx, theta0 + (uint32_t)cursor.n_abs * dtheta, dtheta, quantize, mix_mode == 2 ? 1 : 0);
            x = xmix.as<float2>(); x_stride = nx;
        }
        StreamLauncher l{c.stream};
        auto buf = [&](int slot, size_t bytes) -> void * { ibuf[slot].ensure(bytes); return ibuf[slot].p; };
        const long long ny = interp_launch(l, buf, ip, bank.as<float>(), nstreams, x, x_stride, hist[cur].as<float2>(), cursor.n_abs, nx, y, y_stride);
        launch(k_hist_update, dim3((geo.hcap + 255) / 256, nstreams), dim3(256), 0, c.stream,
               (const float2 *)hist[cur].as<float2>(), hist[cur ^ 1].as<float2>(), x, x_stride, nx, geo.hcap);
        cur ^= 1;
        cursor.n_abs += (unsigned long long)nx;
        return ny;
    }
    long long max_out(long long nx) const
    {
        if (interp) return interp_max_out(ip, nx);
        // pushes <= nx/2^S + 1, each push emits at most ceil(2^24/step) outputs
        double pushes = (double)(nx >> ms.S) + 1.0;
        return (long long)std::ceil(pushes * 16777216.0 / (double)ms.step) + 4;
    }
    // x: device, nstreams x nx at x_stride; y: device, y_stride.  Returns outputs per stream.
    long long run(const Ctx &c, const float2 *x, long long nx, long long x_stride, float2 *y, long long y_stride)
    {
        if (interp) return run_interp(c, x, nx, x_stride, y, y_stride);
        FrontendParams p = geo.base;
        long long ny = fe_prepare_call(geo, cursor, nx, p);
        p.x = x; p.hist = hist[cur].as<float2>(); p.y = y; p.hcap = geo.hcap;
        p.x_stride = x_stride; p.y_stride = y_stride;
        p.mix_mode = mix_mode; p.theta0 = theta0; p.dtheta = dtheta; p.quantize = quantize;
        p.bank = bank.as<float>();
        if (p.ntiles > 0) {
            const int slots = c.sms * ctas_per_sm;      // persistent CTAs: one wave, tiles strided over the grid
            int gx = std::min(p.ntiles, std::max(1, slots / std::max(1, std::min(nstreams, slots))));
            std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
            if (profile) {
                if (ev_pending.size() > 4096) collect(c);
                if (!ev_free.empty()) { ev = ev_free.back(); ev_free.pop_back(); }
                else { CK(cudaEventCreate(&ev.first)); CK(cudaEventCreate(&ev.second)); }
                CK(cudaEventRecord(ev.first, c.stream));
            }
            if (kernel_direct) {
                FeTmap tm{};
                make_tensor_map(x, nx, x_stride, p, tm);
                launch(kernel_direct, dim3(gx, nstreams), dim3(fe_threads), geo.smem_bytes, c.stream, p, tm);
            } else {
                launch(kernel, dim3(gx, nstreams), dim3(fe_threads), geo.smem_bytes, c.stream, p);
            }
            if (profile) { CK(cudaEventRecord(ev.second, c.stream)); ev_pending.push_back(ev); }
        }
        if (nx > 0) {
            launch(k_hist_update, dim3((geo.hcap + 255) / 256, nstreams), dim3(256), 0, c.stream,
                   (const float2 *)hist[cur].as<float2>(), hist[cur ^ 1].as<float2>(), x, x_stride, nx, geo.hcap);
            cur ^= 1;
        }
        return ny;
    }
};

// ---------------------------------------------------------------------------------- back end (dc, agc, fm)
struct Backend {
    int nlanes = 1;
    bool has_dc = false, has_agc = false; int demod = 0;
    float dc_alpha = 0.0005f;
    float agc_bw = 0.1f, agc_thr = 0.f; unsigned agc_timeout = 1000; bool squelch = true, gate = true;
    float kf = 0.3f;
    int L = 512, W = 384, G = 128; bool fixed_L = false;
    DevBuf lane, dc_agg, dc_flag, dc_ticket, dc_state[2], powA, powAB, ss, se, fs, fe, exbits, gatebits, sgnr, sgni, prev_gate, prev_sign, first_bad, bad_list, bad_count, fixups;
    DevBuf ydc, pwbuf, g_first, y_first, y_end, seg_ylast, barrier;
    int pw_ready_n = -1;       // the producer of the input has already written its power for a call of this many samples
    int FW = 3; unsigned long long last_refined = 0; int last_L = 0, last_W = 0, last_nwords = 0;
    // self-tuning warm-up, deterministic: the counters of every call are copied to one of kLag pinned slots; call k waits for
    // the copy of call k - kLag (up to kLag - 1 calls stay in flight, so the wait is normally over before it starts) and
    // lengthens / shortens the warm-up from it
    static constexpr int kLag = 4;        // the warm-up plan of call k follows the counters of call k - kLag
    unsigned long long *h_counters = nullptr; cudaEvent_t ev_counters[kLag] = {}; unsigned long long calls = 0;
    unsigned long long seen[3] = {0, 0, 0}; int W_cur = 0, calm_calls = 0; long long nseg_of[kLag] = {1, 1, 1, 1};
    int sms = 148, dc_cur = 0, dc_depth = 1; unsigned dc_epoch = 0;
    ~Backend() { if (h_counters) cudaFreeHost(h_counters); for (auto e : ev_counters) if (e) cudaEventDestroy(e); }

    void init(const Ctx &c, int lanes, float g0 = 1000.0f, int mode0 = SQ_ENABLED)
    {
        nlanes = lanes; sms = c.sms;
        // segment and warm-up lengths in whole 32-sample words (one warp handles a dc group, one ballot a word)
        L = (std::max(64, g_options[CSDR_OPT_AGC_SEGMENT]) + 31) / 32 * 32; W = (std::max(32, g_options[CSDR_OPT_AGC_WARMUP]) + 31) / 32 * 32;
        fixed_L = g_options[CSDR_OPT_AGC_SEGMENT] != 512;      // an explicit setting is taken literally
        G = 128;
        while (L % G || W % G) G /= 2;
        std::vector<LaneState> ls(nlanes);
        for (auto &l : ls) { l.dc_re = l.dc_im = 0; l.g = g0; l.y2p = 1.0f; l.mode = mode0; l.timer = 0; l.fm_re = l.fm_im = 0; }
        lane.ensure(sizeof(LaneState) * nlanes);
        CK(cudaMemcpyAsync(lane.p, ls.data(), sizeof(LaneState) * nlanes, cudaMemcpyHostToDevice, c.stream));
        prev_gate.ensure(sizeof(unsigned) * nlanes); CK(cudaMemsetAsync(prev_gate.p, 0, prev_gate.cap, c.stream));
        prev_sign.ensure(sizeof(unsigned) * nlanes); CK(cudaMemsetAsync(prev_sign.p, 0, prev_sign.cap, c.stream));
        first_bad.ensure(sizeof(unsigned) * 2 * nlanes); CK(cudaMemsetAsync(first_bad.p, 0xff, first_bad.cap, c.stream));
        // A^k, A = c^G, for the dc blocker's group-boundary states; AB^k, AB = A^kDcGB, for the look-back over blocks
        {
            const double cc = -(double)(-1.0f + dc_alpha);
            double A = 1.0;
            for (int i = 0; i < G; i++) A *= cc;
            std::vector<double> pw(kDcGB + 1);
            pw[0] = 1.0;
            for (int k = 1; k <= kDcGB; k++) pw[k] = pw[k - 1] * A;
            powA.ensure(sizeof(double) * pw.size());
            CK(cudaMemcpyAsync(powA.p, pw.data(), sizeof(double) * pw.size(), cudaMemcpyHostToDevice, c.stream));
            const double AB = pw[kDcGB];
            dc_depth = 1;
            if (AB > 0.0 && AB < 1.0) dc_depth = (int)std::min(65536.0, std::ceil(std::log(1e-13) / std::log(AB)));
            else if (AB >= 1.0) dc_depth = 65536;
            dc_depth = std::max(1, dc_depth);
            std::vector<double> pb((size_t)dc_depth + 1);
            pb[0] = 1.0;
            for (int k = 1; k <= dc_depth; k++) pb[k] = pb[k - 1] * AB;
            powAB.ensure(sizeof(double) * pb.size());
            CK(cudaMemcpyAsync(powAB.p, pb.data(), sizeof(double) * pb.size(), cudaMemcpyHostToDevice, c.stream));
            dc_ticket.ensure(2 * sizeof(unsigned)); CK(cudaMemsetAsync(dc_ticket.p, 0, dc_ticket.cap, c.stream));
            raise_dyn_smem(k_dc_scan<4>, kDcSmem); raise_dyn_smem(k_dc_scan<2>, kDcSmem); raise_dyn_smem(k_dc_scan<1>, kDcSmem);
            for (auto &d : dc_state) { d.ensure(sizeof(float2) * nlanes); CK(cudaMemsetAsync(d.p, 0, d.cap, c.stream)); }
            c.sync();
        }
        if (has_agc) {
            int per_sm = 0;
            raise_dyn_smem(k_agc_emit<false, false>, kAgcSmem); raise_dyn_smem(k_agc_emit<false, true>, kAgcSmem);
            raise_dyn_smem(k_agc_emit<true, false>, kAgcSmem); raise_dyn_smem(k_agc_emit<true, true>, kAgcSmem);
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_agc_emit<false, true>, kAgcT, kAgcSmem));
            agc_slots = std::max(1, per_sm) * c.sms;
        }
        fixups.ensure(3 * sizeof(unsigned long long)); CK(cudaMemsetAsync(fixups.p, 0, fixups.cap, c.stream));
        bad_list.ensure(sizeof(unsigned) * 2 * 65536); bad_count.ensure(4 * sizeof(unsigned)); CK(cudaMemsetAsync(bad_count.p, 0, bad_count.cap, c.stream));
        barrier.ensure(4 * sizeof(unsigned)); CK(cudaMemsetAsync(barrier.p, 0, barrier.cap, c.stream));
        y_end.ensure(sizeof(float2) * nlanes); CK(cudaMemsetAsync(y_end.p, 0, y_end.cap, c.stream));

        c.sync();
    }
    struct Launcher {
        cudaStream_t st; int sms;
        template <class... Args, class... Act>
        void coop(void (*k)(Args...), dim3 block, Act &&...a) const { launch_coop(k, 2 * sms, block, st, sms, std::forward<Act>(a)...); }
        // CSDR_OPT_DEBUG: print the first speculation misses (start state vs predecessor's end state)
        void debug_after_verify(const BackendParams &b) const
        {
            if (!g_options[CSDR_OPT_DEBUG]) return;
            CK(cudaStreamSynchronize(st));
            unsigned cnt = 0;
            CK(cudaMemcpy(&cnt, b.bad_count, sizeof(cnt), cudaMemcpyDeviceToHost));
            unsigned show = std::min(cnt, 12u);
            std::vector<unsigned> lst(show);
            if (show) CK(cudaMemcpy(lst.data(), b.bad_list, show * sizeof(unsigned), cudaMemcpyDeviceToHost));
            fprintf(stderr, "[csdr debug] n=%d L=%d W=%d nseg=%d: %u segments do not continue their predecessor\n", b.n, b.L, b.W, b.nseg, cnt);
            for (unsigned i = 0; i < show; i++) {
                SegState a, e;
                CK(cudaMemcpy(&a, b.seg_start + lst[i], sizeof(a), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(&e, b.seg_end + lst[i] - 1, sizeof(e), cudaMemcpyDeviceToHost));     // (the list is filled by k_be_finish: previous call's)
                fprintf(stderr, "  seg %u: start g=%.9g y2p=%.9g | pred end g=%.9g y2p=%.9g | rel %.3g %.3g\n", lst[i], a.g, a.y2p, e.g,
                        e.y2p, fabs(a.g - e.g) / fmax(a.g, e.g), fabs(a.y2p - e.y2p) / fmax(a.y2p, e.y2p));
            }
        }
        template <class... Args, class... Act>
        void operator()(void (*k)(Args...), dim3 grid, dim3 block, size_t smem, Act &&...a) const
        {
            launch(k, grid, block, smem, st, std::forward<Act>(a)...);
        }
    };
    int agc_slots = 0;           // CTAs of k_agc_emit that are resident at once (whole device)
    int pick_segment(int n, int W) const
    {
        const long long slots = std::max(1, agc_slots);
        long long best_cost = -1; int best = 256;
        for (int L = 64; L <= 1024; L += 32) {
            const long long nseg = ((long long)n + L - 1) / L;
            const long long ctas = (long long)nlanes * ((nseg + kAgcT - 1) / kAgcT);
            const long long cost = ((ctas + slots - 1) / slots) * (long long)(W + L);
            if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = L; }
        }
        return best;
    }
    DcParams dc_params(const float2 *in, long long in_stride, float2 *out, long long out_stride, int n, cudaStream_t st)
    {
        DcParams d{};
        d.in = in; d.in_lane_stride = in_stride; d.out = out; d.out_lane_stride = out_stride;
        d.n = n; d.nlanes = nlanes; d.G = G; d.ngrp = (n + G - 1) / G; d.nblk = (d.ngrp + kDcGB - 1) / kDcGB;
        d.a1 = -1.0f + dc_alpha; d.c = -(double)d.a1; d.has_dc = 1;
        { double cs = 1.0; for (int i = 0; i < G / 32; i++) cs *= d.c; for (int k = 0; k < 5; k++) { d.cS[k] = cs; cs *= cs; } }
        d.powA = powA.as<double>(); d.powAB = powAB.as<double>(); d.depth = dc_depth;
        return d;
    }
    // one pass of k_dc_scan over d (look-back buffers, ticket, the two copies of the carried filter state)
    template <class L> void launch_dc(L &l, DcParams &d, cudaStream_t st)
    {
        const size_t blocks = (size_t)nlanes * d.nblk;
        dc_agg.ensure(sizeof(SelfValid16) * blocks);
        CK(cudaMemsetAsync(dc_agg.p, 0xff, sizeof(SelfValid16) * blocks, st));      // "not published yet"
        d.agg = dc_agg.as<SelfValid16>(); d.ticket = dc_ticket.as<unsigned>();
        d.dc_in = dc_state[dc_cur].as<float2>(); d.dc_out = dc_state[dc_cur ^ 1].as<float2>();
        dc_cur ^= 1;
        be_launch_dc(l, d, kDcMinB * sms);
    }
    // dc blocker only, out may alias in
    // rot: also multiply by the conjugate NCO phasor (theta0 + i * dtheta), the channelizer's pre-rotation
    void run_dc_only(const Ctx &c, const float2 *in, long long in_stride, float2 *out, long long out_stride, int n,
                     bool rot = false, uint32_t rot_theta = 0, uint32_t rot_dtheta = 0, int rot_quantize = 1)
    {
        if (n <= 0) return;
        DcParams d = dc_params(in, in_stride, out, out_stride, n, c.stream);
        d.rot = rot ? 1 : 0; d.rot_theta = rot_theta; d.rot_dtheta = rot_dtheta; d.rot_quantize = rot_quantize;
        Launcher l{c.stream, sms};
        launch_dc(l, d, c.stream);
    }
    // Where the producer of the next run()'s input (n samples per lane, no dc blocker here) may write |x|^2 itself:
    // returns the power array and its lane stride; run() then skips its own power pass.
    float *pw_target(int n, long long *stride)
    {
        if (!has_agc || has_dc || n <= 0) return nullptr;
        const long long pws = ((long long)n + 3) / 4 * 4;
        pwbuf.ensure(sizeof(float) * (size_t)nlanes * pws);
        pw_ready_n = n;
        *stride = pws;
        return pwbuf.as<float>();
    }
    // [dc] -> [agc+gate] -> [fm]; out: float (demod) or float2
    void run(const Ctx &c, const float2 *in, long long in_stride, void *out, long long out_stride, int n, void *const *out_table = nullptr)
    {
        run_on(c.stream, in, in_stride, out, out_stride, n, out_table);
    }
    void run_on(cudaStream_t st, const float2 *in, long long in_stride, void *out, long long out_stride, int n, void *const *out_table = nullptr)
    {
        if (n <= 0) return;
        Launcher l{st, sms};
        if (has_dc && !has_agc && demod == 0 && !out_table) {
            // dc blocker only (config 1): the output pass writes the caller's buffer directly
            DcParams d = dc_params(in, in_stride, (float2 *)out, out_stride, n, st);
            launch_dc(l, d, st);
            return;
        }
        // segment length: every chain of a wave is resident at once and a wave lasts one window (W + L samples), so the
        // cost of a call is  waves x (W + L);  the multiple of 32 in [64, 1024] that minimises it is taken (ties: longer
        // segments = less redundant warm-up work)
        int L = this->L;
        if (has_agc && !fixed_L) L = pick_segment(n, this->W_cur > 0 ? this->W_cur : this->W);
        int W = this->W;
        const int slot = (int)(calls % kLag);
        if (has_agc && !fixed_L) {
            if (!h_counters) {
                CK(cudaHostAlloc((void **)&h_counters, kLag * 3 * sizeof(unsigned long long), cudaHostAllocDefault));
                // (spinning wait: a blocking one costs an interrupt wake-up per call, ~1 ms on a box with eight busy ranks)
                for (auto &e : ev_counters) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                W_cur = this->W;
            }
            if (calls >= (unsigned long long)kLag) {
                CK(cudaEventSynchronize(ev_counters[slot]));           // counters as they stood after call k - kLag
                const unsigned long long *hc = h_counters + 3 * slot;
                const unsigned long long seq = hc[0] - seen[0], refined = hc[2] - seen[2];
                seen[0] = hc[0]; seen[1] = hc[1]; seen[2] = hc[2];
                if (seq > 0 || refined * 200 > (unsigned long long)nseg_of[slot]) { W_cur = std::min(W_cur * 2, 6144); calm_calls = 0; }
                else if (refined == 0 && ++calm_calls >= 8 && W_cur > this->W) { W_cur /= 2; calm_calls = 0; }
            }
            W = W_cur;
        }
        last_L = L; last_W = W;
        int ngrp = (n + G - 1) / G, nseg = (n + L - 1) / L, nwords = (n + 31) / 32;
        last_nwords = has_agc ? nwords : 0;
        size_t segs = (size_t)nlanes * nseg;
        ss.ensure(sizeof(SegState) * segs); se.ensure(sizeof(SegState) * segs);
        fs.ensure(sizeof(FsmState) * segs); fe.ensure(sizeof(FsmState) * segs);
        exbits.ensure(sizeof(unsigned) * (size_t)nlanes * nwords); gatebits.ensure(sizeof(unsigned) * (size_t)nlanes * nwords);
        sgnr.ensure(sizeof(unsigned) * (size_t)nlanes * nwords); sgni.ensure(sizeof(unsigned) * (size_t)nlanes * nwords);
        g_first.ensure(sizeof(float) * nlanes); y_first.ensure(sizeof(float2) * nlanes);
        const long long pws = ((long long)n + 3) / 4 * 4;          // lane stride of the per-sample work arrays
        BackendParams b{};
        b.ydc = in; b.ydc_stride = in_stride;
        {
            // dc blocker (states, then samples) and the power sequence the gain loop runs on
            DcParams d = dc_params(in, in_stride, nullptr, 0, n, st);
            d.has_dc = has_dc ? 1 : 0;
            if (has_dc) {
                ydc.ensure(sizeof(float2) * (size_t)nlanes * pws);
                d.out = ydc.as<float2>(); d.out_lane_stride = pws;
                b.ydc = ydc.as<float2>(); b.ydc_stride = pws;
            } else if ((const void *)in == (const void *)out && demod == 0) {
                // in-place cf32 call: k_be_emit reads the sample before the one it writes
                ydc.ensure(sizeof(float2) * (size_t)nlanes * pws);
                CK(cudaMemcpy2DAsync(ydc.p, sizeof(float2) * pws, in, sizeof(float2) * (size_t)(nlanes > 1 ? in_stride : n), sizeof(float2) * (size_t)n,
                                     nlanes, cudaMemcpyDeviceToDevice, st));
                b.ydc = ydc.as<float2>(); b.ydc_stride = pws;
            }
            if (has_agc) {
                pwbuf.ensure(sizeof(float) * (size_t)nlanes * pws);
                d.pw = pwbuf.as<float>(); d.pw_stride = pws;
            }
            if (has_dc) launch_dc(l, d, st);
            else if (has_agc && pw_ready_n != n) be_launch_prep(l, d);
            pw_ready_n = -1;
        }
        b.in = in; b.in_lane_stride = in_stride; b.out = out; b.out_lane_stride = out_stride; b.out_table = out_table;
        b.n = n; b.nlanes = nlanes; b.L = L; b.W = W; b.G = G; b.nseg = nseg; b.ngrp = ngrp;
        b.has_dc = has_dc; b.has_agc = has_agc; b.demod = demod;
        b.alpha = agc_bw; b.one_minus_alpha_f = (float)(1.0 - (double)agc_bw); b.neg_half_alpha = -0.5f * agc_bw;
        b.g_thr = design::agc_gain_threshold(agc_thr); b.timeout = agc_timeout;
        b.fm_ref = (float)(1.0f / (2 * design::kPi * kf));
        b.squelch_enabled = squelch ? 1 : 0; b.gate = gate ? 1 : 0;
        b.exact_math = g_options[CSDR_OPT_AGC_EXACT_MATH] ? 1 : 0;
        b.lane = lane.as<LaneState>(); b.seg_start = ss.as<SegState>(); b.seg_end = se.as<SegState>();
        b.pw = pwbuf.as<float>(); b.pw_stride = pws;
        b.g_first = g_first.as<float>(); b.y_first = y_first.as<float2>();
        seg_ylast.ensure(sizeof(float2) * segs);
        b.y_end = y_end.as<float2>(); b.seg_ylast = seg_ylast.as<float2>(); b.barrier = barrier.as<unsigned>();
        b.nwords = nwords;
        // the FSM forgets its entry state after timeout + 4 samples: replay that many bits (in whole segments)
        b.FW = (int)std::min<unsigned>(64u, (agc_timeout + 8 + (unsigned)L - 1) / (unsigned)L);
        b.exbits = exbits.as<unsigned>(); b.gatebits = gatebits.as<unsigned>();
        b.fsm_start = fs.as<FsmState>(); b.fsm_end = fe.as<FsmState>();
        b.prev_gate = prev_gate.as<unsigned>(); b.prev_sign = prev_sign.as<unsigned>();
        b.sgnr = sgnr.as<unsigned>(); b.sgni = sgni.as<unsigned>();
        b.first_bad = first_bad.as<unsigned>();
        b.bad_list = bad_list.as<unsigned>(); b.bad_count = bad_count.as<unsigned>(); b.bad_cap = 65536;
        b.fixups = fixups.as<unsigned long long>();
        be_launch(l, b);
        if (has_agc && !fixed_L) {
            CK(cudaMemcpyAsync(h_counters + 3 * slot, fixups.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ev_counters[slot], st));
            nseg_of[slot] = (long long)nlanes * nseg;
        }
        calls++;
    }
    unsigned long long read_fixups(const Ctx &c)
    {
        // sequential (in-order) repairs only: these are the ones that cost time; parallel refinements are v[2]
        unsigned long long v[3] = {0, 0, 0};
        CK(cudaMemcpyAsync(v, fixups.p, sizeof(v), cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        last_refined = v[2];
        return v[0] + v[1];
    }
    LaneState read_lane(const Ctx &c, int i)
    {
        LaneState l;
        CK(cudaMemcpyAsync(&l, lane.as<LaneState>() + i, sizeof(l), cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        return l;
    }
    void write_lane(const Ctx &c, int i, const LaneState &l)
    {
        CK(cudaMemcpyAsync(lane.as<LaneState>() + i, &l, sizeof(l), cudaMemcpyHostToDevice, c.stream));
        c.sync();
    }
};

// ---------------------------------------------------------------------------------- channelizer
struct Channelizer {
    unsigned M = 0, m = 0, P = 0; float As = 0;
    std::vector<float> h;
    DevBuf hd, tw, xr[2]; int cur = 0;       // xr: [(P-1)*M history | new samples], ping-pong for the history
    int log2M = -1, F = 1;
    size_t smem = 0;
    void (*tile_kernel)(PfbTileParams) = nullptr; PfbTileParams tp{}; size_t tile_smem = 0; bool tile_two = false;   // M = 2..32, m = 7
    bool ring_ok = false; int ring_ctas = 1; void (*ring_kernel)(PfbRingParams) = nullptr;                                                 // M = 128..1024, m = 7
    bool stream_ok = false; int stream_ctas = 1; DevBuf perm; void (*stream_kernel)(PfbStreamParams) = nullptr;
    void (*stream_pair)(PfbStreamParams) = nullptr;     // firpfbch2: even / odd passes as clusters of two CTAs (CSDR_OPT_PFB_VARIANT = 3: two launches)                                                                              // M = 128..1024, m = 7 (default)
    bool over2 = false; unsigned long long frames_done = 0;     // firpfbch2 analyzer: hop M/2, generic kernel

    void init(const Ctx &c, unsigned M_, unsigned m_, float As_, bool over2_ = false)
    {
        M = M_; m = m_; As = As_; P = 2 * m; over2 = over2_;
        h = over2 ? design::design_firpfbch2(M, m, As) : design::design_firpfbch(M, m, As);
        hd.ensure(h.size() * sizeof(float));
        CK(cudaMemcpyAsync(hd.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
        std::vector<float2> t(M);
        for (unsigned i = 0; i < M; i++) {
            t[i].x = (float)std::cos(-2.0 * design::kPi * (double)i / (double)M);
            t[i].y = (float)std::sin(-2.0 * design::kPi * (double)i / (double)M);
        }
        tw.ensure(M * sizeof(float2));
        CK(cudaMemcpyAsync(tw.p, t.data(), M * sizeof(float2), cudaMemcpyHostToDevice, c.stream));
        log2M = -1;
        if (M > 1 && (M & (M - 1)) == 0) { log2M = 0; while ((1u << log2M) < M) log2M++; }
        int elems = 8192;                              // F*M complex samples per CTA (64 KB)
        F = std::max(1, elems / (int)M);
        if (F > 256) F = 256;
        smem = (size_t)F * M * sizeof(float2) * (log2M >= 0 ? 1 : 2);
        if (smem > 200 * 1024) throw CudaError{"firpfbch: channel count too large for one CTA tile"};
        raise_dyn_smem(k_pfb, smem);
        tile_kernel = nullptr;
        if (log2M >= (over2 ? 2 : 1) && (int)M <= kPfbTileMaxM && (int)P == kPfbTileP) {
            switch (log2M) {
            case 1: tile_kernel = k_pfb_tile<1>; break;
            case 2: tile_kernel = k_pfb_tile<2>; break;
            case 3: tile_kernel = k_pfb_tile<3>; break;
            case 4: tile_kernel = k_pfb_tile<4>; break;
            default: tile_kernel = k_pfb_tile<5>; break;
            }
            // M = 8, 16: two frames per thread (half the shared-memory reads); option CSDR_OPT_PFB_VARIANT = 2 keeps k_pfb_tile
            tile_two = (log2M == 3 || log2M == 4) && g_options[CSDR_OPT_PFB_VARIANT] != 2;
            if (tile_two) tile_kernel = log2M == 3 ? k_pfb_tile2<3> : k_pfb_tile2<4>;
            tp = PfbTileParams{};
            for (unsigned i = 0; i < M / 2; i++) tp.tw[i] = t[i];
            for (unsigned k = 0; k < P; k++) for (unsigned n = 0; n < M; n++) tp.h[k * M + n] = h[(M - 1 - n) + k * M];
            tile_smem = (size_t)(kPfbTileF + kPfbTileP - 1) * (M + (tile_two ? 0 : 2)) * sizeof(float2);
            raise_dyn_smem(tile_kernel, tile_smem);
        } else if (log2M < 0 && M % 2 == 0 && M <= 24 && (int)P == kPfbTileP && (!over2 || M % 4 == 0)) {
            // even channel counts that are not powers of two (the reference's published run: 20): register DFT by one radix-2
            // split; firpfbch2 needs xr + M/2 on a 16-byte boundary
            switch (M) {
            case 6: tile_kernel = k_pfb_tile_any<6>; break;
            case 10: tile_kernel = k_pfb_tile_any<10>; break;
            case 12: tile_kernel = k_pfb_tile_any<12>; break;
            case 14: tile_kernel = k_pfb_tile_any<14>; break;
            case 18: tile_kernel = k_pfb_tile_any<18>; break;
            case 20: tile_kernel = k_pfb_tile_any<20>; break;
            case 22: tile_kernel = k_pfb_tile_any<22>; break;
            default: tile_kernel = k_pfb_tile_any<24>; break;
            }
            tile_two = false;
            tp = PfbTileParams{};
            for (unsigned i = 0; i < M / 2; i++) tp.tw[i] = t[i];
            for (unsigned k = 0; k < P; k++) for (unsigned n = 0; n < M; n++) tp.h[k * M + n] = h[(M - 1 - n) + k * M];
            tile_smem = (size_t)(kPfbTileF + kPfbTileP - 1) * (M + 2) * sizeof(float2);
            raise_dyn_smem(tile_kernel, tile_smem);
        }
        ring_ok = !over2 && log2M >= 7 && M <= 1024 && (int)P == kPfbRingP;
        if (ring_ok) {
            ring_kernel = pfb_ring_lfz(log2M) == 4 ? k_pfb_ring<4> : k_pfb_ring<5>;
            raise_dyn_smem(ring_kernel, pfb_ring_smem((int)M, log2M));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ring_ctas, ring_kernel, 2 * (int)M / kPfbRingCPT, pfb_ring_smem((int)M, log2M)));
            if (ring_ctas < 1) ring_ok = false;
        }
        stream_ok = log2M >= 7 && M <= 1024 && (int)P == kPfbStP && (over2 || g_options[CSDR_OPT_PFB_VARIANT] == 0);
        if (stream_ok) {
            std::vector<unsigned short> pm(M);
            pfb_stream_perm((int)M, pm.data());
            perm.ensure(M * sizeof(unsigned short));
            CK(cudaMemcpyAsync(perm.p, pm.data(), M * sizeof(unsigned short), cudaMemcpyHostToDevice, c.stream));
            stream_kernel = log2M == 7 ? k_pfb_stream<7> : log2M == 8 ? k_pfb_stream<8> : log2M == 9 ? k_pfb_stream<9> : k_pfb_stream<10>;
            raise_dyn_smem(stream_kernel, pfb_stream_smem((int)M));
            stream_pair = nullptr;
            if (over2 && g_options[CSDR_OPT_PFB_VARIANT] != 3) {
                stream_pair = log2M == 7 ? k_pfb_stream<7, true> : log2M == 8 ? k_pfb_stream<8, true> : log2M == 9 ? k_pfb_stream<9, true> : k_pfb_stream<10, true>;
                raise_dyn_smem(stream_pair, pfb_stream_smem((int)M));
            }
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&stream_ctas, stream_kernel, (int)M, pfb_stream_smem((int)M)));
            if (stream_ctas < 1) stream_ok = false;
            c.sync();
        }
        size_t hb = hist_samples() * sizeof(float2);
        for (auto &b : xr) { b.ensure(hb); CK(cudaMemsetAsync(b.p, 0, b.cap, c.stream)); }
        c.sync();
    }
    unsigned hop() const { return over2 ? M / 2 : M; }
    size_t hist_samples() const { return (size_t)(P - 1) * M + (over2 ? M / 2 : 0); }
    // where the caller must place n = nf*hop() (pre-rotated) samples before calling run()
    float2 *input_slot(const Ctx &c, size_t n)
    {
        size_t need = (hist_samples() + n) * sizeof(float2);
        if (need > xr[cur].cap) {
            // grow while keeping the carried history
            DevBuf nb; nb.ensure(need);
            CK(cudaMemcpyAsync(nb.p, xr[cur].p, hist_samples() * sizeof(float2), cudaMemcpyDeviceToDevice, c.stream));
            c.sync();
            std::swap(xr[cur].p, nb.p); std::swap(xr[cur].cap, nb.cap);
        }
        return xr[cur].as<float2>() + hist_samples();
    }
    // pw (optional): also write |y|^2 there, [M][pw_stride] (saves the per-channel back end its power pass)
    void run(const Ctx &c, int nf, float2 *y, long long y_stride, float *pw = nullptr, long long pw_stride = 0)
    {
        if (nf <= 0) return;
        PfbParams p{};
        p.pw = pw; p.pw_stride = pw_stride;
        p.xr = xr[cur].as<float2>(); p.y = y; p.y_stride = y_stride;
        p.M = (int)M; p.P = (int)P; p.nf = nf; p.F = F; p.log2M = log2M;
        p.h = hd.as<float>(); p.tw = tw.as<float2>();
        p.hop = (int)hop(); p.over2 = over2 ? 1 : 0; p.parity0 = (int)(frames_done & 1); p.scale = 1.0f / (float)M;
        frames_done += (unsigned long long)nf;
        // firpfbch2 on the fast kernels: its even frames are the critically sampled filterbank on xr, its odd frames the same
        // filterbank on xr + M/2 (window of frame t starts at t M/2); two passes, output columns interleaved, and the sign
        // (-1)^(c t) of the per-channel factor is constant within a pass
        const int npass = (over2 && (stream_ok || tile_kernel)) ? 2 : 1;
        if (npass == 2 && stream_ok && stream_pair) {
            // both passes in one launch: clusters of two CTAs (even / odd frames of a stretch) that exchange their output tiles
            // through distributed shared memory and write whole 128-byte lines
            PfbStreamParams sp{};
            const int nfe = (nf + 1) / 2, nfo = nf / 2;
            sp.xr = p.xr; sp.y = y; sp.y_stride = y_stride; sp.pw = pw; sp.pw_stride = pw_stride; sp.nf = nfe; sp.nf_odd = nfo; sp.M = (int)M; sp.log2M = log2M;
            sp.h = hd.as<float>(); sp.tw = tw.as<float2>(); sp.perm = perm.as<unsigned short>();
            sp.ocs = 2; sp.oco = 0; sp.over2 = 1; sp.sc_even = p.scale;
            sp.sc_odd = (p.parity0 & 1) ? -p.scale : p.scale; sp.sc_odd1 = ((p.parity0 + 1) & 1) ? -p.scale : p.scale;
            const int slots = std::max(1, c.sms * stream_ctas / 2);
            int T = (nfe + slots - 1) / slots;
            T = std::max(2 * kPfbStTF, (T + kPfbStTF - 1) / kPfbStTF * kPfbStTF);
            sp.T = T;
            launch_cluster(stream_pair, dim3(2 * ((nfe + T - 1) / T)), dim3(M), pfb_stream_smem((int)M), c.stream, 2u, sp);
        } else
        for (int e = 0; e < npass; e++) {
            const int nfp = npass == 2 ? (nf - e + 1) / 2 : nf;
            if (nfp <= 0) continue;
            const float2 *xin = p.xr + (npass == 2 ? (size_t)e * (M / 2) : 0);
            const int ocs = npass, oco = e, o2 = npass == 2 ? 1 : 0;
            const float sc_even = p.scale, sc_odd = ((p.parity0 + e) & 1) ? -p.scale : p.scale;
            if (stream_ok) {
                PfbStreamParams sp{};
                sp.xr = xin; sp.y = y; sp.y_stride = y_stride; sp.pw = pw; sp.pw_stride = pw_stride; sp.nf = nfp; sp.M = (int)M; sp.log2M = log2M;
                sp.h = hd.as<float>(); sp.tw = tw.as<float2>(); sp.perm = perm.as<unsigned short>();
                sp.ocs = ocs; sp.oco = oco; sp.over2 = o2; sp.sc_even = sc_even; sp.sc_odd = sc_odd;
                // one wave of CTAs where possible: frames per CTA = nf / (SMs * CTAs per SM), in whole output tiles
                const int slots = std::max(1, c.sms * stream_ctas);
                int T = (nfp + slots - 1) / slots;
                T = std::max(2 * kPfbStTF, (T + kPfbStTF - 1) / kPfbStTF * kPfbStTF);
                sp.T = T;
                launch(stream_kernel, dim3((nfp + T - 1) / T), dim3(M), pfb_stream_smem((int)M), c.stream, sp);
            } else if (ring_ok) {
                PfbRingParams rp{};
                rp.xr = p.xr; rp.y = y; rp.y_stride = y_stride; rp.nf = nf; rp.M = (int)M; rp.log2M = log2M;
                rp.pw = pw; rp.pw_stride = pw_stride;
                rp.h = hd.as<float>(); rp.tw = tw.as<float2>();
                // one wave of CTAs where possible: frames per CTA = nf / (SMs * CTAs per SM), in whole output tiles
                const int slots = std::max(1, c.sms * ring_ctas);
                int T = (nf + slots - 1) / slots;
                T = std::max(2 * kPfbRingTF, (T + kPfbRingTF - 1) / kPfbRingTF * kPfbRingTF);
                rp.T = T;
                launch(ring_kernel, dim3((nf + T - 1) / T), dim3(2 * M / kPfbRingCPT), pfb_ring_smem((int)M, log2M), c.stream, rp);
            } else if (tile_kernel) {
                tp.xr = xin; tp.y = y; tp.y_stride = y_stride; tp.nf = nfp; tp.pw = pw; tp.pw_stride = pw_stride;
                tp.ocs = ocs; tp.oco = oco; tp.over2 = o2; tp.sc_even = sc_even; tp.sc_odd = sc_odd;
                launch(tile_kernel, dim3((nfp + kPfbTileF - 1) / kPfbTileF), dim3(tile_two ? kPfbTileF / 2 : kPfbTileF), tile_smem, c.stream, tp);
            } else {
                launch(k_pfb, dim3((nf + F - 1) / F), dim3(256), smem, c.stream, p);
            }
        }
        int H = (int)hist_samples();
        xr[cur ^ 1].ensure((size_t)H * sizeof(float2));
        launch(k_copy_tail, dim3((H + 255) / 256), dim3(256), 0, c.stream, (const float2 *)xr[cur].as<float2>(),
               xr[cur ^ 1].as<float2>(), (long long)nf * hop(), H);
        cur ^= 1;
    }
};

// ---------------------------------------------------------------------------------- wide-band FM tail (real-valued)
// iirfilt_rrrf in second-order sections: per section three launches (wbfm.cuh); state (v1, v2) per lane on the device
struct Iir2Filter {
    std::vector<design::Sos> sos; int lanes = 1;
    DevBuf state, qtab, segz, segin;
    void init(const Ctx &c, const std::vector<design::Sos> &s, int lanes_)
    {
        sos = s; lanes = lanes_;
        state.ensure(sizeof(float2) * sos.size() * lanes);
        CK(cudaMemsetAsync(state.p, 0, sizeof(float2) * sos.size() * lanes, c.stream));
        std::vector<double> q(sos.size() * 128);
        for (size_t i = 0; i < sos.size(); i++) iir2_q_table(sos[i].a[1], sos[i].a[2], q.data() + i * 128);
        qtab.ensure(q.size() * sizeof(double));
        CK(cudaMemcpyAsync(qtab.p, q.data(), q.size() * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        c.sync();
    }
    // y may be x (every thread reads its samples before it writes them)
    void run(const Ctx &c, const float *x, long long x_stride, float *y, long long y_stride, int n)
    {
        if (n <= 0) return;
        for (size_t i = 0; i < sos.size(); i++) {
            Iir2Params p{};
            p.x = i ? y : x; p.x_stride = i ? y_stride : x_stride; p.y = y; p.y_stride = y_stride;
            iir2_plan(p, sos[i].b, sos[i].a, n);
            segz.ensure(sizeof(double2) * (size_t)p.nseg * lanes);
            segin.ensure(sizeof(double2) * (size_t)p.nseg * lanes);
            p.seg_z = segz.as<double2>(); p.seg_in = segin.as<double2>();
            p.Q = qtab.as<double>() + i * 128;
            p.state = state.as<float2>() + i * lanes;
            iir2_launch(StreamLauncher{c.stream}, p, lanes);
        }
    }
};

// firdecim_rrrf as a stream: whole blocks of M samples are consumed as they arrive.  Device layout of a call, per
// lane:  z = [ Lh-1 samples of history | < M pending | the call's new samples ]; the producer writes the new samples
// straight into input_slot().
struct FirDecimator {
    unsigned M = 1; int Lh = 0, lanes = 1;
    DevBuf h, hist, z; long long z_stride = 0; size_t fill = 0;
    long long hist_stride() const { return (long long)Lh - 1 + M; }
    void init(const Ctx &c, unsigned M_, unsigned m, float As, int lanes_)
    {
        M = M_; lanes = lanes_;
        std::vector<float> taps = design::design_firdecim(M, m, As);
        Lh = (int)taps.size();
        h.ensure(taps.size() * sizeof(float));
        CK(cudaMemcpyAsync(h.p, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice, c.stream));
        hist.ensure(sizeof(float) * hist_stride() * lanes);
        CK(cudaMemsetAsync(hist.p, 0, sizeof(float) * hist_stride() * lanes, c.stream));
        c.sync();
    }
    float *input_slot(const Ctx &c, int n, long long *stride)
    {
        const int head = Lh - 1 + (int)fill;
        z_stride = ((long long)head + n + 3) & ~3LL;
        z.ensure(sizeof(float) * (size_t)z_stride * lanes);
        launch(k_rows_copy, dim3((head + 255) / 256, lanes), dim3(256), 0, c.stream, (const float *)hist.as<float>(), hist_stride(), 0LL,
               z.as<float>(), z_stride, 0LL, head);
        *stride = z_stride;
        return z.as<float>() + head;
    }
    // after n new samples were written to input_slot(): y[lane][0 .. blocks) ; returns blocks
    int run(const Ctx &c, int n, float *y, long long y_stride)
    {
        const size_t tot = fill + (size_t)n;
        const int nb = (int)(tot / M);
        FirDecimParams p{};
        p.z = z.as<float>(); p.z_stride = z_stride; p.y = y; p.y_stride = y_stride; p.h = h.as<float>();
        p.nout = nb; p.M = (int)M; p.Lh = Lh;
        firdecim_launch(StreamLauncher{c.stream}, p, lanes, c.sms * 8);
        fill = tot - (size_t)nb * M;
        const int keep = Lh - 1 + (int)fill;
        launch(k_rows_copy, dim3((keep + 255) / 256, lanes), dim3(256), 0, c.stream, (const float *)z.as<float>(), z_stride,
               (long long)nb * M, hist.as<float>(), hist_stride(), 0LL, keep);
        return nb;
    }
};

// wbFMDemodulator's tail after freqdem (Liquid.chs:652-656): de-emphasis, then the output decimator
struct WbfmTail {
    Iir2Filter deemph; FirDecimator dec;
    void init(const Ctx &c, double quad_rate, unsigned decim, int lanes)
    {
        deemph.init(c, design::butter_lowpass_sos(2, (float)(5000.0 / quad_rate)), lanes);
        dec.init(c, decim, 10, 60.0f, lanes);
    }
    size_t max_out(size_t n) const { return (n + dec.fill) / dec.M; }
    // fm: [lanes][n] demodulated samples; y: [lanes][>= max_out(n)]
    int run(const Ctx &c, const float *fm, long long fm_stride, int n, float *y, long long y_stride)
    {
        long long zs = 0;
        float *slot = dec.input_slot(c, n, &zs);
        deemph.run(c, fm, fm_stride, slot, zs, n);
        return dec.run(c, n, y, y_stride);
    }
};

}  // namespace

// ------------------------------------------------------------------------------------------ handles
// Every handle starts with a tag word: the liquid-named alias library (liquid_compat.c) uses it to tell a handle of this
// library from an object created by the real libliquid where a liquid family has constructors on both sides
// (iirfilt_crcf_create_prototype stays liquid's, Liquid.chs:553-571), and the C ABI uses it to reject foreign pointers.
constexpr uint64_t kTagBase = 0x4353445242323030ULL;      // "CSDRB200"
enum { TAG_NCO = 1, TAG_MSRESAMP, TAG_IIRFILT, TAG_FIRPFBCH, TAG_FIRPFBCH2, TAG_AGC, TAG_FREQDEM, TAG_AMPMODEM, TAG_IIRFILT_RRRF,
       TAG_FIRDECIM, TAG_CHAIN };
struct Tagged { uint64_t tag; explicit Tagged(int kind) : tag(kTagBase + (uint64_t)kind) {} ~Tagged() { tag = 0; } };
struct csdr_nco_s : Tagged {
    Ctx ctx; Staging st; int type; uint32_t theta = 0, dtheta = 0;
    float pll_alpha = 0.1f, pll_beta = 0.31622776f;        // NCO_PLL_BANDWIDTH_DEFAULT = 0.1, beta = sqrt(alpha) (liquid nco.c)
    csdr_nco_s(int t) : Tagged(TAG_NCO), ctx(-1), type(t) {}
    int quantize() const { return (type == 1 && g_options[CSDR_OPT_VCO_DIRECT]) ? 0 : 1; }
};
struct csdr_msresamp_s : Tagged { Ctx ctx; Staging st; Frontend fe; csdr_msresamp_s() : Tagged(TAG_MSRESAMP), ctx(-1) {} };
struct csdr_iirfilt_s : Tagged { Ctx ctx; Staging st; Backend be; float alpha; csdr_iirfilt_s() : Tagged(TAG_IIRFILT), ctx(-1) {} };
struct csdr_firpfbch_s : Tagged { Ctx ctx; Staging st; Channelizer ch; DevBuf tmp; csdr_firpfbch_s() : Tagged(TAG_FIRPFBCH), ctx(-1) {} };
struct csdr_firpfbch2_s : Tagged { Ctx ctx; Staging st; Channelizer ch; csdr_firpfbch2_s() : Tagged(TAG_FIRPFBCH2), ctx(-1) {} };
struct csdr_agc_s : Tagged {
    Ctx ctx; Staging st; Backend be; bool started = false;
    float bw = 1e-2f, g = 1.0f, thr = 0.0f; unsigned timeout = 100; int mode = SQ_DISABLED;
    csdr_agc_s() : Tagged(TAG_AGC), ctx(-1) {}
};
struct csdr_freqdem_s : Tagged { Ctx ctx; Staging st; float kf, ref; float2 prev; csdr_freqdem_s() : Tagged(TAG_FREQDEM), ctx(-1) { prev.x = prev.y = 0; } };
struct csdr_ampmodem_s : Tagged { Ctx ctx; Staging st; AmDemod am; csdr_ampmodem_s() : Tagged(TAG_AMPMODEM), ctx(-1) {} };
struct csdr_iirfilt_rrrf_s : Tagged { Ctx ctx; Staging st; Iir2Filter f; csdr_iirfilt_rrrf_s() : Tagged(TAG_IIRFILT_RRRF), ctx(-1) {} };
struct csdr_firdecim_s : Tagged { Ctx ctx; Staging st; FirDecimator d; csdr_firdecim_s() : Tagged(TAG_FIRDECIM), ctx(-1) {} };

namespace {
// NULL / foreign handle: report instead of dereferencing (create() returns NULL on failure, e.g. without a CUDA device)
bool bad_handle(const void *q, int kind, const char *what)
{
    if (q && static_cast<const Tagged *>(q)->tag == kTagBase + (uint64_t)kind) return false;
    set_err(std::string(what) + (q ? ": not a handle of this library (or already destroyed)" : ": NULL handle"));
    return true;
}
}  // namespace
#define REQUIRE(q, kind, ret) do { if (bad_handle((q), (kind), __func__)) return ret; } while (0)

#define API_BEGIN clear_err(); try {
#define API_END(ret_fail)                                                    \
    } catch (const CudaError &e) { set_err(e.msg); return ret_fail; }        \
      catch (const std::exception &e) { set_err(e.what()); return ret_fail; }
#define API_END_VOID                                                         \
    } catch (const CudaError &e) { set_err(e.msg); }                         \
      catch (const std::exception &e) { set_err(e.what()); }

extern "C" {

const char *csdr_version(void) { return "csdr_b200 0.1 (sm_100a)"; }
const char *csdr_last_error(void) { return t_err.c_str(); }
int csdr_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
int csdr_set_device(int device)
{
    API_BEGIN CK(cudaSetDevice(device)); return 0; API_END(-1)
}
void *csdr_host_alloc(size_t bytes)
{
    API_BEGIN void *p = nullptr; CK(cudaHostAlloc(&p, bytes, cudaHostAllocDefault)); return p; API_END(nullptr)
}
void csdr_host_free(void *p) { if (p) cudaFreeHost(p); }
uint64_t csdr_kernel_launches(void) { return g_launches.load(); }
int csdr_synchronize(void) { API_BEGIN CK(cudaDeviceSynchronize()); return 0; API_END(-1) }
int csdr_set_option(int opt, int value) { if (opt < 0 || opt >= 16) return -1; g_options[opt] = value; return 0; }
int csdr_get_option(int opt) { return (opt < 0 || opt >= 16) ? -1 : g_options[opt]; }

// ---------------------------------------------------------------- nco_crcf
csdr_nco csdr_nco_crcf_create(int type)
{
    API_BEGIN return new csdr_nco_s(type); API_END(nullptr)
}
void csdr_nco_crcf_destroy(csdr_nco q) { if (!q) return; REQUIRE(q, TAG_NCO, ); delete q; }
void csdr_nco_crcf_print(csdr_nco q)
{
    REQUIRE(q, TAG_NCO, );
    if (q) printf("nco [phase: 0x%.8x rad, freq: 0x%.8x rad/sample]\n", q->theta, q->dtheta);
}
void csdr_nco_crcf_set_frequency(csdr_nco q, float dtheta) { REQUIRE(q, TAG_NCO, ); q->dtheta = design::nco_constrain(dtheta); }
void csdr_nco_crcf_set_phase(csdr_nco q, float theta) { REQUIRE(q, TAG_NCO, ); q->theta = design::nco_constrain(theta); }
uint32_t csdr_nco_crcf_get_phase_word(csdr_nco q) { REQUIRE(q, TAG_NCO, 0); return q->theta; }
uint32_t csdr_nco_crcf_get_freq_word(csdr_nco q) { REQUIRE(q, TAG_NCO, 0); return q->dtheta; }
static void nco_mix(csdr_nco q, const csdr_cf32 *x, csdr_cf32 *y, unsigned n, int up)
{
    API_BEGIN
    if (!n) return;
    q->ctx.use();
    size_t bytes = (size_t)n * sizeof(float2);
    const float2 *xd = (const float2 *)q->st.to_dev(q->ctx, x, bytes);
    float2 *yd = (float2 *)q->st.out_dev(y, bytes);
    launch(k_nco_mix, dim3(grid_for(n, 256, q->ctx.sms)), dim3(256), 0, q->ctx.stream, xd, yd, (long long)n, q->theta,
           q->dtheta, q->quantize(), up);
    q->theta += (uint32_t)n * q->dtheta;
    q->st.finish(q->ctx, y, yd, bytes);
    API_END_VOID
}
// ---- the scalar members of liquid's nco_crcf family (nco.c, uint32 phase): the reference's stereo-FM pilot PLL
// (pllCreate / pllStep, Liquid.chs:959-988) calls them on handles made by nco_crcf_create, so a library that
// answers to nco_crcf_create has to answer to all of them.  Pure host arithmetic on the handle's phase / frequency words.
// NCO(_get_phase): 2.0f*M_PI*(float)theta / (float)(1LLU<<32), evaluated in double, returned as float
static float nco_word_to_rad(uint32_t w) { return (float)(2.0 * design::kPi * (double)(float)w / 4294967296.0); }
void csdr_nco_crcf_adjust_frequency(csdr_nco q, float df) { REQUIRE(q, TAG_NCO, ); q->dtheta += design::nco_constrain(df); }
void csdr_nco_crcf_adjust_phase(csdr_nco q, float dphi) { REQUIRE(q, TAG_NCO, ); q->theta += design::nco_constrain(dphi); }
void csdr_nco_crcf_step(csdr_nco q) { REQUIRE(q, TAG_NCO, ); q->theta += q->dtheta; }
void csdr_nco_crcf_reset(csdr_nco q) { REQUIRE(q, TAG_NCO, ); q->theta = 0; q->dtheta = 0; }
float csdr_nco_crcf_get_phase(csdr_nco q) { REQUIRE(q, TAG_NCO, 0.0f); return nco_word_to_rad(q->theta); }
float csdr_nco_crcf_get_frequency(csdr_nco q)
{
    REQUIRE(q, TAG_NCO, 0.0f);
    const float d = nco_word_to_rad(q->dtheta);
    return d > (float)design::kPi ? d - 2.0f * (float)design::kPi : d;
}
void csdr_nco_crcf_cexpf(csdr_nco q, csdr_cf32 *y)
{
    REQUIRE(q, TAG_NCO, );
    if (!y) return;
    if (q->quantize()) {
        // NCO(_index): round the phase to 1024 table levels; sintab[i] = sinf(2 pi i / 1024), cos = sintab[i + 256]
        const unsigned i = ((q->theta + (1u << 21)) >> 22) & 0x3ffu;
        y->im = sinf((float)(2.0 * design::kPi * (double)i / 1024.0));
        y->re = sinf((float)(2.0 * design::kPi * (double)((i + 256u) & 0x3ffu) / 1024.0));
    } else {
        const float th = nco_word_to_rad(q->theta);
        y->re = cosf(th); y->im = sinf(th);
    }
}
void csdr_nco_crcf_pll_set_bandwidth(csdr_nco q, float bw)
{
    REQUIRE(q, TAG_NCO, );
    if (bw < 0.0f) { set_err("nco_crcf_pll_set_bandwidth: bandwidth must be positive"); return; }
    q->pll_alpha = bw; q->pll_beta = sqrtf(bw);
}
void csdr_nco_crcf_pll_step(csdr_nco q, float dphi)
{
    REQUIRE(q, TAG_NCO, );
    q->dtheta += design::nco_constrain(dphi * q->pll_alpha);      // adjust_frequency
    q->theta += design::nco_constrain(dphi * q->pll_beta);        // adjust_phase
}
int csdr_handle_kind(const void *h)
{
    if (!h) return 0;
    const uint64_t t = static_cast<const Tagged *>(h)->tag;
    return (t > kTagBase && t <= kTagBase + TAG_CHAIN) ? (int)(t - kTagBase) : 0;
}
void csdr_nco_crcf_mix_block_down(csdr_nco q, const csdr_cf32 *x, csdr_cf32 *y, unsigned n) { REQUIRE(q, TAG_NCO, ); nco_mix(q, x, y, n, 0); }
void csdr_nco_crcf_mix_block_up(csdr_nco q, const csdr_cf32 *x, csdr_cf32 *y, unsigned n) { REQUIRE(q, TAG_NCO, ); nco_mix(q, x, y, n, 1); }

// ---------------------------------------------------------------- msresamp_crcf
csdr_msresamp csdr_msresamp_crcf_create(float r, float As)
{
    API_BEGIN
    if (!(r > 0.0f)) throw CudaError{"msresamp_crcf_create: rate must be positive"};
    std::unique_ptr<csdr_msresamp_s> q(new csdr_msresamp_s());
    q->fe.init(q->ctx, r, As, 1);
    return q.release();
    API_END(nullptr)
}
void csdr_msresamp_crcf_destroy(csdr_msresamp q) { if (!q) return; REQUIRE(q, TAG_MSRESAMP, ); delete q; }
float csdr_msresamp_crcf_get_rate(csdr_msresamp q) { REQUIRE(q, TAG_MSRESAMP, 0.0f); return q->fe.ms.rate; }
void csdr_msresamp_crcf_print(csdr_msresamp q)
{
    REQUIRE(q, TAG_MSRESAMP, );
    const auto &ms = q->fe.ms;
    printf("multi-stage resampler (csdr_b200)\n  composite rate      : %12.10f\n", ms.rate);
    printf("  type                : %s\n", ms.interp ? "interp" : "decim");
    printf("  num halfband stages : %u (rate 2^-%u)\n", ms.S, ms.S);
    for (unsigned s = 0; s < ms.S; s++) printf("    stage[%u] m = %u (%u taps)\n", s, ms.st[s].m, 4 * ms.st[s].m + 1);
    printf("  arbitrary resampler : rate %12.10f, npfb %u, m %u, step 0x%08x\n", ms.rate_arb, ms.npfb, ms.m_arb, ms.step);
}
void csdr_msresamp_crcf_execute(csdr_msresamp q, const csdr_cf32 *x, unsigned nx, csdr_cf32 *y, unsigned *ny)
{
    REQUIRE(q, TAG_MSRESAMP, );
    if (ny) *ny = 0;
    API_BEGIN
    q->ctx.use();
    const float2 *xd = (const float2 *)q->st.to_dev(q->ctx, x, (size_t)nx * sizeof(float2));
    size_t cap = (size_t)q->fe.max_out(nx) * sizeof(float2);
    float2 *yd = (float2 *)q->st.out_dev(y, cap);
    long long n = q->fe.run(q->ctx, xd, nx, 0, yd, 0);
    q->st.finish(q->ctx, y, yd, (size_t)n * sizeof(float2));
    if (ny) *ny = (unsigned)n;
    API_END_VOID
}
unsigned csdr_msresamp_num_stages(csdr_msresamp q) { REQUIRE(q, TAG_MSRESAMP, 0); return q->fe.ms.S; }
unsigned csdr_msresamp_stage_m(csdr_msresamp q, unsigned s) { REQUIRE(q, TAG_MSRESAMP, 0); return s < q->fe.ms.S ? q->fe.ms.st[s].m : 0; }
int csdr_msresamp_stage_taps(csdr_msresamp q, unsigned s, float *h1)
{
    REQUIRE(q, TAG_MSRESAMP, -1);
    if (s >= q->fe.ms.S) return -1;
    memcpy(h1, q->fe.ms.st[s].h1.data(), q->fe.ms.st[s].h1.size() * sizeof(float));
    return 0;
}
uint32_t csdr_msresamp_resamp_step(csdr_msresamp q) { REQUIRE(q, TAG_MSRESAMP, 0); return q->fe.ms.step; }
int csdr_msresamp_resamp_bank(csdr_msresamp q, float *bank, unsigned *npfb)
{
    REQUIRE(q, TAG_MSRESAMP, -1);
    if (npfb) *npfb = q->fe.ms.npfb;
    if (bank) memcpy(bank, q->fe.ms.bank.data(), q->fe.ms.bank.size() * sizeof(float));
    return 0;
}

// ---------------------------------------------------------------- iirfilt_crcf dc blocker
csdr_iirfilt csdr_iirfilt_crcf_create_dc_blocker(float alpha)
{
    API_BEGIN
    std::unique_ptr<csdr_iirfilt_s> q(new csdr_iirfilt_s());
    q->alpha = alpha;
    q->be.has_dc = true; q->be.dc_alpha = alpha;
    q->be.init(q->ctx, 1);
    return q.release();
    API_END(nullptr)
}
void csdr_iirfilt_crcf_destroy(csdr_iirfilt q) { if (!q) return; REQUIRE(q, TAG_IIRFILT, ); delete q; }
void csdr_iirfilt_crcf_print(csdr_iirfilt q)
{
    REQUIRE(q, TAG_IIRFILT, );
    printf("iir filter [normal]:\n  b :   %12.8f %12.8f\n  a :   %12.8f %12.8f\n", 1.0f, -1.0f, 1.0f, -1.0f + q->alpha);
}
void csdr_iirfilt_crcf_execute_block(csdr_iirfilt q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y)
{
    REQUIRE(q, TAG_IIRFILT, );
    API_BEGIN
    if (!n) return;
    q->ctx.use();
    size_t bytes = (size_t)n * sizeof(float2);
    const float2 *xd = (const float2 *)q->st.to_dev(q->ctx, x, bytes);
    float2 *yd = (float2 *)q->st.out_dev(y, bytes);
    q->be.run_dc_only(q->ctx, xd, 0, yd, 0, (int)n);
    q->st.finish(q->ctx, y, yd, bytes);
    API_END_VOID
}

// ---------------------------------------------------------------- firpfbch_crcf
csdr_firpfbch csdr_firpfbch_crcf_create_kaiser(int type, unsigned M, unsigned m, float As)
{
    API_BEGIN
    if (type != 0) throw CudaError{"firpfbch_crcf_create_kaiser: only LIQUID_ANALYZER (0) is implemented"};
    if (M < 2 || m < 1) throw CudaError{"firpfbch_crcf_create_kaiser: need M >= 2 and m >= 1"};
    std::unique_ptr<csdr_firpfbch_s> q(new csdr_firpfbch_s());
    q->ch.init(q->ctx, M, m, As);
    return q.release();
    API_END(nullptr)
}
void csdr_firpfbch_crcf_destroy(csdr_firpfbch q) { if (!q) return; REQUIRE(q, TAG_FIRPFBCH, ); delete q; }
void csdr_firpfbch_crcf_print(csdr_firpfbch q)
{
    REQUIRE(q, TAG_FIRPFBCH, );
    printf("firpfbch (analyzer) [%u channels]:\n", q->ch.M);
    for (size_t i = 0; i < q->ch.h.size(); i++) printf("  h[%3zu] = %12.8f + %12.8f*j\n", i, q->ch.h[i], 0.0f);
}
int csdr_firpfbch_taps(csdr_firpfbch q, float *h) { REQUIRE(q, TAG_FIRPFBCH, -1); memcpy(h, q->ch.h.data(), q->ch.h.size() * sizeof(float)); return 0; }
int csdr_firpfbch_execute_block(csdr_firpfbch q, csdr_nco nco, const csdr_cf32 *x, unsigned n, csdr_cf32 *y)
{
    REQUIRE(q, TAG_FIRPFBCH, -1);
    API_BEGIN
    if (!n) return 0;
    q->ctx.use();
    const unsigned M = q->ch.M;
    const unsigned nf = n / M;
    size_t bytes_in = (size_t)n * sizeof(float2), bytes_out = (size_t)nf * M * sizeof(float2);
    const float2 *xd = (const float2 *)q->st.to_dev(q->ctx, x, bytes_in);
    float2 *yd = (float2 *)q->st.out_dev(y, bytes_out);
    // pre-rotate the whole chunk (Liquid.chs:847), tail included, into the channelizer's input slot
    float2 *slot = q->ch.input_slot(q->ctx, (size_t)nf * M + M);
    if (nco) {
        if (bad_handle(nco, TAG_NCO, "csdr_firpfbch_execute_block (nco)")) return -1;
        launch(k_nco_mix, dim3(grid_for(n, 256, q->ctx.sms)), dim3(256), 0, q->ctx.stream, xd, slot, (long long)n,
               nco->theta, nco->dtheta, nco->quantize(), 0);
        nco->theta += (uint32_t)n * nco->dtheta;
    } else {
        CK(cudaMemcpyAsync(slot, xd, bytes_in, cudaMemcpyDeviceToDevice, q->ctx.stream));
    }
    q->ch.run(q->ctx, (int)nf, yd, nf);
    q->st.finish(q->ctx, y, yd, bytes_out);
    return 0;
    API_END(-1)
}
void csdr_firpfbch_crcf_analyzer_execute(csdr_firpfbch q, const csdr_cf32 *x, csdr_cf32 *y)
{
    REQUIRE(q, TAG_FIRPFBCH, );
    csdr_firpfbch_execute_block(q, nullptr, x, q->ch.M, y);
}

// ---------------------------------------------------------------- firpfbch2_crcf (2x oversampled analyzer)
csdr_firpfbch2 csdr_firpfbch2_crcf_create_kaiser(int type, unsigned M, unsigned m, float As)
{
    API_BEGIN
    if (type != 0) throw CudaError{"firpfbch2_crcf_create_kaiser: only LIQUID_ANALYZER (0) is implemented"};
    if (M < 2 || (M & 1) || m < 1) throw CudaError{"firpfbch2_crcf_create_kaiser: need an even M >= 2 and m >= 1"};
    std::unique_ptr<csdr_firpfbch2_s> q(new csdr_firpfbch2_s());
    q->ch.init(q->ctx, M, m, As, true);
    return q.release();
    API_END(nullptr)
}
void csdr_firpfbch2_crcf_destroy(csdr_firpfbch2 q) { if (!q) return; REQUIRE(q, TAG_FIRPFBCH2, ); delete q; }
void csdr_firpfbch2_crcf_print(csdr_firpfbch2 q)
{
    REQUIRE(q, TAG_FIRPFBCH2, );
    printf("firpfbch2_crcf: analyzer, channels: %u, semi-length: %u, %zu taps\n", q->ch.M, q->ch.m, q->ch.h.size());
}
int csdr_firpfbch2_taps(csdr_firpfbch2 q, float *h) { REQUIRE(q, TAG_FIRPFBCH2, -1); memcpy(h, q->ch.h.data(), q->ch.h.size() * sizeof(float)); return 0; }
int csdr_firpfbch2_execute_block(csdr_firpfbch2 q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y)
{
    REQUIRE(q, TAG_FIRPFBCH2, -1);
    API_BEGIN
    q->ctx.use();
    const unsigned M = q->ch.M, M2 = M / 2, nf = n / M2;
    if (!nf) return 0;
    const size_t bytes_in = (size_t)nf * M2 * sizeof(float2), bytes_out = (size_t)nf * M * sizeof(float2);
    float2 *yd = (float2 *)q->st.out_dev(y, bytes_out);
    float2 *slot = q->ch.input_slot(q->ctx, (size_t)nf * M2);
    CK(cudaMemcpyAsync(slot, x, bytes_in, cudaMemcpyDefault, q->ctx.stream));
    q->ch.run(q->ctx, (int)nf, yd, nf);
    q->st.finish(q->ctx, y, yd, bytes_out);
    return 0;
    API_END(-1)
}
void csdr_firpfbch2_crcf_execute(csdr_firpfbch2 q, const csdr_cf32 *x, csdr_cf32 *y)
{
    REQUIRE(q, TAG_FIRPFBCH2, );
    csdr_firpfbch2_execute_block(q, x, q->ch.M / 2, y);
}

// ---------------------------------------------------------------- agc_crcf
csdr_agc csdr_agc_crcf_create(void)
{
    API_BEGIN return new csdr_agc_s(); API_END(nullptr)
}
void csdr_agc_crcf_destroy(csdr_agc q) { if (!q) return; REQUIRE(q, TAG_AGC, ); delete q; }
static void agc_push_config(csdr_agc q)
{
    // (re)build the device state from the host-side configuration; called lazily before the first execute
    if (q->started) return;
    q->be.has_dc = false; q->be.has_agc = true; q->be.demod = 0;
    q->be.agc_bw = q->bw; q->be.agc_thr = q->thr; q->be.agc_timeout = q->timeout; q->be.squelch = q->mode != SQ_DISABLED;
    q->be.init(q->ctx, 1, q->g, q->mode);
    q->started = true;
}
static LaneState agc_lane(csdr_agc q) { agc_push_config(q); return q->be.read_lane(q->ctx, 0); }
void csdr_agc_crcf_print(csdr_agc q)
{
    REQUIRE(q, TAG_AGC, );
    API_BEGIN
    LaneState l = agc_lane(q);
    printf("agc [rssi: %12.4f dB, output gain: %.3f dB, bw: %12.4e, locked: no, squelch: %s]:\n",
           -20 * log10((double)l.g), 0.0, q->bw, q->mode == SQ_DISABLED ? "disabled" : "enabled");
    API_END_VOID
}
void csdr_agc_crcf_set_bandwidth(csdr_agc q, float bt) { REQUIRE(q, TAG_AGC, ); q->bw = bt; if (q->started) q->be.agc_bw = bt; }
void csdr_agc_crcf_set_signal_level(csdr_agc q, float x2)
{
    REQUIRE(q, TAG_AGC, );
    API_BEGIN
    q->g = 1.0f / x2;
    if (q->started) { LaneState l = q->be.read_lane(q->ctx, 0); l.g = q->g; l.y2p = 1.0f; q->be.write_lane(q->ctx, 0, l); }
    API_END_VOID
}
void csdr_agc_crcf_squelch_enable(csdr_agc q)
{
    REQUIRE(q, TAG_AGC, );
    API_BEGIN
    q->mode = SQ_ENABLED;
    if (q->started) { q->be.squelch = true; LaneState l = q->be.read_lane(q->ctx, 0); l.mode = SQ_ENABLED; q->be.write_lane(q->ctx, 0, l); }
    API_END_VOID
}
void csdr_agc_crcf_squelch_set_threshold(csdr_agc q, float t) { REQUIRE(q, TAG_AGC, ); q->thr = t; if (q->started) q->be.agc_thr = t; }
void csdr_agc_crcf_squelch_set_timeout(csdr_agc q, unsigned t) { REQUIRE(q, TAG_AGC, ); q->timeout = t; if (q->started) q->be.agc_timeout = t; }
float csdr_agc_crcf_get_rssi(csdr_agc q)
{
    REQUIRE(q, TAG_AGC, 0.0f);
    API_BEGIN LaneState l = agc_lane(q); return (float)(-20 * log10((double)l.g)); API_END(0.0f)
}
int csdr_agc_crcf_squelch_get_status(csdr_agc q)
{
    REQUIRE(q, TAG_AGC, -1);
    API_BEGIN LaneState l = agc_lane(q); return l.mode; API_END(0)
}
static int agc_exec(csdr_agc q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y, bool gate)
{
    API_BEGIN
    if (!n) return 0;
    q->ctx.use();
    agc_push_config(q);
    // liquid's agc_crcf_execute_block does not gate; the gate is the Haskell wrapper's (Liquid.chs:700-704)
    size_t bytes = (size_t)n * sizeof(float2);
    const float2 *xd = (const float2 *)q->st.to_dev(q->ctx, x, bytes);
    float2 *yd = (float2 *)q->st.out_dev(y, bytes);
    q->be.squelch = (q->mode != SQ_DISABLED);
    q->be.gate = gate;
    q->be.run(q->ctx, xd, 0, yd, 0, (int)n);
    q->st.finish(q->ctx, y, yd, bytes);
    return 0;
    API_END(-1)
}
void csdr_agc_crcf_execute_block(csdr_agc q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y) { REQUIRE(q, TAG_AGC, ); agc_exec(q, x, n, y, false); }
int csdr_agc_squelch_execute_block(csdr_agc q, const csdr_cf32 *x, unsigned n, csdr_cf32 *y) { REQUIRE(q, TAG_AGC, -1); return agc_exec(q, x, n, y, true); }

// ---------------------------------------------------------------- freqdem
csdr_freqdem csdr_freqdem_create(float kf)
{
    API_BEGIN
    if (!(kf > 0.0f)) throw CudaError{"freqdem_create: modulation factor must be positive"};
    std::unique_ptr<csdr_freqdem_s> q(new csdr_freqdem_s());
    q->kf = kf; q->ref = (float)(1.0f / (2 * design::kPi * kf));
    return q.release();
    API_END(nullptr)
}
void csdr_freqdem_destroy(csdr_freqdem q) { if (!q) return; REQUIRE(q, TAG_FREQDEM, ); delete q; }
void csdr_freqdem_print(csdr_freqdem q) { REQUIRE(q, TAG_FREQDEM, ); printf("freqdem:\n    mod. factor :   %8.4f\n", q->kf); }
void csdr_freqdem_demodulate_block(csdr_freqdem q, const csdr_cf32 *r, unsigned n, float *m)
{
    REQUIRE(q, TAG_FREQDEM, );
    API_BEGIN
    if (!n) return;
    q->ctx.use();
    const float2 *rd = (const float2 *)q->st.to_dev(q->ctx, r, (size_t)n * sizeof(float2));
    float *md = (float *)q->st.out_dev(m, (size_t)n * sizeof(float));
    launch(k_freqdem, dim3(grid_for(n, 256, q->ctx.sms)), dim3(256), 0, q->ctx.stream, rd, md, (long long)n, q->prev, q->ref);
    CK(cudaMemcpyAsync(&q->prev, rd + (n - 1), sizeof(float2), cudaMemcpyDeviceToHost, q->ctx.stream));
    q->st.finish(q->ctx, m, md, (size_t)n * sizeof(float));
    API_END_VOID
}

// ---------------------------------------------------------------- ampmodem
csdr_ampmodem csdr_ampmodem_create(float mod_index, int type, int suppressed)
{
    API_BEGIN
    if (type != 0 || suppressed != 0) throw CudaError{"ampmodem_create: only DSB with carrier (type 0, suppressed 0) is implemented"};
    std::unique_ptr<csdr_ampmodem_s> q(new csdr_ampmodem_s());
    q->am.init(q->ctx.stream, 1, mod_index, g_options[CSDR_OPT_AMPMODEM_PLL] != 0);
    q->am.spec = g_options[CSDR_OPT_AM_PLL_SEQUENTIAL] == 0;
    CK(cudaStreamSynchronize(q->ctx.stream));
    return q.release();
    API_END(nullptr)
}
void csdr_ampmodem_destroy(csdr_ampmodem q) { if (!q) return; REQUIRE(q, TAG_AMPMODEM, ); delete q; }
void csdr_ampmodem_print(csdr_ampmodem q)
{
    REQUIRE(q, TAG_AMPMODEM, );
    printf("ampmodem:\n    type            :   double side-band\n    supp. carrier   :   no\n    mod. index      :   %-8.4f\n", q->am.mod_index);
}
void csdr_ampmodem_demodulate_block(csdr_ampmodem q, const csdr_cf32 *r, unsigned n, float *m)
{
    REQUIRE(q, TAG_AMPMODEM, );
    API_BEGIN
    if (!n) return;
    q->ctx.use();
    const float2 *rd = (const float2 *)q->st.to_dev(q->ctx, r, (size_t)n * sizeof(float2));
    float *md = (float *)q->st.out_dev(m, (size_t)n * sizeof(float));
    q->am.run(q->ctx.stream, rd, 0, md, 0, (int)n);
    g_launches.fetch_add(q->am.take_launches());
    q->st.finish(q->ctx, m, md, (size_t)n * sizeof(float));
    API_END_VOID
}

// ---------------------------------------------------------------- iirfilt_rrrf (Butterworth low-pass prototype)
csdr_iirfilt_rrrf csdr_iirfilt_rrrf_create_prototype(int ftype, int btype, int format, unsigned order, float fc, float f0,
                                                     float ap, float as)
{
    API_BEGIN
    (void)f0; (void)ap; (void)as;
    if (ftype != 0 || btype != 0 || format != 0)
        throw CudaError{"iirfilt_rrrf_create_prototype: only Butterworth / low-pass / second-order sections (0, 0, 0) is implemented"};
    if (order < 1 || order > 16) throw CudaError{"iirfilt_rrrf_create_prototype: order must be in [1, 16]"};
    if (!(fc > 0.0f && fc < 0.5f)) throw CudaError{"iirfilt_rrrf_create_prototype: cutoff must be in (0, 0.5)"};
    std::unique_ptr<csdr_iirfilt_rrrf_s> q(new csdr_iirfilt_rrrf_s());
    q->f.init(q->ctx, design::butter_lowpass_sos(order, fc), 1);
    return q.release();
    API_END(nullptr)
}
void csdr_iirfilt_rrrf_destroy(csdr_iirfilt_rrrf q) { if (!q) return; REQUIRE(q, TAG_IIRFILT_RRRF, ); delete q; }
void csdr_iirfilt_rrrf_print(csdr_iirfilt_rrrf q)
{
    REQUIRE(q, TAG_IIRFILT_RRRF, );
    printf("iir filter [sos]:\n");
    for (size_t i = 0; i < q->f.sos.size(); i++) {
        const design::Sos &s = q->f.sos[i];
        printf("  section %zu: b = %12.8f %12.8f %12.8f   a = %12.8f %12.8f %12.8f\n", i, s.b[0], s.b[1], s.b[2], s.a[0], s.a[1], s.a[2]);
    }
}
unsigned csdr_iirfilt_rrrf_coefficients(csdr_iirfilt_rrrf q, float *b, float *a)
{
    REQUIRE(q, TAG_IIRFILT_RRRF, 0);
    for (size_t i = 0; i < q->f.sos.size(); i++)
        for (int j = 0; j < 3; j++) { if (b) b[3 * i + j] = q->f.sos[i].b[j]; if (a) a[3 * i + j] = q->f.sos[i].a[j]; }
    return (unsigned)q->f.sos.size();
}
void csdr_iirfilt_rrrf_execute_block(csdr_iirfilt_rrrf q, const float *x, unsigned n, float *y)
{
    REQUIRE(q, TAG_IIRFILT_RRRF, );
    API_BEGIN
    if (!n) return;
    q->ctx.use();
    const float *xd = (const float *)q->st.to_dev(q->ctx, x, (size_t)n * sizeof(float));
    float *yd = (float *)q->st.out_dev(y, (size_t)n * sizeof(float));
    q->f.run(q->ctx, xd, 0, yd, 0, (int)n);
    q->st.finish(q->ctx, y, yd, (size_t)n * sizeof(float));
    API_END_VOID
}

// ---------------------------------------------------------------- firdecim_rrrf
csdr_firdecim csdr_firdecim_rrrf_create_kaiser(unsigned M, unsigned m, float as)
{
    API_BEGIN
    if (M < 1 || M > 4096) throw CudaError{"firdecim_rrrf_create_kaiser: decimation factor must be in [1, 4096]"};
    if (m < 1 || (size_t)2 * M * m + 1 > 40000) throw CudaError{"firdecim_rrrf_create_kaiser: filter too long"};
    std::unique_ptr<csdr_firdecim_s> q(new csdr_firdecim_s());
    q->d.init(q->ctx, M, m, as, 1);
    return q.release();
    API_END(nullptr)
}
void csdr_firdecim_rrrf_destroy(csdr_firdecim q) { if (!q) return; REQUIRE(q, TAG_FIRDECIM, ); delete q; }
void csdr_firdecim_rrrf_print(csdr_firdecim q) { REQUIRE(q, TAG_FIRDECIM, ); printf("firdecim_rrrf: M = %u, %d taps\n", q->d.M, q->d.Lh); }
void csdr_firdecim_rrrf_execute_block(csdr_firdecim q, const float *x, unsigned n, float *y)
{
    REQUIRE(q, TAG_FIRDECIM, );
    API_BEGIN
    if (!n) return;
    q->ctx.use();
    const size_t nin = (size_t)n * q->d.M;
    if (nin > 0x7fffffffULL) throw CudaError{"firdecim_rrrf_execute_block: block too large"};
    long long zs = 0;
    float *slot = q->d.input_slot(q->ctx, (int)nin, &zs);
    CK(cudaMemcpyAsync(slot, x, nin * sizeof(float), cudaMemcpyDefault, q->ctx.stream));
    float *yd = (float *)q->st.out_dev(y, (size_t)n * sizeof(float));
    q->d.run(q->ctx, (int)nin, yd, 0);
    q->st.finish(q->ctx, y, yd, (size_t)n * sizeof(float));
    API_END_VOID
}

}  // extern "C"

#include "chain.inl"
