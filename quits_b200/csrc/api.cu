// quits_b200/csrc/api.cu -- the C ABI (include/quits_b200.h): contexts, device residency, batching, launches.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/quits_b200.h"
#include "qb_device.h"
#include "qb_host.h"

namespace {

thread_local std::string g_err;

struct cuda_error : std::runtime_error { using std::runtime_error::runtime_error; };
struct arg_error : std::runtime_error { using std::runtime_error::runtime_error; };

#define CK(expr)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            throw cuda_error(std::string(#expr) + ": " + cudaGetErrorString(e__) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

template <typename F>
int guard(F&& f) {
    try {
        f();
        return QB_OK;
    } catch (const qb::value_error& e) { g_err = e.what(); return QB_EVALUE;
    } catch (const qb::unsupported_error& e) { g_err = e.what(); return QB_ENOTIMPL;
    } catch (const cuda_error& e) { g_err = e.what(); return QB_ECUDA;
    } catch (const arg_error& e) { g_err = e.what(); return QB_EARG;
    } catch (const std::bad_alloc&) { g_err = "out of host memory"; return QB_ECUDA;
    } catch (const std::exception& e) { g_err = e.what(); return QB_ECUDA; }
}

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
    ~DevBuf() { if (p) cudaFree(p); }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
};

template <typename T>
void upload(DevBuf& b, const std::vector<T>& v, cudaStream_t st, size_t min_elems = 1) {
    b.ensure(std::max(v.size(), min_elems) * sizeof(T) + 16);
    if (!v.empty()) CK(cudaMemcpyAsync(b.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st));
}

struct EventTimer {          // accumulates CUDA-event time of one kernel class on the launching stream
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    size_t used = 0;
    void begin(cudaStream_t st) {
        if (used == spans.size()) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            spans.emplace_back(a, b);
        }
        CK(cudaEventRecord(spans[used].first, st));
    }
    void end(cudaStream_t st) { CK(cudaEventRecord(spans[used].second, st)); ++used; }
    double collect() {       // call after the stream has been synchronised
        double ms = 0;
        for (size_t i = 0; i < used; ++i) {
            float t = 0;
            CK(cudaEventElapsedTime(&t, spans[i].first, spans[i].second));
            ms += t;
        }
        used = 0;
        return ms;
    }
    ~EventTimer() { for (auto& s : spans) { cudaEventDestroy(s.first); cudaEventDestroy(s.second); } }
};

}  // namespace

struct qb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    // sampler scratch
    DevBuf det_rows, obs_rows, det_bytes, obs_bytes, inj_start, inj_tgt, inj_code, inj_shot, counts;
    DevBuf det_words;          // frame kernel scratch (qb::frame_scratch_bytes)
    EventTimer t_frame, t_total;
    // side streams of the decoder: a batch is decoded as kMaxLanes independent sub-batches so that the latency-bound OSD
    // kernel of one sub-batch shares the SMs with the BP kernel of another
    static constexpr int kMaxLanes = 4;
    cudaStream_t aux[kMaxLanes - 1] = {};
    cudaEvent_t ev_fork = nullptr, ev_join[kMaxLanes - 1] = {};
    // copy stream of the host-buffer entry points: transfers of one chunk overlap the kernels of the next
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_ready[2] = {}, ev_done[2] = {};
    DevBuf det_rows_alt, obs_rows_alt, det_bytes_alt, obs_bytes_alt;
    cudaStream_t copy_stream() {
        if (!copy) {
            CK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
            for (int i = 0; i < 2; ++i) {
                CK(cudaEventCreateWithFlags(&ev_ready[i], cudaEventDisableTiming));
                CK(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
            }
        }
        return copy;
    }
    cudaStream_t lane_stream(int lane) {
        if (lane == 0) return stream;
        if (!aux[lane - 1]) {
            CK(cudaStreamCreateWithFlags(&aux[lane - 1], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ev_join[lane - 1], cudaEventDisableTiming));
        }
        if (!ev_fork) CK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        return aux[lane - 1];
    }
};

struct qb_circuit {
    qb::FlatCircuit fc;
    qb::Tape tape;
    std::vector<int32_t> noise_tape_of_flat;      // flat op -> tape op (noise ops only, else -1)
    // device copies (created on first use with a context)
    int dev = -1;
    DevBuf d_ops, d_targets, d_detptr, d_detidx, d_ctab;
};

struct qb_dem {
    qb::Dem dem;
    qb::CheckMatrix cm;
};

struct qb_plan {
    qb::WindowPlan plan;
};

namespace {

struct WinOwned {
    qb::WinDev dev{};
    DevBuf ss_rec, ss_v0, ss_s0, ss_rsum0, ss_neg0;
    bool serial_slab = false;
    DevBuf colE, llr0f, llr0d, osd_wt, lmask, uptr, uidx, cptr, crow, rptr, rcol, colrec, ptabf, ptabd, rlen, rsum0f, rsum0d, neg0;
    size_t bp_smem = 0;
    bool vglobal = false;
    int bp_grid = 0;            // persistent grid of the VGLOBAL variant (0: one CTA per shot)
    int sort_grid = 0, elim_grid = 0, fast_grid = 0, lsd_grid = 0;
    bool osd_big = false;       // OSD-0 through the slab kernel (more than 768 checks)
};

}  // namespace

struct qb_sw {
    qb_ctx* ctx = nullptr;
    qb::WindowPlan plan;
    qb_bp_opts opts{};
    bool single = false;
    bool use_osd = true;
    bool use_lsd = false;         // BpLsdDecoder post-processing (LSD) instead of OSD
    bool lsd_hi = false;          // lsd_e / lsd_cs with order > 0: candidate sweep inside every cluster
    bool use_slab = false;        // some window runs OSD-0 through the slab kernel (more than 768 checks)
    bool serial = false;          // ldpc schedule='serial'
    bool osd_hi = false;          // osd_e / osd_cs with order > 0: full elimination + candidate sweeps
    int max_iter = 0;
    int precision = 64;
    std::vector<std::unique_ptr<WinOwned>> wins;
    DevBuf alpha;
    // batch state
    int cap = 0;
    int lanes = 1, lanes_used = 1;    // concurrent sub-batches per batch (decode_batch)
    int DW = 0, KW = 0, carryW = 0, synW = 0;
    size_t llr_stride = 0;
    DevBuf det_rows, det_bytes, det_bytes_alt, carry, acc, llr, syn, fail_list, ovf_list, order, sel_key, sel_idx, sel_cnt, counters, stats, pred, ehat, iters, conv, vscratch, lsd_scratch, sort_scratch;
    size_t lsd_slab = 0;
    size_t serial_slab = 0;       // serial schedule: bytes of one CTA's message slab, CTAs of the persistent grid
    int serial_grid = 0;
    int lsd_cols = 0;
    int lsd_grid = 0, lsd_slabs_per_lane = 0;
    EventTimer t_bp, t_osd;
};

namespace {

constexpr size_t kStatSlots = 8;
constexpr size_t kCounterSlots = 8;
constexpr size_t kBatchSlots = 4;          // batches of one fused-run chunk in flight

void use_device(qb_ctx* ctx) { CK(cudaSetDevice(ctx->device)); }

void circuit_to_device(qb_ctx* ctx, qb_circuit* c) {
    if (c->dev == ctx->device) return;
    {
        std::vector<qb::TapeOp> ops(c->tape.ops);
        ops.push_back(qb::TapeOp{});                     // the frame kernel reads one op header ahead
        upload(c->d_ops, ops, ctx->stream);
    }
    upload(c->d_targets, c->tape.targets, ctx->stream, 2);
    upload(c->d_detptr, c->tape.detptr, ctx->stream, 2);
    upload(c->d_detidx, c->tape.detidx, ctx->stream);
    upload(c->d_ctab, c->tape.ctab, ctx->stream, 64);
    CK(cudaStreamSynchronize(ctx->stream));
    c->dev = ctx->device;
}

qb::FrameArgs frame_args(qb_circuit* c, uint64_t seed) {
    qb::FrameArgs a{};
    a.ops = c->d_ops.as<qb::TapeOp>();
    a.n_ops = static_cast<int>(c->tape.ops.size());
    a.targets = c->d_targets.as<uint32_t>();
    a.detptr = c->d_detptr.as<uint32_t>();
    a.detidx = c->d_detidx.as<uint32_t>();
    a.ctab = c->d_ctab.as<uint64_t>();
    a.n_qubits = c->fc.n_qubits;
    a.n_det = c->fc.n_det;
    a.n_obs = c->fc.n_obs;
    a.ring = c->tape.ring;
    a.DW = std::max(1, (c->fc.n_det + 63) / 64);
    a.KW = std::max(1, (c->fc.n_obs + 63) / 64);
    a.seed = seed;
    return a;
}

constexpr uint64_t kSampleChunk = 1ull << 18;      // shots per sampler launch when results go back to the host

// ------------------------------------------------------------------------------------------------ decoder set-up
// GF(2) rank of a window matrix (bit-packed rows, plain elimination); once per window at set-up
int gf2_rank(const qb::Window& hw) {
    const int rows = hw.rows, ncols = hw.ncols, nw = (ncols + 63) / 64;
    std::vector<uint64_t> a(static_cast<size_t>(rows) * nw, 0);
    for (int j = 0; j < ncols; ++j)
        for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) a[static_cast<size_t>(hw.crow[e]) * nw + (j >> 6)] |= 1ull << (j & 63);
    int rank = 0;
    for (int j = 0; j < ncols && rank < rows; ++j) {
        const int wd = j >> 6;
        const uint64_t bit = 1ull << (j & 63);
        int piv = -1;
        for (int i = rank; i < rows; ++i)
            if (a[static_cast<size_t>(i) * nw + wd] & bit) { piv = i; break; }
        if (piv < 0) continue;
        if (piv != rank)
            for (int q = 0; q < nw; ++q) std::swap(a[static_cast<size_t>(piv) * nw + q], a[static_cast<size_t>(rank) * nw + q]);
        for (int i = rank + 1; i < rows; ++i)
            if (a[static_cast<size_t>(i) * nw + wd] & bit)
                for (int q = wd; q < nw; ++q) a[static_cast<size_t>(i) * nw + q] ^= a[static_cast<size_t>(rank) * nw + q];
        ++rank;
    }
    return rank;
}

// ---- serial schedule on the message slab (bp_serial.cu): dependency levels of the column sequence, steps of independent columns
void build_serial_slab(qb_ctx* ctx, const qb::Window& hw, int precision, int method, WinOwned& wo) {
    qb::WinDev& d = wo.dev;
    const int rows = hw.rows, ncols = hw.ncols;
    const size_t nnz = hw.crow.size();
    int cw = 0;
    for (int j = 0; j < ncols; ++j) cw = std::max(cw, static_cast<int>(hw.cptr[j + 1] - hw.cptr[j]));
    if (cw > 16 || rows >= 65535 || ncols >= 65535) return;
    const int lpc = cw <= 6 ? 6 : (cw <= 8 ? 8 : 16), ng = 32 / lpc;
    // CSR position of every CSC edge (rows ascending, columns ascending inside a row: the order the serial sweep visits a row in)
    std::vector<int32_t> rptr(static_cast<size_t>(rows) + 1, 0);
    for (int32_t r : hw.crow) rptr[static_cast<size_t>(r) + 1]++;
    for (int r = 0; r < rows; ++r) rptr[r + 1] += rptr[r];
    std::vector<uint32_t> csr_of(nnz);
    {
        std::vector<int32_t> fill(rptr.begin(), rptr.end() - 1);
        for (int j = 0; j < ncols; ++j)
            for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) csr_of[static_cast<size_t>(e)] = static_cast<uint32_t>(fill[hw.crow[e]]++);
    }
    // level of a column = 1 + the highest level among the earlier columns it shares a row with
    std::vector<int> lastlvl(static_cast<size_t>(rows), 0);
    std::vector<std::vector<int>> bylevel(1);
    for (int j = 0; j < ncols; ++j) {
        int l = 1;
        for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) l = std::max(l, lastlvl[hw.crow[e]] + 1);
        for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) lastlvl[hw.crow[e]] = l;
        if (static_cast<size_t>(l) >= bylevel.size()) bylevel.resize(static_cast<size_t>(l) + 1);
        bylevel[l].push_back(j);
    }
    std::vector<uint32_t> rec;                                // uint2 per lane
    int nsteps = 0;
    auto push_step = [&](const int* cj, int n) {
        for (int lane = 0; lane < 32; ++lane) {
            const int g = lane / lpc, q = lane % lpc;
            uint32_t x = static_cast<uint32_t>(nnz) + static_cast<uint32_t>(lane), row = static_cast<uint32_t>(rows), extra = q == 0 ? 0xFFFFu : 0u;
            if (g < ng && g < n) {
                const int j = cj[g];
                const int wt = static_cast<int>(hw.cptr[j + 1] - hw.cptr[j]);
                if (q < wt) { x = csr_of[static_cast<size_t>(hw.cptr[j] + q)]; row = static_cast<uint32_t>(hw.crow[hw.cptr[j] + q]); }
                if (q == 0) extra = static_cast<uint32_t>(j);
            }
            rec.push_back(x);
            rec.push_back(row | (extra << 16));
        }
        ++nsteps;
    };
    for (size_t l = 1; l < bylevel.size(); ++l)
        for (size_t i = 0; i < bylevel[l].size(); i += static_cast<size_t>(ng))
            push_step(bylevel[l].data() + i, static_cast<int>(std::min<size_t>(static_cast<size_t>(ng), bylevel[l].size() - i)));
    const int real_steps = nsteps;
    for (int k = 0; k < 4; ++k) push_step(nullptr, 0);        // the records in flight past the last step land here
    upload(wo.ss_rec, rec, ctx->stream, 2);
    // initial messages per CSR edge and iteration-0 row summaries
    std::vector<double> v0d(nnz), s0sum(static_cast<size_t>(rows) * 2, DBL_MAX);
    std::vector<float> v0f(nnz), s0sumf(static_cast<size_t>(rows) * 2, FLT_MAX);
    std::vector<uint8_t> neg0(static_cast<size_t>(rows), 0);
    for (int j = 0; j < ncols; ++j) {
        const double l0 = std::log((1.0 - hw.priors[j]) / hw.priors[j]);
        const float l0f = static_cast<float>(l0);
        for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) {
            const int r = hw.crow[e];
            v0d[csr_of[static_cast<size_t>(e)]] = l0;
            v0f[csr_of[static_cast<size_t>(e)]] = l0f;
            if (l0 <= 0.0) neg0[r] ^= 1;
            const double ad = std::fabs(l0);
            const float af = std::fabs(l0f);
            double& m1 = s0sum[2 * r]; double& m2 = s0sum[2 * r + 1];
            m2 = std::min(m2, std::max(m1, ad)); m1 = std::min(m1, ad);
            float& f1 = s0sumf[2 * r]; float& f2 = s0sumf[2 * r + 1];
            f2 = std::min(f2, std::max(f1, af)); f1 = std::min(f1, af);
        }
    }
    const size_t esz = precision == 32 ? 4 : 8;
    wo.ss_v0.ensure((nnz + 32) * esz + 16);
    wo.ss_s0.ensure((nnz + 32) * esz + 16);
    if (precision == 32) {
        if (nnz) CK(cudaMemcpyAsync(wo.ss_v0.p, v0f.data(), nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
        upload(wo.ss_rsum0, s0sumf, ctx->stream, 2);
    } else {
        if (nnz) CK(cudaMemcpyAsync(wo.ss_v0.p, v0d.data(), nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
        upload(wo.ss_rsum0, s0sum, ctx->stream, 2);
    }
    upload(wo.ss_neg0, neg0, ctx->stream, 2);
    CK(cudaStreamSynchronize(ctx->stream));
    d.nnz = static_cast<int>(nnz);
    d.ss_nsteps = real_steps;
    d.ss_lpc = lpc;
    d.ss_rec = wo.ss_rec.as<uint2>();
    d.ss_v0f = wo.ss_v0.as<float>(); d.ss_v0d = wo.ss_v0.as<double>();
    d.ss_s0f = wo.ss_s0.as<float>(); d.ss_s0d = wo.ss_s0.as<double>();
    d.ss_rsum0f = wo.ss_rsum0.as<float2>(); d.ss_rsum0d = wo.ss_rsum0.as<double2>();
    d.ss_neg0 = wo.ss_neg0.as<uint8_t>();
    if (method == 1) {                                        // product-sum: tanh(LLR / 2) per edge and its suffix products, by the device
        CK(qb::launch_serial_slab_ps_tables(d, precision, wo.ss_v0.p, wo.ss_s0.p, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    wo.serial_slab = true;
}

void build_window(qb_ctx* ctx, const qb::Window& hw, int KW, int precision, int method, WinOwned& wo) {
    qb::WinDev& d = wo.dev;
    const int rows = hw.rows, ncols = hw.ncols;
    if (rows <= 0) throw qb::value_error("a decoding window has no detector rows");
    if (ncols <= 0) throw qb::value_error("a decoding window has no fault columns");
    if (rows > 8192 || ncols >= 65535)
        throw qb::unsupported_error("window of " + std::to_string(rows) + " x " + std::to_string(ncols) +
                                    " exceeds what the BP kernels handle (rows <= 8192, columns < 65535)");
    std::vector<int> fillr(static_cast<size_t>(rows), 0);
    int cw = 0;
    for (int j = 0; j < ncols; ++j) cw = std::max(cw, static_cast<int>(hw.cptr[j + 1] - hw.cptr[j]));
    const int cw_alloc = cw <= 6 ? 6 : (cw <= 8 ? 8 : 16);
    if (cw > 16) throw qb::unsupported_error("column weight " + std::to_string(cw) + " > 16 is not supported by the BP kernel");
    for (int32_t r : hw.crow) fillr[r]++;
    int rs = 1;
    for (int r = 0; r < rows; ++r) rs = std::max(rs, fillr[r]);
    if (rs > 255) throw qb::unsupported_error("row weight > 255 is not supported by the BP kernel");
    rs |= 1;                                             // odd row stride: conflict-free thread-per-row sweeps
    // slot of every edge in its row and order of the column records: chosen so that the bit sweep's shared-memory
    // gathers are bank-conflict free (layout.cpp); any choice gives the same arithmetic
    qb::BpLayout layout;
    if (cw <= 6) {
        // flooding min-sum runs bp_kernel_ms2, whose row summaries are single values (bank classes as for the messages)
        qb::optimize_bp_layout(hw, rs, precision, layout, method == 0 && qb::bp_ms2_enabled() ? (precision == 32 ? 32 : 16) : 0);
    } else {
        layout.order.resize(static_cast<size_t>(ncols));
        std::iota(layout.order.begin(), layout.order.end(), 0);
        std::stable_sort(layout.order.begin(), layout.order.end(), [&](int a, int b) { return (hw.cptr[a + 1] - hw.cptr[a]) > (hw.cptr[b + 1] - hw.cptr[b]); });
        layout.slot.resize(hw.crow.size());
        std::vector<int> fill(static_cast<size_t>(rows), 0);
        for (size_t e = 0; e < hw.crow.size(); ++e) layout.slot[e] = fill[hw.crow[e]]++;
    }
    const int npad = (ncols + 31) / 32 * 32;
    std::vector<uint32_t> colE(static_cast<size_t>(cw_alloc) * npad, qb::kNoEdge);
    for (int j = 0; j < ncols; ++j) {
        int q = 0;
        for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e, ++q)
            colE[static_cast<size_t>(q) * npad + j] = (static_cast<uint32_t>(hw.crow[e]) << 8) | static_cast<uint32_t>(layout.slot[e]);
    }
    std::vector<float> llr0f(static_cast<size_t>(npad), 0.0f);
    std::vector<double> llr0d(static_cast<size_t>(npad), 0.0);
    for (int j = 0; j < ncols; ++j) {
        const double pr = hw.priors[j];
        if (!(pr > 0.0 && pr < 1.0)) throw qb::value_error("fault prior outside (0, 1)");
        llr0d[j] = std::log((1.0 - pr) / pr);
        llr0f[j] = static_cast<float>(llr0d[j]);
    }
    std::vector<uint64_t> lmask(static_cast<size_t>(std::max(hw.ncommit, 1)) * KW, 0);
    for (int j = 0; j < hw.ncommit; ++j)
        for (int64_t e = hw.lptr[j]; e < hw.lptr[j + 1]; ++e) lmask[static_cast<size_t>(j) * KW + hw.lidx[e] / 64] ^= 1ull << (hw.lidx[e] % 64);
    std::vector<int32_t> uptr(hw.uptr.begin(), hw.uptr.end());
    if (uptr.empty()) uptr.assign(static_cast<size_t>(hw.ncommit) + 1, 0);
    std::vector<uint16_t> uidx(hw.uidx.begin(), hw.uidx.end());
    std::vector<int32_t> cptr(hw.cptr.begin(), hw.cptr.end());
    std::vector<uint16_t> crow(hw.crow.begin(), hw.crow.end());
    // ---- compact form for the BP kernel's fast path (see bp.cu): columns sorted by weight, one 16-byte record each
    d.compact = 0;
    {
        std::vector<double> ptab;
        std::vector<int> pidx(static_cast<size_t>(ncols));
        for (int j = 0; j < ncols; ++j) {
            size_t k = 0;
            while (k < ptab.size() && ptab[k] != llr0d[j]) ++k;
            if (k == ptab.size()) { if (ptab.size() > 4096) break; ptab.push_back(llr0d[j]); }
            pidx[j] = static_cast<int>(k);
        }
        const bool fits = cw <= 6 && static_cast<long long>(rows) * rs + 512 + 2 * rs < 65535 && ncols < 65535 && ptab.size() < 4095 && rs <= 255;
        if (fits) {
            const uint32_t magic = static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(rs)) + 1u;
            for (uint32_t a = 0; a <= static_cast<uint32_t>(rows) * rs + 512 + static_cast<uint32_t>(rs); ++a)
                if (static_cast<uint32_t>((static_cast<uint64_t>(a) * magic) >> 32) != a / rs) throw std::runtime_error("internal: row magic is not exact");
            const std::vector<int>& order = layout.order;
            // records: 6 x u16 message address (dummy edges point at the private dummy slot rows*rs + r % NT of the thread that
            // handles record r), then w = original column | prior index << 16 | weight of the heaviest column of the record's warp << 28
            const uint32_t dummy0 = static_cast<uint32_t>(rows) * static_cast<uint32_t>(rs);
            const uint32_t nthreads = precision == 32 ? 256u : 512u;          // threads of bp_kernel_compact (bp.cu)
            const int nrec = std::max(npad, 512) + 1024;     // the kernels prefetch up to two records per thread past the last chunk
            std::vector<uint32_t> rec(static_cast<size_t>(nrec) * 4, 0);
            for (int r = 0; r < nrec; ++r) {
                uint32_t* o = &rec[static_cast<size_t>(r) * 4];
                const uint32_t dummy = dummy0 + static_cast<uint32_t>(r) % nthreads;
                uint32_t e[6] = {dummy, dummy, dummy, dummy, dummy, dummy};
                // padding record: no column, dummy edges only, and a positive prior (the extra last entry of the prior table) so
                // that its posterior is never <= 0
                uint32_t word3 = 0xFFFFu | (static_cast<uint32_t>(ptab.size()) << 16);
                if (r < ncols) {
                    const int j = order[r];
                    const int wt = static_cast<int>(hw.cptr[j + 1] - hw.cptr[j]);
                    for (int q = 0; q < wt; ++q) {
                        const uint32_t ce = colE[static_cast<size_t>(q) * npad + j];
                        e[q] = (ce >> 8) * static_cast<uint32_t>(rs) + (ce & 255u);
                    }
                    word3 = static_cast<uint32_t>(j) | (static_cast<uint32_t>(pidx[j]) << 16);
                }
                const int lead = r / 32 * 32;                 // columns are sorted by weight: the warp's first lane is its heaviest
                const int wmax = lead < ncols ? static_cast<int>(hw.cptr[order[lead] + 1] - hw.cptr[order[lead]]) : 0;
                o[0] = e[0] | (e[1] << 16);
                o[1] = e[2] | (e[3] << 16);
                o[2] = e[4] | (e[5] << 16);
                o[3] = word3 | (static_cast<uint32_t>(wmax) << 28);
            }
            std::vector<uint8_t> rlen(static_cast<size_t>(rows)), neg0(static_cast<size_t>(rows), 0);
            std::vector<double> s0d(static_cast<size_t>(rows) * 2, DBL_MAX);
            std::vector<float> s0f(static_cast<size_t>(rows) * 2, FLT_MAX);
            for (int r = 0; r < rows; ++r) rlen[r] = static_cast<uint8_t>(fillr[r]);
            for (int j = 0; j < ncols; ++j)
                for (int64_t e2 = hw.cptr[j]; e2 < hw.cptr[j + 1]; ++e2) {
                    const int r = hw.crow[e2];
                    const double ad = std::fabs(llr0d[j]);
                    const float af = std::fabs(llr0f[j]);
                    if (llr0d[j] <= 0.0) neg0[r] ^= 1;
                    double& m1 = s0d[2 * r]; double& m2 = s0d[2 * r + 1];
                    m2 = std::min(m2, std::max(m1, ad)); m1 = std::min(m1, ad);
                    float& f1 = s0f[2 * r]; float& f2 = s0f[2 * r + 1];
                    f2 = std::min(f2, std::max(f1, af)); f1 = std::min(f1, af);
                }
            std::vector<double> ptab_x(ptab);
            ptab_x.push_back(1.0);                            // prior LLR of the padding records
            std::vector<float> ptf(ptab_x.begin(), ptab_x.end());
            upload(wo.colrec, rec, ctx->stream);
            upload(wo.ptabd, ptab_x, ctx->stream);
            upload(wo.ptabf, ptf, ctx->stream);
            upload(wo.rlen, rlen, ctx->stream);
            upload(wo.neg0, neg0, ctx->stream);
            upload(wo.rsum0d, s0d, ctx->stream);
            upload(wo.rsum0f, s0f, ctx->stream);
            // chunks of 32 records by the weight of their heaviest (= first) column, for the segment loops of bp_kernel_ms2
            {
                const int nch = npad / 32;
                for (int wt = 0; wt <= 6; ++wt) d.chunk_end[wt] = 0;
                for (int ch = 0; ch < nch; ++ch) {
                    const int lead = ch * 32;
                    const int wmax = lead < ncols ? static_cast<int>(hw.cptr[order[lead] + 1] - hw.cptr[order[lead]]) : 0;
                    for (int wt = 0; wt <= wmax; ++wt) d.chunk_end[wt] = ch + 1;
                }
            }
            d.compact = 1;
            d.n_ptab = static_cast<int>(ptab_x.size());
            d.rs_magic = magic;
            d.colrec = wo.colrec.as<uint4>();
            d.ptabf = wo.ptabf.as<float>(); d.ptabd = wo.ptabd.as<double>();
            d.rlen = wo.rlen.as<uint8_t>(); d.neg0 = wo.neg0.as<uint8_t>();
            d.rsum0f = wo.rsum0f.as<float2>(); d.rsum0d = wo.rsum0d.as<double2>();
        }
    }
    upload(wo.colE, colE, ctx->stream);
    upload(wo.llr0f, llr0f, ctx->stream);
    upload(wo.llr0d, llr0d, ctx->stream);
    {
        std::vector<double> owt(static_cast<size_t>(npad), 0.0);
        for (int j = 0; j < ncols; ++j) owt[j] = std::log(1.0 / hw.priors[j]);
        upload(wo.osd_wt, owt, ctx->stream);
    }
    upload(wo.lmask, lmask, ctx->stream);
    upload(wo.uptr, uptr, ctx->stream, 2);
    upload(wo.uidx, uidx, ctx->stream, 2);
    upload(wo.cptr, cptr, ctx->stream, 2);
    upload(wo.crow, crow, ctx->stream, 2);
    {
        // CSR view of the window (ascending columns inside a row): the growth-candidate scan of the LSD kernel
        std::vector<int32_t> rptr(static_cast<size_t>(rows) + 1, 0);
        for (int32_t r : hw.crow) rptr[static_cast<size_t>(r) + 1]++;
        for (int r = 0; r < rows; ++r) rptr[r + 1] += rptr[r];
        std::vector<uint16_t> rcol(hw.crow.size());
        std::vector<int32_t> fill(rptr.begin(), rptr.end() - 1);
        for (int j = 0; j < ncols; ++j)
            for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) rcol[static_cast<size_t>(fill[hw.crow[e]]++)] = static_cast<uint16_t>(j);
        upload(wo.rptr, rptr, ctx->stream, 2);
        upload(wo.rcol, rcol, ctx->stream, 2);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    d.rows = rows; d.ncols = ncols; d.ncols_pad = npad; d.RS = rs; d.cw = cw_alloc; d.ncommit = hw.ncommit;
    d.row0 = hw.row0; d.carry_rows = hw.urows; d.KW = KW;
    d.rowsW32 = (rows + 31) / 32; d.nW32 = (ncols + 31) / 32;
    d.rank = gf2_rank(hw);
    d.full_row_rank = d.rank == rows ? 1 : 0;
    {
        double lmin = DBL_MAX;
        for (int j = 0; j < ncols; ++j) lmin = std::min(lmin, llr0d[j]);
        d.bin_scale = lmin > 1e-3 ? 10.0 / lmin : 10.0;
    }
    d.osd_wt = wo.osd_wt.as<double>();
    d.colE = wo.colE.as<uint32_t>(); d.llr0f = wo.llr0f.as<float>(); d.llr0d = wo.llr0d.as<double>(); d.lmask = wo.lmask.as<uint64_t>();
    d.uptr = wo.uptr.as<int32_t>(); d.uidx = wo.uidx.as<uint16_t>(); d.cptr = wo.cptr.as<int32_t>(); d.crow = wo.crow.as<uint16_t>();
    d.rptr = wo.rptr.as<int32_t>(); d.rcol = wo.rcol.as<uint16_t>();
}

void finish_decoder(qb_sw* sw) {
    qb_ctx* ctx = sw->ctx;
    const qb_bp_opts& o = sw->opts;
    if (o.bp_method != 0 && o.bp_method != 1) throw qb::value_error("bp_method must be 0 (minimum_sum) or 1 (product_sum)");
    if (o.schedule != 0 && o.schedule != 1) throw qb::value_error("schedule must be 0 (parallel) or 1 (serial)");
    sw->serial = o.schedule == 1;
    if (o.osd_order < 0) throw qb::value_error("osd_order must be >= 0");
    if (o.osd_method == 1 && o.osd_order > 12) throw qb::unsupported_error("osd_e beyond order 12 (4095 patterns per shot) is not supported on the GPU path");
    if (o.osd_method == 2 && o.osd_order > 32) throw qb::unsupported_error("osd_cs beyond order 32 is not supported on the GPU path");
    if (o.osd_method == 4 && o.osd_order > 12) throw qb::unsupported_error("lsd_e beyond order 12 (4095 patterns per cluster) is not supported on the GPU path");
    if (o.osd_method == 5 && o.osd_order > 32) throw qb::unsupported_error("lsd_cs beyond order 32 is not supported on the GPU path");
    sw->osd_hi = (o.osd_method == 1 || o.osd_method == 2) && o.osd_order > 0;
    if (o.ms_scaling_factor < 0) throw qb::value_error("ms_scaling_factor must be >= 0");
    if (o.precision != 0 && o.precision != 32 && o.precision != 64) throw qb::value_error("precision must be 32 or 64");
    sw->precision = o.precision == 32 ? 32 : 64;
    const int prec = sw->precision;
    if (o.osd_method > 5)
        throw qb::value_error("osd_method must be -1 (off), 0 (osd_0), 1 (osd_e), 2 (osd_cs), 3 (lsd_0), 4 (lsd_e) or 5 (lsd_cs)");
    sw->use_osd = o.osd_method >= 0 && o.osd_method <= 2;
    sw->use_lsd = o.osd_method >= 3;
    sw->lsd_hi = o.osd_method >= 4 && o.osd_order > 0;          // order 0 is LSD-0 whatever the method
    int max_npad = 0, max_rowsW = 0, max_iter = 0, big_rows = 0;
    size_t max_slab = 0, sort_slab = 0, serial_slab = 0, tall_hi_slab = 0;
    int max_sort_grid = 0, serial_grid = 0, tall_hi_grid = 0;
    bool any_big = false;
    for (auto& w : sw->wins) {
        // messages in shared memory when they fit, else in an L2-resident global slab per CTA
        w->dev.unit_alpha = o.ms_scaling_factor == 1.0 ? 1 : 0;
        if (sw->serial) {
            // one warp per shot, messages in a global slab per persistent CTA, row summaries in shared memory (bp_serial.cu)
            if (!w->serial_slab || !qb::bp_serial_slab_supported(w->dev, prec, o.bp_method))
                throw qb::unsupported_error("schedule 'serial': window of " + std::to_string(w->dev.rows) + " x " + std::to_string(w->dev.ncols) +
                                            " exceeds what the serial kernel handles (column weight <= 16, rows < 65535, row summaries in shared memory)");
            CK(qb::bp_serial_slab_configure(w->dev, prec, o.bp_method));
            w->bp_grid = 148 * qb::bp_serial_slab_ctas_per_sm(w->dev, prec, o.bp_method);
            serial_slab = std::max(serial_slab, qb::bp_serial_slab_bytes(w->dev, prec, o.bp_method));
            serial_grid = std::max(serial_grid, w->bp_grid);
            max_npad = std::max(max_npad, w->dev.ncols_pad);
            max_rowsW = std::max(max_rowsW, w->dev.rowsW32);
            max_iter = std::max(max_iter, o.max_iter > 0 ? o.max_iter : w->dev.ncols);
        }
        w->vglobal = !sw->serial && qb::bp_smem_bytes(w->dev, prec, false) > 227 * 1024;
        if (!sw->serial) {
        w->bp_smem = qb::bp_smem_bytes(w->dev, prec, w->vglobal);
        if (w->bp_smem > 227 * 1024)
            throw qb::unsupported_error("window of " + std::to_string(w->dev.rows) + " rows is too tall for the BP kernel's per-row shared-memory state");
        if (!qb::bp_supports(w->dev, o.bp_method, w->vglobal))
            throw qb::unsupported_error("bp_method 'product_sum' needs a window that fits the compact shared-memory kernel (column weight <= 6)");
        CK(qb::bp_configure(w->dev, prec, w->vglobal, o.bp_method));
        if (w->vglobal) {
            w->bp_grid = 148 * 2;
            max_slab = std::max(max_slab, qb::bp_slab_bytes(w->dev, prec));
        }
        }
        max_npad = std::max(max_npad, w->dev.ncols_pad);
        max_rowsW = std::max(max_rowsW, w->dev.rowsW32);
        max_iter = std::max(max_iter, o.max_iter > 0 ? o.max_iter : w->dev.ncols);
        if (sw->use_osd && !qb::osd_supported(w->dev, prec) && sw->osd_hi) {
            // taller than the shared-memory elimination takes, higher-order sweeps: the same elimination with T in a global slab
            if (!qb::osd_tall_hi_supported(w->dev, prec))
                throw qb::unsupported_error("window of " + std::to_string(w->dev.rows) + " x " + std::to_string(w->dev.ncols) +
                                            " exceeds what the higher-order OSD kernels handle (rows <= 2304)");
            CK(qb::osd_tall_hi_configure(w->dev, prec));
            const int per_sm = static_cast<int>((227 * 1024) / (qb::osd_elim_smem_bytes(w->dev, true) + 1024));
            w->elim_grid = 148 * std::max(1, std::min(per_sm, 4));
            w->sort_grid = 148 * 4;
            tall_hi_slab = std::max(tall_hi_slab, qb::osd_tall_hi_slab_bytes(w->dev));
            tall_hi_grid = std::max(tall_hi_grid, w->elim_grid);
        } else if (sw->use_osd && !qb::osd_supported(w->dev, prec)) {
            // taller than the shared-memory elimination takes: OSD-0 through the slab kernel (lsd.cu, osd_big_kernel)
            if (!qb::osd_big_supported(w->dev))
                throw qb::unsupported_error("window of " + std::to_string(w->dev.rows) + " x " + std::to_string(w->dev.ncols) +
                                            " exceeds what the OSD kernels handle (order 0: rows <= 3072; higher orders: rows <= 2304)");
            w->osd_big = true;
            CK(qb::osd_sort_configure(w->dev, prec));
            CK(qb::osd_big_configure(w->dev));
            const int per_sm = static_cast<int>((227 * 1024) / (qb::osd_big_smem_bytes(w->dev) + 1024));
            w->elim_grid = 148 * std::max(1, std::min(per_sm, 8));
            w->sort_grid = 148 * 4;
            any_big = true;
            big_rows = std::max(big_rows, w->dev.rows);
        } else if (sw->use_osd) {
            CK(qb::osd_configure(w->dev, prec));
            const int sort_per_sm = static_cast<int>((227 * 1024) / (qb::osd_sort_smem_bytes(w->dev, prec) + 1024));
            const int elim_per_sm = static_cast<int>((227 * 1024) / (qb::osd_elim_smem_bytes(w->dev, sw->osd_hi) + 1024));
            w->sort_grid = 148 * std::max(1, std::min(sort_per_sm, 6));
            w->elim_grid = 148 * std::max(1, std::min(elim_per_sm, 16));
            const int fast_per_sm = static_cast<int>((227 * 1024) / (qb::osd_fast_smem_bytes(w->dev) + 1024));
            w->fast_grid = 148 * std::max(1, std::min(fast_per_sm, 16));
        }
        if (sw->use_osd) {
            // windows too wide for the radix sort's shared memory keep its keys / index buffers in a global slab per CTA -- short, very
            // wide windows that still take the shared-memory elimination included
            sort_slab = std::max(sort_slab, qb::osd_sort_slab_bytes(w->dev, prec));
            max_sort_grid = std::max(max_sort_grid, w->sort_grid);
        }
        if (sw->use_lsd) {
            if (!qb::lsd_supported(w->dev))
                throw qb::unsupported_error("window of " + std::to_string(w->dev.rows) + " x " + std::to_string(w->dev.ncols) +
                                            " exceeds what the LSD kernel handles (rows <= 3072, columns < 65535)");
            CK(qb::lsd_configure(w->dev, prec, sw->lsd_hi));
            const int per_sm = static_cast<int>((227 * 1024) / (qb::lsd_smem_bytes(w->dev) + 1024));
            w->lsd_grid = 148 * std::max(1, std::min(per_sm, 16));
        }
    }
    if (max_slab) sw->vscratch.ensure(max_slab * 148 * 2 + 16);
    sw->serial_slab = serial_slab;
    sw->serial_grid = serial_grid;
    if (sw->use_lsd) {
        // one slab per persistent warp: bit owners (0xFFFF = none between shots), column-order links, operation vectors
        int max_rows = 0;
        for (auto& w : sw->wins) max_rows = std::max(max_rows, w->dev.rows);
        sw->lsd_cols = max_npad;
        sw->lsd_slab = qb::lsd_slab_bytes(max_npad, max_rows);
        for (auto& w : sw->wins) sw->lsd_grid = std::max(sw->lsd_grid, w->lsd_grid);
    }
    if (any_big) {
        // the slab kernel's row operations (one slab per persistent warp) and the wide sort's key / index buffers (one per CTA)
        sw->use_slab = true;
        sw->lsd_slab = qb::osd_big_slab_bytes(big_rows);
        for (auto& w : sw->wins) if (w->osd_big) sw->lsd_grid = std::max(sw->lsd_grid, w->elim_grid);
    }
    if (tall_hi_slab) {
        // higher-order OSD on tall windows: one T slab per persistent warp (shares the LSD / slab-OSD scratch buffer)
        sw->use_slab = true;
        sw->lsd_slab = std::max(sw->lsd_slab, tall_hi_slab);
        sw->lsd_grid = std::max(sw->lsd_grid, tall_hi_grid);
    }
    if (sort_slab) sw->sort_scratch.ensure(sort_slab * static_cast<size_t>(std::max(max_sort_grid, 1)) + 16);
    if (o.max_iter == 0 && sw->wins.size() > 1) {
        // ldpc's "0 => number of columns" differs per window; the kernel takes one value
        for (auto& w : sw->wins)
            if (w->dev.ncols != sw->wins[0]->dev.ncols) throw qb::unsupported_error("max_iter = 0 with windows of different widths; pass max_iter explicitly");
    }
    sw->max_iter = max_iter;
    std::vector<double> alpha(static_cast<size_t>(max_iter) + 1, 1.0);
    for (int it = 1; it <= max_iter; ++it) alpha[it] = o.ms_scaling_factor == 0.0 ? 1.0 - std::pow(2.0, -1.0 * it) : o.ms_scaling_factor;
    upload(sw->alpha, alpha, ctx->stream);
    // batches start on 64-shot word boundaries (the sampler numbers shots by word): capacity is rounded down to a multiple of 64
    static const int env_cap = [] { const char* e = getenv("QB_CAPACITY"); return e ? atoi(e) : 0; }();      // tuning knob
    const int want_cap = o.capacity > 0 ? o.capacity : (env_cap > 0 ? env_cap : 262144);
    sw->cap = std::max(64, want_cap / 64 * 64);
    {
        // the posterior scratch is capacity x widest window: keep it under 8 GiB for very wide windows
        const size_t per_shot = static_cast<size_t>(max_npad) * (prec / 8);
        const size_t fit = (size_t(8) << 30) / std::max<size_t>(per_shot, 1);
        if (static_cast<size_t>(sw->cap) > fit) sw->cap = static_cast<int>(std::max<size_t>(64, fit / 64 * 64));
    }
    // LSD: a few shots grow clusters of hundreds of bits and keep one warp busy for milliseconds after the rest of the launch has
    // drained; with sub-batches on side streams the BP kernel of another sub-batch fills the machine meanwhile
    sw->lanes = o.lanes > 0 ? std::min(o.lanes, static_cast<int32_t>(qb_ctx::kMaxLanes)) : (sw->use_lsd ? static_cast<int>(qb_ctx::kMaxLanes) : 1);
    if (max_slab || any_big || sort_slab || tall_hi_slab) sw->lanes = 1;           // the global message / sort slabs are indexed by CTA, not by sub-batch
    sw->DW = std::max(1, (sw->plan.D + 63) / 64);
    sw->KW = std::max(1, (sw->plan.K + 63) / 64);
    sw->carryW = (sw->plan.m + 31) / 32 + 1;
    sw->synW = max_rowsW;
    sw->llr_stride = static_cast<size_t>(max_npad);
    CK(cudaStreamSynchronize(ctx->stream));
}

void ensure_batch(qb_sw* sw, int n) {
    const size_t N = static_cast<size_t>(n);
    sw->carry.ensure(N * sw->carryW * 4 + 16);
    sw->acc.ensure(N * sw->KW * 8 + 16);
    sw->llr.ensure(N * sw->llr_stride * (sw->precision / 8) + 16);
    sw->syn.ensure(N * sw->synW * 4 + 16);
    sw->fail_list.ensure(N * 4 + 16);
    if (sw->use_osd) {
        sw->ovf_list.ensure(N * 4 + 16);
        sw->sel_key.ensure(N * qb::kOsdSelCap * (sw->precision / 8) + 16);
        sw->sel_idx.ensure(N * qb::kOsdSelCap * 2 + 16);
        sw->sel_cnt.ensure(N * 4 + 16);
    }
    if (sw->use_lsd || sw->use_slab) {
        // one slab per persistent warp and sub-batch lane: bit owners (0xFFFF = none between shots), column-order links, operation vectors
        const int per_lane = std::min(sw->lsd_grid, n);
        if (per_lane > sw->lsd_slabs_per_lane) {
            const size_t bytes = sw->lsd_slab * static_cast<size_t>(per_lane) * static_cast<size_t>(std::max(sw->lanes, 1)) + 16;
            sw->lsd_scratch.ensure(bytes);
            CK(cudaMemsetAsync(sw->lsd_scratch.p, 0xFF, bytes, sw->ctx->stream));
            sw->lsd_slabs_per_lane = per_lane;
        }
    }
    if (sw->serial_slab) sw->vscratch.ensure(sw->serial_slab * static_cast<size_t>(sw->serial_grid) * static_cast<size_t>(std::max(sw->lanes, 1)) + 64);
    // counters and statistics: one set per (batch slot, sub-batch lane, window); batch slots let the fused run queue several
    // batches before it reads anything back
    const size_t nw = sw->wins.size() * qb_ctx::kMaxLanes * kBatchSlots;
    sw->counters.ensure(nw * kCounterSlots * sizeof(int) + 16);
    sw->stats.ensure(nw * kStatSlots * sizeof(unsigned long long) + 16);
}

// decode n (<= cap) shots whose packed detector rows are on the device; leaves acc[n][KW] on the device.
// The batch is cut into `lanes` contiguous sub-batches, each walking the windows on its own stream (the windows of one
// shot are sequential -- carry dependency, sliding_window.py:169,174 -- but shots are independent).
void decode_batch(qb_sw* sw, const uint64_t* d_det_rows, int n, bool want_ehat, bool want_llr, qb_stats* stats, int batch_slot = 0) {
    qb_ctx* ctx = sw->ctx;
    cudaStream_t st = ctx->stream;
    ensure_batch(sw, n);
    if (want_llr && sw->use_osd) sw->order.ensure(static_cast<size_t>(n) * sw->llr_stride * 2 + 16);
    const size_t nw = sw->wins.size();
    int lanes = sw->lanes;
    if (want_ehat || n < 4096) lanes = 1;
    sw->lanes_used = lanes;
    CK(cudaMemsetAsync(sw->acc.p, 0, static_cast<size_t>(n) * sw->KW * 8, st));
    CK(cudaMemsetAsync(sw->carry.p, 0, static_cast<size_t>(n) * sw->carryW * 4, st));
    const size_t slot0 = static_cast<size_t>(batch_slot) * nw * qb_ctx::kMaxLanes;          // first (lane, window) set of this batch slot
    CK(cudaMemsetAsync(sw->counters.as<int>() + kCounterSlots * slot0, 0, nw * lanes * kCounterSlots * sizeof(int), st));
    CK(cudaMemsetAsync(sw->stats.as<unsigned long long>() + kStatSlots * slot0, 0, nw * lanes * kStatSlots * sizeof(unsigned long long), st));
    if (lanes > 1) {
        for (int l = 1; l < lanes; ++l) ctx->lane_stream(l);
        CK(cudaEventRecord(ctx->ev_fork, st));
        for (int l = 1; l < lanes; ++l) CK(cudaStreamWaitEvent(ctx->lane_stream(l), ctx->ev_fork, 0));
    }
    qb::BpParams bp{};
    bp.max_iter = sw->max_iter;
    bp.method = sw->opts.bp_method;
    bp.alpha = sw->alpha.as<double>();
    const size_t esz = static_cast<size_t>(sw->precision / 8);
    for (size_t k = 0; k < nw; ++k) {
        WinOwned& w = *sw->wins[k];
        for (int l = 0; l < lanes; ++l) {
            const size_t s0 = static_cast<size_t>(n) * l / lanes, s1 = static_cast<size_t>(n) * (l + 1) / lanes;
            const int nl = static_cast<int>(s1 - s0);
            if (nl == 0) continue;
            cudaStream_t ls = ctx->lane_stream(l);
            qb::BatchDev b{};
            b.n_shots = nl;
            b.det_stride32 = 2 * sw->DW;
            b.det32 = reinterpret_cast<const uint32_t*>(d_det_rows) + s0 * b.det_stride32;
            b.in_carry_rows = k == 0 ? 0 : sw->wins[k - 1]->dev.carry_rows;
            b.carry_stride32 = sw->carryW;
            b.carry = sw->carry.as<uint32_t>() + s0 * sw->carryW;
            b.acc = sw->acc.as<uint64_t>() + s0 * sw->KW;
            b.llr_stride = sw->llr_stride;
            b.llr_esize = sw->precision / 8;
            b.order_alt = (want_llr && sw->use_osd) ? sw->order.as<uint16_t>() + s0 * sw->llr_stride : nullptr;   // keep the posteriors intact for the caller
            b.llr_buf = static_cast<unsigned char*>(sw->llr.p) + s0 * sw->llr_stride * esz;
            b.vscratch = sw->serial_slab ? static_cast<void*>(static_cast<unsigned char*>(sw->vscratch.p) + static_cast<size_t>(l) * sw->serial_slab * sw->serial_grid)
                                         : sw->vscratch.p;
            b.syn_stride32 = sw->synW;
            b.syn_buf = sw->syn.as<uint32_t>() + s0 * sw->synW;
            b.fail_list = sw->fail_list.as<int>() + s0;          // shot indices local to the sub-batch
            const size_t slot = slot0 + static_cast<size_t>(l) * nw + k;
            int* ctr = sw->counters.as<int>() + kCounterSlots * slot;
            b.fail_count = ctr;
            b.fast_next = ctr + 1;
            b.ovf_count = ctr + 2;
            b.sort_next = ctr + 3;
            b.osd_next = ctr + 4;
            b.bp_next = ctr + 5;
            b.ovf_list = sw->use_osd ? sw->ovf_list.as<int>() + s0 : nullptr;
            b.osd_method = sw->opts.osd_method;
            b.osd_order = sw->opts.osd_order;
            b.lsd_scratch = sw->lsd_scratch.p ? static_cast<unsigned char*>(sw->lsd_scratch.p) + static_cast<size_t>(l) * sw->lsd_slab * sw->lsd_slabs_per_lane : nullptr;
            b.lsd_slab = sw->lsd_slab;
            b.lsd_cols = sw->lsd_cols;
            b.lsd_method = sw->lsd_hi ? (sw->opts.osd_method == 4 ? 1 : 2) : 0;
            b.lsd_order = sw->lsd_hi ? sw->opts.osd_order : 0;
            b.sort_scratch = sw->sort_scratch.p;
            if (sw->use_osd && !sw->osd_hi && !w.osd_big) {
                b.sel_key = static_cast<unsigned char*>(sw->sel_key.p) + s0 * qb::kOsdSelCap * esz;
                b.sel_idx = sw->sel_idx.as<uint16_t>() + s0 * qb::kOsdSelCap;
                b.sel_cnt = sw->sel_cnt.as<int>() + s0;
            }
            b.stats = sw->stats.as<unsigned long long>() + kStatSlots * slot;
            b.ehat_out = want_ehat ? sw->ehat.as<uint32_t>() : nullptr;
            b.ehat_stride32 = w.dev.nW32;
            b.iters_out = want_ehat ? sw->iters.as<int32_t>() : nullptr;
            b.conv_out = want_ehat ? sw->conv.as<uint8_t>() : nullptr;
            b.write_llr_always = want_llr ? 1 : 0;
            b.commit_unconverged = (!sw->use_osd && !sw->use_lsd) ? 1 : 0;
            if (sw->opts.profile) sw->t_bp.begin(ls);
            if (sw->serial) CK(qb::launch_bp_serial_slab(w.dev, b, bp, sw->precision, std::min(w.bp_grid, nl), ls));
            else {
                const int pg = qb::bp_persistent_grid(w.dev, sw->precision, w.vglobal, bp.method);
                if (!pg) b.bp_next = nullptr;
                CK(qb::launch_bp(w.dev, b, bp, sw->precision, w.vglobal, w.vglobal ? std::min(w.bp_grid, nl) : (pg ? std::min(pg, nl) : nl), ls));
            }
            if (sw->opts.profile) sw->t_bp.end(ls);
            if (stats) stats->bp_launches++;
            if (sw->use_lsd) {
                if (sw->opts.profile) sw->t_osd.begin(ls);
                CK(qb::launch_lsd(w.dev, b, sw->precision, std::min(w.lsd_grid, nl), ls));
                if (sw->opts.profile) sw->t_osd.end(ls);
                if (stats) stats->osd_launches += 1;
            }
            if (sw->use_osd && w.osd_big) {
                if (sw->opts.profile) sw->t_osd.begin(ls);
                CK(qb::launch_osd_sort(w.dev, b, sw->precision, std::min(w.sort_grid, nl), ls));
                CK(qb::launch_osd_big(w.dev, b, std::min(w.elim_grid, nl), ls));
                if (sw->opts.profile) sw->t_osd.end(ls);
                if (stats) stats->osd_launches += 2;
                continue;
            }
            if (sw->use_osd) {
                if (sw->opts.profile) sw->t_osd.begin(ls);
                if (sw->osd_hi) {
                    // higher-order OSD needs the complete elimination over the complete column order
                    CK(qb::launch_osd_sort(w.dev, b, sw->precision, std::min(w.sort_grid, nl), ls));
                    CK(qb::launch_osd_elim(w.dev, b, true, std::min(w.elim_grid, nl), ls));
                    if (sw->opts.profile) sw->t_osd.end(ls);
                    if (stats) stats->osd_launches += 2;
                    continue;
                }
                // fast path: per-warp selection of the least reliable columns + elimination; the (rare) shots it cannot finish
                // go through the full sort + elimination, which read the overflow list instead of the fail list
                CK(qb::launch_osd_fast(w.dev, b, sw->precision, std::min(w.fast_grid, nl), ls));
                qb::BatchDev bo = b;
                bo.fail_list = b.ovf_list;
                bo.fail_count = b.ovf_count;
                CK(qb::launch_osd_sort(w.dev, bo, sw->precision, std::min(w.sort_grid, std::max(1, nl / 64)), ls));
                CK(qb::launch_osd_elim(w.dev, bo, false, std::min(w.elim_grid, std::max(1, nl / 64)), ls));
                if (sw->opts.profile) sw->t_osd.end(ls);
                if (stats) stats->osd_launches += 3;
            }
        }
    }
    for (int l = 1; l < lanes; ++l) {
        CK(cudaEventRecord(ctx->ev_join[l - 1], ctx->lane_stream(l)));
        CK(cudaStreamWaitEvent(st, ctx->ev_join[l - 1], 0));
    }
}

void collect_stats(qb_sw* sw, int n, qb_stats* stats, int batch_slot = 0) {        // stream must be synchronised
    if (!stats) return;
    const size_t nw = sw->wins.size();
    const int lanes = sw->lanes_used;
    std::vector<unsigned long long> h(nw * lanes * kStatSlots);
    CK(cudaMemcpy(h.data(), sw->stats.as<unsigned long long>() + kStatSlots * static_cast<size_t>(batch_slot) * nw * qb_ctx::kMaxLanes,
                  h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    stats->shots += n;
    stats->windows += static_cast<int64_t>(nw) * n;
    for (int l = 0; l < lanes; ++l)
        for (size_t k = 0; k < nw; ++k) {
            const unsigned long long* hk = &h[(static_cast<size_t>(l) * nw + k) * kStatSlots];
            stats->bp_converged += static_cast<int64_t>(hk[0]);
            stats->bp_iterations += static_cast<int64_t>(hk[1]);
            stats->osd_calls += static_cast<int64_t>(hk[2]);
            stats->osd_columns += static_cast<int64_t>(hk[3]);
            stats->osd_pivots += static_cast<int64_t>(hk[4]);
            stats->osd_max_columns = std::max(stats->osd_max_columns, static_cast<int64_t>(hk[5]));
            stats->osd_overflows += static_cast<int64_t>(hk[6]);
            const qb::WinDev& d = sw->wins[k]->dev;
            const double nnz = static_cast<double>(sw->plan.windows[k].crow.size());
            const size_t s0 = static_cast<size_t>(n) * l / lanes, s1 = static_cast<size_t>(n) * (l + 1) / lanes;
            const double io = 8.0 * ((d.rows + 63) / 64) + 8.0 * sw->KW + 8.0 * ((d.carry_rows + 63) / 64);
            stats->bp_alg_bytes += static_cast<double>(hk[1]) * 4.0 * nnz * (sw->precision / 8) + io * static_cast<double>(s1 - s0);
            stats->bp_edge_iters += static_cast<double>(hk[1]) * nnz;
            stats->osd_alg_bytes += static_cast<double>(hk[2]) * 2.0 * d.rows * 8.0 * ((d.ncols + 63) / 64);
        }
    if (sw->opts.profile) {
        stats->bp_ms += sw->t_bp.collect();
        stats->osd_ms += sw->t_osd.collect();
    }
}

void parse_opts(const qb_bp_opts* in, qb_bp_opts& out) {
    if (in) out = *in;
    else {
        out = qb_bp_opts{};
        out.max_iter = 10;
        out.ms_scaling_factor = 1.0;
    }
}

}  // namespace

template <typename V>
static void export_csr(const V& lists, int64_t* ptr, int32_t* idx) {
    int64_t pos = 0;
    for (size_t i = 0; i < lists.size(); ++i) {
        if (ptr) ptr[i] = pos;
        if (idx) for (size_t j = 0; j < lists[i].size(); ++j) idx[pos + static_cast<int64_t>(j)] = lists[i][j];
        pos += static_cast<int64_t>(lists[i].size());
    }
    if (ptr) ptr[lists.size()] = pos;
}


// ================================================================================================ C ABI
extern "C" {

const char* qb_last_error(void) { return g_err.c_str(); }
int qb_version(void) { return 100; }

#ifndef QB_SRC_HASH
#define QB_SRC_HASH "unknown"
#endif
#define QB_STR2(x) #x
#define QB_STR(x) QB_STR2(x)
const char* qb_build_info(void) {
    return "quits_b200 abi=100 arch=sm_100a nvcc=" QB_STR(__CUDACC_VER_MAJOR__) "." QB_STR(__CUDACC_VER_MINOR__) "." QB_STR(__CUDACC_VER_BUILD__)
           " fmad=off lineinfo=on src=" QB_SRC_HASH;
}

int qb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int qb_ctx_create(int device, qb_ctx** out) {
    return guard([&] {
        if (!out) throw arg_error("out is NULL");
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            throw cuda_error("no CUDA device is available: the quits_b200 engine has no CPU fallback");
        }
        if (device < 0 || device >= n) throw arg_error("device index out of range");
        CK(cudaSetDevice(device));
        std::unique_ptr<qb_ctx> ctx(new qb_ctx());
        ctx->device = device;
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        *out = ctx.release();
    });
}

void qb_ctx_destroy(qb_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    for (int i = 0; i < qb_ctx::kMaxLanes - 1; ++i) {
        if (ctx->aux[i]) { cudaStreamSynchronize(ctx->aux[i]); cudaStreamDestroy(ctx->aux[i]); }
        if (ctx->ev_join[i]) cudaEventDestroy(ctx->ev_join[i]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->copy) { cudaStreamSynchronize(ctx->copy); cudaStreamDestroy(ctx->copy); }
    for (int i = 0; i < 2; ++i) {
        if (ctx->ev_ready[i]) cudaEventDestroy(ctx->ev_ready[i]);
        if (ctx->ev_done[i]) cudaEventDestroy(ctx->ev_done[i]);
    }
    delete ctx;
}

int qb_ctx_synchronize(qb_ctx* ctx) {
    return guard([&] { use_device(ctx); CK(cudaStreamSynchronize(ctx->stream)); });
}

int qb_host_alloc(size_t bytes, void** out) {
    return guard([&] {
        if (!out) throw arg_error("out is NULL");
        *out = nullptr;
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); throw cuda_error("no CUDA device is available"); }
        CK(cudaHostAlloc(out, std::max<size_t>(bytes, 16), cudaHostAllocPortable));
    });
}

void qb_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// ------------------------------------------------------------------------------------------------ circuit
int qb_circuit_parse(const char* text, size_t len, qb_circuit** out) {
    return guard([&] {
        if (!text || !out) throw arg_error("NULL argument");
        std::unique_ptr<qb_circuit> c(new qb_circuit());
        qb::parse_flatten(text, len, c->fc);
        qb::build_tape(c->fc, c->tape);
        c->noise_tape_of_flat.assign(c->fc.ops.size(), -1);
        for (size_t t = 0; t < c->tape.ops.size(); ++t) {
            const int k = c->tape.ops[t].kind;
            if (k >= qb::OP_XERR && k <= qb::OP_DEP2) c->noise_tape_of_flat[c->tape.ops[t].flat] = static_cast<int32_t>(t);
        }
        *out = c.release();
    });
}

void qb_circuit_free(qb_circuit* c) { delete c; }

int qb_circuit_get_info(const qb_circuit* c, qb_circuit_info* info) {
    return guard([&] {
        if (!c || !info) throw arg_error("NULL argument");
        info->n_qubits = c->fc.n_qubits;
        info->n_measurements = c->fc.n_meas;
        info->n_detectors = c->fc.n_det;
        info->n_observables = c->fc.n_obs;
        info->n_flat_ops = static_cast<int64_t>(c->fc.ops.size());
        info->n_tape_ops = static_cast<int64_t>(c->tape.ops.size());
        info->n_noise_sites = c->fc.n_sites;
        info->ring = c->tape.ring;
    });
}

int qb_circuit_flat(const qb_circuit* c, int32_t* kind, double* arg, int64_t* tstart, int32_t* targets, int64_t* n_targets_out) {
    return guard([&] {
        if (!c) throw arg_error("NULL argument");
        int64_t pos = 0;
        for (size_t i = 0; i < c->fc.ops.size(); ++i) {
            const qb::FlatOp& op = c->fc.ops[i];
            if (kind) kind[i] = op.kind;
            if (arg) arg[i] = op.arg;
            if (tstart) tstart[i] = pos;
            if (targets) for (size_t j = 0; j < op.targets.size(); ++j) targets[pos + static_cast<int64_t>(j)] = op.targets[j];
            pos += static_cast<int64_t>(op.targets.size());
        }
        if (tstart) tstart[c->fc.ops.size()] = pos;
        if (n_targets_out) *n_targets_out = pos;
    });
}

// ------------------------------------------------------------------------------------------------ sampler
static void sample_impl(qb_ctx* ctx, qb_circuit* c, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint8_t* det, uint8_t* obs,
                        uint64_t* det_rows, uint64_t* obs_rows) {
    if (!ctx || !c) throw arg_error("NULL argument");
    if (shot0 & 63) throw arg_error("shot0 must be a multiple of 64");
    use_device(ctx);
    circuit_to_device(ctx, c);
    qb::FrameArgs a = frame_args(c, seed);
    if (qb::frame_smem_per_warp(a) > 200 * 1024) throw qb::unsupported_error("circuit too large for the shared-memory frame kernel");
    const int D = c->fc.n_det, K = c->fc.n_obs;
    const bool to_host = det || obs || det_rows || obs_rows;
    // results that go back to the host are produced in chunks of 65536 shots into two buffer sets: the device-to-host copies of a
    // chunk run on the copy stream while the frame kernel of the next chunk runs
    const uint64_t chunk = to_host ? (1ull << 16) : kSampleChunk;
    cudaStream_t cs = to_host ? ctx->copy_stream() : nullptr;
    int k = 0;
    for (uint64_t done = 0; done < n_shots; done += chunk, ++k) {
        const uint64_t n = std::min(chunk, n_shots - done);
        const uint64_t nwords = (n + 63) / 64;
        const int bsel = to_host ? (k & 1) : 0;
        DevBuf& rows_d = bsel ? ctx->det_rows_alt : ctx->det_rows;
        DevBuf& rows_o = bsel ? ctx->obs_rows_alt : ctx->obs_rows;
        DevBuf& bytes_d = bsel ? ctx->det_bytes_alt : ctx->det_bytes;
        DevBuf& bytes_o = bsel ? ctx->obs_bytes_alt : ctx->obs_bytes;
        if (to_host && k >= 2) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_done[bsel], 0));      // the copies of chunk k-2 have left this set
        rows_d.ensure(std::min<uint64_t>(chunk, n_shots) / 64 * 64 * a.DW * 8 + 64 * a.DW * 8 + 16);
        rows_o.ensure(std::min<uint64_t>(chunk, n_shots) / 64 * 64 * a.KW * 8 + 64 * a.KW * 8 + 16);
        a.word0 = (shot0 + done) / 64;
        a.n_words = nwords;
        a.det_rows = rows_d.as<uint64_t>();
        a.obs_rows = rows_o.as<uint64_t>();
        ctx->det_words.ensure(qb::frame_scratch_bytes(a)); a.det_words = ctx->det_words.as<uint64_t>(); CK(qb::launch_frame(a, ctx->stream));
        if (det || obs) {
            bytes_d.ensure(std::min<uint64_t>(chunk, n_shots) * std::max(D, 1) + 16);
            bytes_o.ensure(std::min<uint64_t>(chunk, n_shots) * std::max(K, 1) + 16);
            CK(qb::launch_unpack_bits(a.det_rows, a.DW, D, n, bytes_d.as<uint8_t>(), ctx->stream));
            CK(qb::launch_unpack_bits(a.obs_rows, a.KW, K, n, bytes_o.as<uint8_t>(), ctx->stream));
        }
        if (!to_host) continue;
        CK(cudaEventRecord(ctx->ev_ready[bsel], ctx->stream));
        CK(cudaStreamWaitEvent(cs, ctx->ev_ready[bsel], 0));
        if (det && D) CK(cudaMemcpyAsync(det + done * D, bytes_d.p, n * D, cudaMemcpyDeviceToHost, cs));
        if (obs && K) CK(cudaMemcpyAsync(obs + done * K, bytes_o.p, n * K, cudaMemcpyDeviceToHost, cs));
        if (det_rows) CK(cudaMemcpyAsync(det_rows + done * a.DW, a.det_rows, n * a.DW * 8, cudaMemcpyDeviceToHost, cs));
        if (obs_rows) CK(cudaMemcpyAsync(obs_rows + done * a.KW, a.obs_rows, n * a.KW * 8, cudaMemcpyDeviceToHost, cs));
        CK(cudaEventRecord(ctx->ev_done[bsel], cs));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    if (cs) CK(cudaStreamSynchronize(cs));
}

int qb_sample(qb_ctx* ctx, qb_circuit* c, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint8_t* det, uint8_t* obs) {
    return guard([&] { sample_impl(ctx, c, seed, shot0, n_shots, det, obs, nullptr, nullptr); });
}

int qb_sample_packed(qb_ctx* ctx, qb_circuit* c, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint64_t* det_rows, uint64_t* obs_rows) {
    return guard([&] { sample_impl(ctx, c, seed, shot0, n_shots, nullptr, nullptr, det_rows, obs_rows); });
}

int qb_sample_faults(qb_ctx* ctx, qb_circuit* c, int64_t n_faults, const int32_t* op, const int32_t* tgt, const int32_t* code,
                     const int64_t* shot, uint64_t n_shots, uint8_t* det, uint8_t* obs) {
    return guard([&] {
        if (!ctx || !c) throw arg_error("NULL argument");
        if (n_faults < 0 || (n_faults > 0 && (!op || !tgt || !code || !shot))) throw arg_error("bad fault list");
        use_device(ctx);
        circuit_to_device(ctx, c);
        const size_t T = c->tape.ops.size();
        std::vector<int32_t> tape_of(static_cast<size_t>(n_faults));
        for (int64_t f = 0; f < n_faults; ++f) {
            if (op[f] < 0 || static_cast<size_t>(op[f]) >= c->fc.ops.size() || c->noise_tape_of_flat[op[f]] < 0)
                throw arg_error("fault " + std::to_string(f) + " does not address a noise instruction");
            const qb::FlatOp& fo = c->fc.ops[op[f]];
            const int ns = fo.kind == qb::OP_DEP2 ? static_cast<int>(fo.targets.size() / 2) : static_cast<int>(fo.targets.size());
            if (tgt[f] < 0 || tgt[f] >= ns) throw arg_error("fault " + std::to_string(f) + ": target index out of range");
            if (shot[f] < 0 || static_cast<uint64_t>(shot[f]) >= n_shots) throw arg_error("fault " + std::to_string(f) + ": shot out of range");
            if (code[f] < 1 || code[f] > (fo.kind == qb::OP_DEP2 ? 15 : 3)) throw arg_error("fault " + std::to_string(f) + ": bad Pauli code");
            tape_of[f] = c->noise_tape_of_flat[op[f]];
        }
        std::vector<int64_t> order(static_cast<size_t>(n_faults));
        std::iota(order.begin(), order.end(), 0);
        std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return tape_of[a] < tape_of[b]; });
        std::vector<int32_t> start(T + 1, 0), stgt(static_cast<size_t>(n_faults)), scode(static_cast<size_t>(n_faults));
        std::vector<int64_t> sshot(static_cast<size_t>(n_faults));
        for (int64_t i = 0; i < n_faults; ++i) {
            const int64_t f = order[i];
            start[tape_of[f] + 1]++;
            stgt[i] = tgt[f]; scode[i] = code[f]; sshot[i] = shot[f];
        }
        for (size_t t = 0; t < T; ++t) start[t + 1] += start[t];
        upload(ctx->inj_start, start, ctx->stream);
        upload(ctx->inj_tgt, stgt, ctx->stream);
        upload(ctx->inj_code, scode, ctx->stream);
        upload(ctx->inj_shot, sshot, ctx->stream);
        qb::FrameArgs a = frame_args(c, 0);
        const int D = c->fc.n_det, K = c->fc.n_obs;
        const uint64_t nwords = (n_shots + 63) / 64;
        ctx->det_rows.ensure(nwords * 64 * a.DW * 8 + 16);
        ctx->obs_rows.ensure(nwords * 64 * a.KW * 8 + 16);
        a.word0 = 0; a.n_words = nwords;
        a.det_rows = ctx->det_rows.as<uint64_t>();
        a.obs_rows = ctx->obs_rows.as<uint64_t>();
        a.inject = 1;
        a.inj_start = ctx->inj_start.as<int32_t>();
        a.inj_tgt = ctx->inj_tgt.as<int32_t>();
        a.inj_code = ctx->inj_code.as<int32_t>();
        a.inj_shot = ctx->inj_shot.as<int64_t>();
        ctx->det_words.ensure(qb::frame_scratch_bytes(a)); a.det_words = ctx->det_words.as<uint64_t>(); CK(qb::launch_frame(a, ctx->stream));
        ctx->det_bytes.ensure(n_shots * std::max(D, 1) + 16);
        ctx->obs_bytes.ensure(n_shots * std::max(K, 1) + 16);
        CK(qb::launch_unpack_bits(a.det_rows, a.DW, D, n_shots, ctx->det_bytes.as<uint8_t>(), ctx->stream));
        CK(qb::launch_unpack_bits(a.obs_rows, a.KW, K, n_shots, ctx->obs_bytes.as<uint8_t>(), ctx->stream));
        if (det && D) CK(cudaMemcpyAsync(det, ctx->det_bytes.p, n_shots * D, cudaMemcpyDeviceToHost, ctx->stream));
        if (obs && K) CK(cudaMemcpyAsync(obs, ctx->obs_bytes.p, n_shots * K, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    });
}

// ------------------------------------------------------------------------------------------------ DEM
int qb_dem_from_circuit(const qb_circuit* c, qb_dem** out) {
    return guard([&] {
        if (!c || !out) throw arg_error("NULL argument");
        std::unique_ptr<qb_dem> d(new qb_dem());
        qb::analyze(c->fc, d->dem);
        qb::dem_to_matrix(d->dem, d->cm);
        *out = d.release();
    });
}

int qb_dem_from_errors(int32_t n_detectors, int32_t n_observables, int64_t n_errors, const double* probs, const int64_t* det_ptr,
                       const int32_t* det_idx, const int64_t* obs_ptr, const int32_t* obs_idx, qb_dem** out) {
    return guard([&] {
        if (!out || n_errors < 0 || (n_errors && (!probs || !det_ptr || !obs_ptr))) throw arg_error("NULL argument");
        std::unique_ptr<qb_dem> d(new qb_dem());
        d->dem.n_det = n_detectors;
        d->dem.n_obs = n_observables;
        for (int64_t e = 0; e < n_errors; ++e) {
            std::vector<int32_t> dd(det_idx + det_ptr[e], det_idx + det_ptr[e + 1]), oo(obs_idx + obs_ptr[e], obs_idx + obs_ptr[e + 1]);
            for (int32_t v : dd) if (v < 0 || v >= n_detectors) throw arg_error("detector id out of range");
            for (int32_t v : oo) if (v < 0 || v >= n_observables) throw arg_error("observable id out of range");
            d->dem.probs.push_back(probs[e]);
            d->dem.dets.push_back(std::move(dd));
            d->dem.obs.push_back(std::move(oo));
            d->dem.rep_op.push_back(-1); d->dem.rep_tgt.push_back(-1); d->dem.rep_code.push_back(0);
        }
        qb::dem_to_matrix(d->dem, d->cm);
        *out = d.release();
    });
}

void qb_dem_free(qb_dem* d) { delete d; }

int qb_dem_sizes(const qb_dem* d, int64_t sizes[9]) {
    return guard([&] {
        if (!d || !sizes) throw arg_error("NULL argument");
        int64_t nd = 0, no = 0, nh = 0, nl = 0;
        for (auto& v : d->dem.dets) nd += static_cast<int64_t>(v.size());
        for (auto& v : d->dem.obs) no += static_cast<int64_t>(v.size());
        for (auto& v : d->cm.col_dets) nh += static_cast<int64_t>(v.size());
        for (auto& v : d->cm.col_obs) nl += static_cast<int64_t>(v.size());
        sizes[0] = d->dem.n_det; sizes[1] = d->dem.n_obs; sizes[2] = static_cast<int64_t>(d->dem.probs.size());
        sizes[3] = nd; sizes[4] = no; sizes[5] = static_cast<int64_t>(d->cm.priors.size()); sizes[6] = nh; sizes[7] = nl;
        sizes[8] = d->cm.n_detless;
    });
}

int qb_dem_errors(const qb_dem* d, double* probs, int64_t* det_ptr, int32_t* det_idx, int64_t* obs_ptr, int32_t* obs_idx,
                  int32_t* rep_op, int32_t* rep_tgt, int32_t* rep_code) {
    return guard([&] {
        if (!d) throw arg_error("NULL argument");
        const size_t n = d->dem.probs.size();
        if (probs) memcpy(probs, d->dem.probs.data(), n * 8);
        export_csr(d->dem.dets, det_ptr, det_idx);
        export_csr(d->dem.obs, obs_ptr, obs_idx);
        if (rep_op) memcpy(rep_op, d->dem.rep_op.data(), n * 4);
        if (rep_tgt) memcpy(rep_tgt, d->dem.rep_tgt.data(), n * 4);
        if (rep_code) memcpy(rep_code, d->dem.rep_code.data(), n * 4);
    });
}

int qb_dem_matrix(const qb_dem* d, int64_t* h_ptr, int32_t* h_idx, int64_t* l_ptr, int32_t* l_idx, double* priors) {
    return guard([&] {
        if (!d) throw arg_error("NULL argument");
        export_csr(d->cm.col_dets, h_ptr, h_idx);
        export_csr(d->cm.col_obs, l_ptr, l_idx);
        if (priors) memcpy(priors, d->cm.priors.data(), d->cm.priors.size() * 8);
    });
}

// ------------------------------------------------------------------------------------------------ decoder
int qb_plan_create(const qb_dem* d, int32_t m, int32_t W, int32_t F, int32_t n_cor, qb_plan** out) {
    return guard([&] {
        if (!d || !out) throw arg_error("NULL argument");
        std::unique_ptr<qb_plan> p(new qb_plan());
        qb::plan_windows(d->cm, m, W, F, n_cor, p->plan);
        *out = p.release();
    });
}

int qb_plan_create_explicit(int32_t m, int32_t K, int32_t D, int32_t n_windows, const int64_t* dims, const int64_t* h_ptr,
                            const int32_t* h_idx, const double* priors, const int64_t* l_ptr, const int32_t* l_idx,
                            const int64_t* u_ptr, const int32_t* u_idx, qb_plan** out) {
    return guard([&] {
        if (!out || !dims || !h_ptr || !h_idx || !priors || !l_ptr || !u_ptr) throw arg_error("NULL argument");
        if (m <= 0 || K < 0 || D <= 0 || n_windows <= 0) throw arg_error("bad plan dimensions");
        std::unique_ptr<qb_plan> p(new qb_plan());
        qb::WindowPlan& plan = p->plan;
        plan.m = m; plan.K = K; plan.D = D; plan.W = 0; plan.F = 0; plan.num_rounds = D / m - 2; plan.n_cor = n_windows - 1;
        size_t hp = 0, hi = 0, pp = 0, lp = 0, li = 0, up = 0, ui = 0;
        for (int k = 0; k < n_windows; ++k) {
            const int64_t* d = dims + 10 * static_cast<size_t>(k);
            qb::Window w;
            w.row0 = static_cast<int>(d[0]); w.rows = static_cast<int>(d[1]); w.col0 = static_cast<int>(d[2]);
            w.ncols = static_cast<int>(d[3]); w.ncommit = static_cast<int>(d[4]);
            w.urow0 = static_cast<int>(d[8]); w.urows = static_cast<int>(d[9]);
            const int64_t nnz = d[5], nnzl = d[6], nnzu = d[7];
            if (w.rows <= 0 || w.ncols <= 0) throw qb::value_error("a decoding window has no detector rows or no fault columns");
            if (w.row0 < 0 || w.row0 + w.rows > D) throw arg_error("window rows outside the detector range");
            if (w.ncommit < 0 || w.ncommit > w.ncols) throw arg_error("ncommit outside [0, ncols]");
            if (k + 1 < n_windows ? w.urows != m : w.urows != 0) throw arg_error("every window but the last must carry m rows, the last none");
            w.cptr.assign(h_ptr + hp, h_ptr + hp + w.ncols + 1); hp += static_cast<size_t>(w.ncols) + 1;
            if (w.cptr[0] != 0 || w.cptr[w.ncols] != nnz) throw arg_error("window column pointers do not match nnz");
            w.crow.assign(h_idx + hi, h_idx + hi + nnz); hi += static_cast<size_t>(nnz);
            for (int j = 0; j < w.ncols; ++j) {
                if (w.cptr[j + 1] < w.cptr[j]) throw arg_error("column pointers must be non-decreasing");
                for (int64_t e = w.cptr[j]; e < w.cptr[j + 1]; ++e) {
                    if (w.crow[e] < 0 || w.crow[e] >= w.rows) throw arg_error("row index out of range");
                    if (e > w.cptr[j] && w.crow[e] <= w.crow[e - 1]) throw arg_error("rows of a column must be strictly ascending");
                }
            }
            w.priors.assign(priors + pp, priors + pp + w.ncols); pp += static_cast<size_t>(w.ncols);
            w.lptr.assign(l_ptr + lp, l_ptr + lp + w.ncommit + 1); lp += static_cast<size_t>(w.ncommit) + 1;
            if (w.lptr[0] != 0 || w.lptr[w.ncommit] != nnzl) throw arg_error("observable pointers do not match nnz_L");
            if (nnzl) { if (!l_idx) throw arg_error("NULL argument"); w.lidx.assign(l_idx + li, l_idx + li + nnzl); li += static_cast<size_t>(nnzl); }
            for (int32_t v : w.lidx) if (v < 0 || v >= K) throw arg_error("observable index out of range");
            w.uptr.assign(u_ptr + up, u_ptr + up + w.ncommit + 1); up += static_cast<size_t>(w.ncommit) + 1;
            if (w.uptr[0] != 0 || w.uptr[w.ncommit] != nnzu) throw arg_error("carry pointers do not match nnz_U");
            if (nnzu) { if (!u_idx) throw arg_error("NULL argument"); w.uidx.assign(u_idx + ui, u_idx + ui + nnzu); ui += static_cast<size_t>(nnzu); }
            for (int32_t v : w.uidx) if (v < 0 || v >= w.urows) throw arg_error("carry row out of range");
            plan.windows.push_back(std::move(w));
        }
        *out = p.release();
    });
}

void qb_plan_free(qb_plan* p) { delete p; }

int qb_sw_create(qb_ctx* ctx, const qb_plan* plan, const qb_bp_opts* opts, qb_sw** out) {
    return guard([&] {
        if (!ctx || !plan || !out) throw arg_error("NULL argument");
        use_device(ctx);
        std::unique_ptr<qb_sw> sw(new qb_sw());
        sw->ctx = ctx;
        parse_opts(opts, sw->opts);
        sw->plan = plan->plan;
        const int KW = std::max(1, (sw->plan.K + 63) / 64);
        for (const qb::Window& hw : sw->plan.windows) {
            if (hw.urows != 0 && hw.urows != sw->plan.m) throw qb::value_error("carry block of a window is not m rows tall");
            sw->wins.emplace_back(new WinOwned());
            build_window(ctx, hw, KW, sw->opts.precision == 32 ? 32 : 64, (sw->opts.bp_method == 0 && sw->opts.schedule == 0) ? 0 : 1, *sw->wins.back());
            if (sw->opts.schedule == 1) build_serial_slab(ctx, hw, sw->opts.precision == 32 ? 32 : 64, sw->opts.bp_method, *sw->wins.back());
        }
        finish_decoder(sw.get());
        *out = sw.release();
    });
}

int qb_sw_create_single(qb_ctx* ctx, int32_t rows, int32_t cols, const int64_t* indptr, const int32_t* indices, const double* priors,
                        const qb_bp_opts* opts, qb_sw** out) {
    return guard([&] {
        if (!ctx || !indptr || !priors || !out) throw arg_error("NULL argument");
        use_device(ctx);
        std::unique_ptr<qb_sw> sw(new qb_sw());
        sw->ctx = ctx;
        sw->single = true;
        parse_opts(opts, sw->opts);
        qb::Window hw;
        hw.row0 = 0; hw.rows = rows; hw.col0 = 0; hw.ncols = cols; hw.ncommit = 0; hw.urow0 = 0; hw.urows = 0;
        hw.cptr.assign(indptr, indptr + cols + 1);
        if (hw.cptr[0] != 0) throw arg_error("indptr[0] must be 0");
        for (int j = 0; j < cols; ++j) {
            if (hw.cptr[j + 1] < hw.cptr[j]) throw arg_error("indptr must be non-decreasing");
            std::vector<int32_t> r(indices + hw.cptr[j], indices + hw.cptr[j + 1]);
            std::sort(r.begin(), r.end());
            for (size_t i = 0; i < r.size(); ++i) {
                if (r[i] < 0 || r[i] >= rows) throw arg_error("row index out of range");
                if (i && r[i] == r[i - 1]) throw arg_error("duplicate entry in a column");
                hw.crow.push_back(r[i]);
            }
        }
        hw.priors.assign(priors, priors + cols);
        hw.lptr.assign(1, 0);
        hw.uptr.assign(1, 0);
        sw->plan.m = rows; sw->plan.K = 0; sw->plan.D = rows; sw->plan.W = 1; sw->plan.F = 1; sw->plan.n_cor = 0;
        sw->plan.windows.push_back(hw);
        sw->wins.emplace_back(new WinOwned());
        build_window(ctx, sw->plan.windows[0], 1, sw->opts.precision == 32 ? 32 : 64, (sw->opts.bp_method == 0 && sw->opts.schedule == 0) ? 0 : 1, *sw->wins.back());
        if (sw->opts.schedule == 1) build_serial_slab(ctx, sw->plan.windows[0], sw->opts.precision == 32 ? 32 : 64, sw->opts.bp_method, *sw->wins.back());
        finish_decoder(sw.get());
        *out = sw.release();
    });
}

void qb_sw_free(qb_sw* sw) {
    if (!sw) return;
    if (sw->ctx) { cudaSetDevice(sw->ctx->device); cudaStreamSynchronize(sw->ctx->stream); }
    delete sw;
}

int qb_plan_info(const qb_plan* sw, int64_t info[8]) {
    return guard([&] {
        if (!sw || !info) throw arg_error("NULL argument");
        info[0] = static_cast<int64_t>(sw->plan.windows.size()); info[1] = sw->plan.m; info[2] = sw->plan.K; info[3] = sw->plan.D;
        info[4] = sw->plan.W; info[5] = sw->plan.F; info[6] = sw->plan.num_rounds; info[7] = sw->plan.whole_history ? 1 : 0;
    });
}

int qb_plan_window(const qb_plan* sw, int32_t k, int64_t dims[10], int64_t* h_ptr, int32_t* h_idx, double* priors, int64_t* l_ptr,
                 int32_t* l_idx, int64_t* u_ptr, int32_t* u_idx) {
    return guard([&] {
        if (!sw) throw arg_error("NULL argument");
        if (k < 0 || static_cast<size_t>(k) >= sw->plan.windows.size()) throw arg_error("window index out of range");
        const qb::Window& w = sw->plan.windows[k];
        if (dims) {
            dims[0] = w.row0; dims[1] = w.rows; dims[2] = w.col0; dims[3] = w.ncols; dims[4] = w.ncommit;
            dims[5] = static_cast<int64_t>(w.crow.size()); dims[6] = static_cast<int64_t>(w.lidx.size());
            dims[7] = static_cast<int64_t>(w.uidx.size()); dims[8] = w.urow0; dims[9] = w.urows;
        }
        if (h_ptr) memcpy(h_ptr, w.cptr.data(), w.cptr.size() * 8);
        if (h_idx && !w.crow.empty()) memcpy(h_idx, w.crow.data(), w.crow.size() * 4);
        if (priors && !w.priors.empty()) memcpy(priors, w.priors.data(), w.priors.size() * 8);
        if (l_ptr) memcpy(l_ptr, w.lptr.data(), w.lptr.size() * 8);
        if (l_idx && !w.lidx.empty()) memcpy(l_idx, w.lidx.data(), w.lidx.size() * 4);
        if (u_ptr) memcpy(u_ptr, w.uptr.data(), w.uptr.size() * 8);
        if (u_idx && !w.uidx.empty()) memcpy(u_idx, w.uidx.data(), w.uidx.size() * 4);
    });
}

int qb_plan_layout(const qb_plan* p, int32_t k, int32_t precision, int32_t* order, int32_t* slot, double ratios[4]) {
    return guard([&] {
        if (!p) throw arg_error("NULL argument");
        if (k < 0 || static_cast<size_t>(k) >= p->plan.windows.size()) throw arg_error("window index out of range");
        if (precision != 32 && precision != 64) throw arg_error("precision must be 32 or 64");
        const qb::Window& hw = p->plan.windows[k];
        std::vector<int> len(static_cast<size_t>(hw.rows), 0);
        int cw = 0, rs = 1;
        for (int32_t r : hw.crow) rs = std::max(rs, ++len[r]);
        rs |= 1;
        for (int j = 0; j < hw.ncols; ++j) cw = std::max(cw, static_cast<int>(hw.cptr[j + 1] - hw.cptr[j]));
        if (cw > 6) throw qb::unsupported_error("layout search covers column weights <= 6");
        qb::BpLayout naive, lay;
        naive.order.resize(static_cast<size_t>(hw.ncols));
        std::iota(naive.order.begin(), naive.order.end(), 0);
        std::stable_sort(naive.order.begin(), naive.order.end(), [&](int a, int b) { return (hw.cptr[a + 1] - hw.cptr[a]) > (hw.cptr[b + 1] - hw.cptr[b]); });
        naive.slot.resize(hw.crow.size());
        std::fill(len.begin(), len.end(), 0);
        for (size_t e = 0; e < hw.crow.size(); ++e) naive.slot[e] = len[hw.crow[e]]++;
        qb::optimize_bp_layout(hw, rs, precision, lay);
        if (ratios) {
            qb::layout_wavefronts(hw, rs, precision, naive, &ratios[0], &ratios[1]);
            qb::layout_wavefronts(hw, rs, precision, lay, &ratios[2], &ratios[3]);
        }
        if (order) std::copy(lay.order.begin(), lay.order.end(), order);
        if (slot) std::copy(lay.slot.begin(), lay.slot.end(), slot);
    });
}

int qb_sw_decode(qb_sw* sw, const uint8_t* det, uint64_t n, int64_t* pred, qb_stats* stats) {
    return guard([&] {
        if (!sw || (n && (!det || !pred))) throw arg_error("NULL argument");
        if (sw->single) throw arg_error("qb_sw_decode needs a sliding-window decoder (qb_sw_create)");
        qb_ctx* ctx = sw->ctx;
        use_device(ctx);
        cudaStream_t st = ctx->stream;
        const int D = sw->plan.D, K = sw->plan.K;
        // the host-to-device copy of batch k+1 runs on the copy stream while batch k is decoded (two staging buffers; the stream
        // is synchronised at the end of every batch, so a buffer is free again two batches later)
        cudaStream_t cs = ctx->copy_stream();
        // batches: a short first one, so that decoding starts after a small copy and the rest of the input arrives underneath it,
        // then full device batches (the fewer the batches, the fewer post-processing launches with a handful of warps each)
        const uint64_t cap = static_cast<uint64_t>(sw->cap);
        std::vector<uint64_t> cut(1, 0);
        if (n > 32768) cut.push_back(std::min<uint64_t>(cap, std::max<uint64_t>(16384, n / 8 / 64 * 64)));
        while (cut.back() < n) cut.push_back(std::min<uint64_t>(n, cut.back() + cap));
        const size_t nbat = cut.size() - 1;
        size_t big = 0;
        for (size_t k = 0; k < nbat; ++k) big = std::max<size_t>(big, static_cast<size_t>(cut[k + 1] - cut[k]));
        sw->det_bytes.ensure(big * D + 16);
        if (nbat > 1) sw->det_bytes_alt.ensure(big * D + 16);
        auto stage = [&](size_t k) {
            DevBuf& buf = (k & 1) ? sw->det_bytes_alt : sw->det_bytes;
            CK(cudaMemcpyAsync(buf.p, det + cut[k] * D, static_cast<size_t>(cut[k + 1] - cut[k]) * D, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(ctx->ev_ready[k & 1], cs));
        };
        if (nbat) stage(0);
        for (size_t k = 0; k < nbat; ++k) {
            const uint64_t done = cut[k];
            const int nb = static_cast<int>(cut[k + 1] - cut[k]);
            const int which = static_cast<int>(k & 1);
            sw->det_rows.ensure(static_cast<size_t>(nb) * sw->DW * 8 + 16);
            sw->pred.ensure(static_cast<size_t>(nb) * std::max(K, 1) * 8 + 16);
            CK(cudaStreamWaitEvent(st, ctx->ev_ready[which], 0));
            CK(qb::launch_pack_bits((which ? sw->det_bytes_alt : sw->det_bytes).as<uint8_t>(), D, nb, sw->det_rows.as<uint64_t>(), sw->DW, st));
            if (k + 1 < nbat) stage(k + 1);
            decode_batch(sw, sw->det_rows.as<uint64_t>(), nb, false, false, stats);
            CK(qb::launch_expand_pred(sw->acc.as<uint64_t>(), sw->KW, K, nb, sw->pred.as<int64_t>(), st));
            if (K) CK(cudaMemcpyAsync(pred + done * K, sw->pred.p, static_cast<size_t>(nb) * K * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            if (stats) stats->other_launches += 2;
            collect_stats(sw, nb, stats);
        }
    });
}

int qb_sw_decode_packed(qb_sw* sw, const uint64_t* det_rows, uint64_t n, uint64_t* pred_rows, qb_stats* stats) {
    return guard([&] {
        if (!sw || (n && (!det_rows || !pred_rows))) throw arg_error("NULL argument");
        if (sw->single) throw arg_error("qb_sw_decode_packed needs a sliding-window decoder (qb_sw_create)");
        qb_ctx* ctx = sw->ctx;
        use_device(ctx);
        cudaStream_t st = ctx->stream;
        for (uint64_t done = 0; done < n; done += static_cast<uint64_t>(sw->cap)) {
            const int nb = static_cast<int>(std::min<uint64_t>(sw->cap, n - done));
            sw->det_rows.ensure(static_cast<size_t>(nb) * sw->DW * 8 + 16);
            CK(cudaMemcpyAsync(sw->det_rows.p, det_rows + done * sw->DW, static_cast<size_t>(nb) * sw->DW * 8, cudaMemcpyHostToDevice, st));
            decode_batch(sw, sw->det_rows.as<uint64_t>(), nb, false, false, stats);
            CK(cudaMemcpyAsync(pred_rows + done * sw->KW, sw->acc.p, static_cast<size_t>(nb) * sw->KW * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            collect_stats(sw, nb, stats);
        }
    });
}

int qb_bp_decode_batch(qb_sw* sw, const uint8_t* syndromes, uint64_t n, uint8_t* ehat, double* llr, int32_t* iters, uint8_t* converged) {
    return guard([&] {
        if (!sw || (n && !syndromes)) throw arg_error("NULL argument");
        if (!sw->single) throw arg_error("qb_bp_decode_batch needs a single-window decoder (qb_sw_create_single)");
        qb_ctx* ctx = sw->ctx;
        use_device(ctx);
        cudaStream_t st = ctx->stream;
        const qb::WinDev& w = sw->wins[0]->dev;
        const int rows = w.rows, cols = w.ncols;
        for (uint64_t done = 0; done < n; done += static_cast<uint64_t>(sw->cap)) {
            const int nb = static_cast<int>(std::min<uint64_t>(sw->cap, n - done));
            sw->det_bytes.ensure(static_cast<size_t>(nb) * std::max(rows, cols) + 16);
            sw->det_rows.ensure(static_cast<size_t>(nb) * sw->DW * 8 + 16);
            sw->ehat.ensure(static_cast<size_t>(nb) * w.nW32 * 4 + 16);
            sw->iters.ensure(static_cast<size_t>(nb) * 4 + 16);
            sw->conv.ensure(static_cast<size_t>(nb) + 16);
            CK(cudaMemcpyAsync(sw->det_bytes.p, syndromes + done * rows, static_cast<size_t>(nb) * rows, cudaMemcpyHostToDevice, st));
            CK(qb::launch_pack_bits(sw->det_bytes.as<uint8_t>(), rows, nb, sw->det_rows.as<uint64_t>(), sw->DW, st));
            CK(cudaMemsetAsync(sw->ehat.p, 0, static_cast<size_t>(nb) * w.nW32 * 4, st));
            decode_batch(sw, sw->det_rows.as<uint64_t>(), nb, true, llr != nullptr, nullptr);
            if (ehat) {
                // ehat bit rows are u32 words; widen through the u64 unpacker when the stride is even, else via a u32 view
                // (nW32 words per shot) -> unpack treats rows as u64 with words_per_row = nW32/2 only if even; use byte path below
                std::vector<uint32_t> h(static_cast<size_t>(nb) * w.nW32);
                CK(cudaMemcpyAsync(h.data(), sw->ehat.p, h.size() * 4, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                for (int s = 0; s < nb; ++s)
                    for (int j = 0; j < cols; ++j)
                        ehat[(done + s) * cols + j] = static_cast<uint8_t>((h[static_cast<size_t>(s) * w.nW32 + (j >> 5)] >> (j & 31)) & 1u);
            }
            if (llr) {
                const size_t esz = sw->precision / 8;
                std::vector<unsigned char> h(static_cast<size_t>(nb) * cols * esz);
                CK(cudaMemcpy2DAsync(h.data(), static_cast<size_t>(cols) * esz, sw->llr.p, sw->llr_stride * esz, static_cast<size_t>(cols) * esz,
                                     nb, cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                double* out = llr + done * cols;
                if (esz == 8) memcpy(out, h.data(), h.size());
                else for (size_t i = 0; i < static_cast<size_t>(nb) * cols; ++i) out[i] = static_cast<double>(reinterpret_cast<const float*>(h.data())[i]);
            }
            if (iters) CK(cudaMemcpyAsync(iters + done, sw->iters.p, static_cast<size_t>(nb) * 4, cudaMemcpyDeviceToHost, st));
            if (converged) CK(cudaMemcpyAsync(converged + done, sw->conv.p, static_cast<size_t>(nb), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
        }
    });
}

// ------------------------------------------------------------------------------------------------ fused run
int qb_mc_run(qb_ctx* ctx, qb_circuit* c, qb_sw* sw, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint64_t* counts, qb_stats* stats) {
    return guard([&] {
        if (!ctx || !c || !sw || !counts) throw arg_error("NULL argument");
        if (sw->ctx != ctx) throw arg_error("decoder belongs to another context");
        if (sw->single) throw arg_error("qb_mc_run needs a sliding-window decoder");
        if (shot0 & 63) throw arg_error("shot0 must be a multiple of 64");
        if (c->fc.n_det != sw->plan.D || c->fc.n_obs != sw->plan.K) throw arg_error("circuit and decoder disagree on detectors/observables");
        use_device(ctx);
        circuit_to_device(ctx, c);
        cudaStream_t st = ctx->stream;
        qb::FrameArgs a = frame_args(c, seed);
        if (qb::frame_smem_per_warp(a) > 200 * 1024) throw qb::unsupported_error("circuit too large for the shared-memory frame kernel");
        const int K = sw->plan.K;
        ctx->counts.ensure((1 + static_cast<size_t>(a.KW) * 64) * 8);
        CK(cudaMemsetAsync(ctx->counts.p, 0, (1 + static_cast<size_t>(a.KW) * 64) * 8, st));
        const bool prof = sw->opts.profile != 0;
        if (prof) ctx->t_total.begin(st);
        // One frame launch samples a chunk of up to kBatchSlots decoder batches (65536 shots give the frame kernel 1024 warps, a
        // tenth of the machine); the batches of the chunk are then decoded back to back and their statistics read once.
        const uint64_t chunk_cap = static_cast<uint64_t>(sw->cap) * kBatchSlots;
        for (uint64_t done = 0; done < n_shots; done += chunk_cap) {
            const uint64_t nc = std::min<uint64_t>(chunk_cap, n_shots - done);
            const uint64_t nwords = (nc + 63) / 64;
            ctx->det_rows.ensure(nwords * 64 * a.DW * 8 + 16);
            ctx->obs_rows.ensure(nwords * 64 * a.KW * 8 + 16);
            a.word0 = (shot0 + done) / 64;
            a.n_words = nwords;
            a.det_rows = ctx->det_rows.as<uint64_t>();
            a.obs_rows = ctx->obs_rows.as<uint64_t>();
            if (prof) ctx->t_frame.begin(st);
            ctx->det_words.ensure(qb::frame_scratch_bytes(a)); a.det_words = ctx->det_words.as<uint64_t>(); CK(qb::launch_frame(a, st));
            if (prof) ctx->t_frame.end(st);
            if (stats) { stats->frame_launches++; stats->frame_alg_bytes += static_cast<double>(nc) * 8.0 * (a.DW + a.KW); }
            int slot = 0;
            std::vector<int> sizes;
            for (uint64_t off = 0; off < nc; off += static_cast<uint64_t>(sw->cap), ++slot) {
                const uint64_t n = std::min<uint64_t>(sw->cap, nc - off);
                decode_batch(sw, a.det_rows + off * a.DW, static_cast<int>(n), false, false, stats, slot);
                CK(qb::launch_count(sw->acc.as<uint64_t>(), a.obs_rows + off * a.KW, a.KW, K, n, ctx->counts.as<unsigned long long>(), st));
                if (stats) stats->other_launches++;
                sizes.push_back(static_cast<int>(n));
            }
            CK(cudaStreamSynchronize(st));
            for (int b2 = 0; b2 < slot; ++b2) collect_stats(sw, sizes[b2], stats, b2);
        }
        if (prof) { ctx->t_total.end(st); }
        std::vector<unsigned long long> h(1 + static_cast<size_t>(K));
        CK(cudaMemcpyAsync(h.data(), ctx->counts.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (size_t i = 0; i < h.size(); ++i) counts[i] += h[i];
        if (stats && prof) {
            stats->frame_ms += ctx->t_frame.collect();
            stats->total_ms += ctx->t_total.collect();
        }
    });
}

}  // extern "C"
