// quits_b200/csrc/bp_serial.cu -- serial-schedule BP of one sliding window, ONE WARP per shot, messages in an L2-resident slab.
//
// ldpc schedule='serial' is what the reference's wrappers default to (src/quits/decoder/bposd.py:54, bplsd.py:54; every notebook
// keeps it): the columns are updated one after the other in index order, each from the CURRENT messages of its rows
// (oracle/bp_impl.inc, serial branch).  Columns that share no row commute, so the host cuts the column sequence into dependency
// levels and packs each level into steps of independent columns (api.cu, build_serial_slab); a step gives every column LPC lanes
// (one per edge: 6 lanes x 5 columns, 8 x 4 or 16 x 2 -- any column weight up to 16).
//
// What bounds a serial sweep is the dependent chain of ~1100 steps per iteration, so the only way to throughput is MANY shots in
// flight.  The first generation of this kernel (bp.cu, bp_kernel_serial_warp) kept a shot's messages in shared memory: 101 KB per
// shot in fp64, two shots per SM.  Here a shot keeps in shared memory only what the chain depends on -- one summary per ROW:
//   min-sum      (min1 carrying the parity of the syndrome bit and of #{v <= 0} in its sign bit, min2), patched in O(1) when a
//                column writes its new messages; when an edge that held min1 or min2 grows past min2 the warp re-scans that row
//   product-sum  the running product P of the factors tanh(v/2) the sweep has ALREADY renewed in that row
// and the messages live in a global slab per warp (CSR order, so a row is contiguous) that stays L2 resident.  A message is only
// read at its own step and was written a full sweep earlier, so its address is known from the step table and the load is issued
// two steps ahead: the slab's latency never enters the chain.  Product-sum needs "the product of the row's other factors" =
// P (renewed factors before this edge) x S (old factors after it); the suffix products S of a row are rebuilt by the warp at the
// end of every sweep (one multiplicative scan per row) -- O(1) per edge instead of a 35-entry scan, and no division, so every
// intermediate stays in [-1, 1] (the saturation problem a divided-out running product has).
// 5.9 KB (min-sum fp64, gross-code window) or 3 KB (product-sum) of shared memory per shot => 24-32 shots per SM instead of two.
//
// Arithmetic: min-sum is bit-exact with the oracle (same operations in the same order: a minimum does not depend on the order it
// is taken in, the column's prefix / suffix sums are formed as the oracle forms them); product-sum agrees to rounding (the
// association of the row product differs, and CUDA's tanh / log are not libm's).
#include <cstdlib>

#include "bp_common.cuh"
#include "qb_device.h"

namespace qb {

namespace {

constexpr uint32_t kFullMask = 0xFFFFFFFFu;

// off: rsum, syn, cand, accs, car, hist, ebits
__host__ __device__ inline size_t bpss_layout(const WinDev& w, int rsize, bool ps, size_t* off /*[7]*/) {
    size_t o = 0;
    off[0] = o; o += align_up((static_cast<size_t>(w.rows) + 2) * (ps ? 1 : 2) * rsize, 16);
    off[1] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[2] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[3] = o; o += align_up(static_cast<size_t>(w.KW) * 8, 16);
    off[4] = o; o += align_up(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4, 16);
    off[5] = o; o += kSelWords * 4;
    off[6] = o; o += align_up(static_cast<size_t>(w.nW32) * 4, 16);
    return o;
}

template <typename R> struct SlabTabs;
template <> struct SlabTabs<float> {
    static __device__ __forceinline__ const float* v0(const WinDev& w) { return w.ss_v0f; }
    static __device__ __forceinline__ const float* s0(const WinDev& w) { return w.ss_s0f; }
    static __device__ __forceinline__ float2 sum0(const WinDev& w, int i) { return __ldg(w.ss_rsum0f + i); }
    static __device__ __forceinline__ float prior(const WinDev& w, uint32_t j) { return __ldg(w.llr0f + j); }
};
template <> struct SlabTabs<double> {
    static __device__ __forceinline__ const double* v0(const WinDev& w) { return w.ss_v0d; }
    static __device__ __forceinline__ const double* s0(const WinDev& w) { return w.ss_s0d; }
    static __device__ __forceinline__ double2 sum0(const WinDev& w, int i) { return __ldg(w.ss_rsum0d + i); }
    static __device__ __forceinline__ double prior(const WinDev& w, uint32_t j) { return __ldg(w.llr0d + j); }
};

// the two smallest |v| of a row of the slab, by the whole warp; the parity already in the summary's sign bit is kept
template <typename R>
__device__ __noinline__ void slab_rescan_row(const R* vr, const int len, typename Real<R>::pair* srow, const int lane) {
    using RT = Real<R>;
    using CT = Compact<R>;
    R l1 = RT::big(), l2 = RT::big();
    for (int k = lane; k < len; k += 32) {
        const R a = CT::mag(__ldcg(vr + k));
        const bool p = a < l1, q = a < l2;
        l2 = p ? l1 : (q ? a : l2);
        l1 = p ? a : l1;
    }
    const R m1 = warp_min_mag(l1);
    const R m2 = warp_min_mag(l1 == m1 ? l2 : l1);
    const uint32_t holders = __ballot_sync(kFullMask, l1 == m1);
    if (lane == 0) *srow = RT::mk(CT::signed_by(m1, sign_bit<R>(srow->x)), __popc(holders) >= 2 ? m1 : m2);
    __syncwarp();
}

// suffix products of a row's factors: S[k] = f[k+1] * f[k+2] * ... (1 for the last edge), by the whole warp
template <typename R>
__device__ __forceinline__ void slab_suffix_row(const R* f, R* S, const int len, const int lane) {
    using RT = Real<R>;
    R carry = R(1);
    for (int hi = len; hi > 0; hi -= 32) {                    // chunks of 32 from the right
        const int k = hi - 1 - lane;                          // lane 0 holds the rightmost edge of the chunk
        const R x = k >= 0 ? __ldcg(f + k) : R(1);
        R incl = x;                                           // inclusive product over lanes 0..lane (edges k .. hi-1)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const R t = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl = RT::mul(incl, t);
        }
        R excl = __shfl_up_sync(kFullMask, incl, 1);          // product of the edges to the right of k inside the chunk
        if (lane == 0) excl = R(1);
        if (k >= 0) __stcg(S + k, RT::mul(excl, carry));
        carry = RT::mul(carry, __shfl_sync(kFullMask, incl, 31));
    }
}

template <typename R, bool PS, int LPC>
__global__ void __launch_bounds__(32) bp_kernel_serial_slab(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    using CT = Compact<R>;
    using TT = Trans<R>;
    using ST = SlabTabs<R>;
    using Pair = typename RT::pair;
    constexpr int NG = 32 / LPC;                               // columns per step
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off[7];
    bpss_layout(w, sizeof(R), PS, off);
    Pair* rsum = reinterpret_cast<Pair*>(smem_raw + off[0]);   // min-sum
    R* rprod = reinterpret_cast<R*>(smem_raw + off[0]);        // product-sum
    uint32_t* syn = reinterpret_cast<uint32_t*>(smem_raw + off[1]);
    uint32_t* cand = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
    uint32_t* accs = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
    uint32_t* car = reinterpret_cast<uint32_t*>(smem_raw + off[4]);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off[5]);
    uint32_t* ebits = reinterpret_cast<uint32_t*>(smem_raw + off[6]);

    const int lane = threadIdx.x;
    const int grp = lane / LPC, q = lane - grp * LPC, gbase = grp * LPC;
    const int rows = w.rows;
    const uint32_t nnz = static_cast<uint32_t>(w.nnz);
    const size_t slab_elems = static_cast<size_t>(nnz) + 32;   // + one private dummy slot per lane
    R* const V = reinterpret_cast<R*>(b.vscratch) + static_cast<size_t>(blockIdx.x) * slab_elems * (PS ? 2 : 1);
    R* const S = V + slab_elems;                               // product-sum only
    R* const llr_all = reinterpret_cast<R*>(b.llr_buf);
    const uint2* tab = w.ss_rec;
    const int ns = w.ss_nsteps;
    __stcg(V + nnz + lane, R(0));
    if (PS) __stcg(S + nnz + lane, R(0));

    for (int shot = blockIdx.x; shot < b.n_shots; shot += gridDim.x) {
        __syncthreads();
        load_syndrome(w, b, shot, lane, syn, accs, car);
        if (!PS) {                                             // every message starts at its column's prior
            const R* v0 = ST::v0(w);
            for (uint32_t e = lane; e < nnz; e += 32) __stcg(V + e, __ldg(v0 + e));
        }
        __syncthreads();
        for (int i = lane; i < rows + 2; i += 32) {
            if (PS) {
                rprod[i] = i < rows ? R(1) : R(0);             // the dummy row: x = 0, message 0
            } else if (i < rows) {
                const uint32_t sbit = (syn[i >> 5] >> (i & 31)) & 1u;
                const Pair s0 = ST::sum0(w, i);
                rsum[i] = RT::mk(CT::signed_by(s0.x, (sbit + __ldg(w.ss_neg0 + i)) & 1u), s0.y);
            } else {
                rsum[i] = RT::mk(R(0), R(0));
            }
        }
        __syncthreads();
        bool conv = false;
        int it = 1;
        for (; it <= p.max_iter; ++it) {
            const R alpha = static_cast<R>(__ldg(p.alpha + it));
            const bool last = it == p.max_iter;
            for (int i = lane; i < w.rowsW32; i += 32) cand[i] = 0;
            for (int i = lane; i < w.nW32; i += 32) ebits[i] = 0;
            if (last) hist[lane] = 0;
            __syncthreads();
            // product-sum reads the suffix products of the OLD factors: the priors' in the first sweep, the slab's afterwards
            const R* const Sread = PS ? (it == 1 ? ST::s0(w) : S) : nullptr;
            // records of the next four steps in registers (the table is padded by four steps); the slab value and the column's prior
            // of the next two steps are in flight as well
            uint2 r0 = __ldg(tab + lane), r1 = __ldg(tab + 32 + lane), r2 = __ldg(tab + 64 + lane), r3 = __ldg(tab + 96 + lane);
            auto fetch_v = [&](const uint2 rec) -> R {
                if (PS) return rec.x < nnz ? (it == 1 ? __ldg(Sread + rec.x) : __ldcg(Sread + rec.x)) : R(0);
                return __ldcg(V + rec.x);
            };
            auto fetch_l0 = [&](const uint2 rec) -> R {       // lane q == 0 of a column's group carries the column index
                const uint32_t j = rec.y >> 16;
                return (q == 0 && j != 0xFFFFu) ? ST::prior(w, j) : R(0);
            };
            R v0n = fetch_v(r0), v1n = fetch_v(r1);
            R p0n = fetch_l0(r0), p1n = fetch_l0(r1);
#pragma unroll 1
            for (int s = 0; s < ns; ++s) {
                const uint2 rec = r0;
                const R vold = v0n;                            // min-sum: the edge's old message; product-sum: its suffix product
                const R l0q = p0n;
                r0 = r1; r1 = r2; r2 = r3;
                r3 = __ldg(tab + static_cast<size_t>(s + 4) * 32 + lane);
                v0n = v1n; p0n = p1n;
                v1n = fetch_v(r1);                             // step s + 2
                p1n = fetch_l0(r1);
                const uint32_t e = rec.x, row = rec.y & 0xFFFFu;
                const uint32_t j = __shfl_sync(kFullMask, rec.y, gbase) >> 16;
                const R l0 = shfl_real<R>(l0q, gbase);
                const bool active = j != 0xFFFFu && grp < NG;
                const bool real = active && e < nnz;
                R c;
                Pair sm = RT::mk(R(0), R(0));
                R pr = R(0);
                if (PS) {
                    c = R(0);
                    pr = rprod[row];
                    if (real) {
                        const R x = RT::mul(pr, vold);
                        const R lx = TT::lg(TT::div(RT::add(R(1), x), RT::add(R(1), -x)));
                        c = ((syn[row >> 5] >> (row & 31u)) & 1u) ? -lx : lx;
                    }
                } else {
                    sm = rsum[row];
                    const R m1 = CT::mag(sm.x);
                    const R m = CT::mag(vold) == m1 ? sm.y : m1;
                    c = CT::flip(RT::mul(m, alpha), sm.x, vold <= R(0));
                }
                // the column's sums in the oracle's order: v_q = (l0 + c_0 + .. + c_{q-1}) + (c_{W-1} + .. + c_{q+1}); dummy edges add +-0
                R t = l0, pre = R(0), suf = R(0);
#pragma unroll
                for (int k = 0; k < LPC; ++k) {
                    const R ck = shfl_real<R>(c, gbase + k);
                    pre = k == q ? t : pre;
                    t = RT::add(t, ck);
                }
                const R llr = t;
                t = R(0);
#pragma unroll
                for (int k = LPC - 1; k >= 0; --k) {
                    const R ck = shfl_real<R>(c, gbase + k);
                    suf = k == q ? t : suf;
                    t = RT::add(t, ck);
                }
                const R vn = RT::add(pre, suf);
                bool rescan = false;
                if (real) {
                    if (PS) {
                        const R f = TT::th(RT::mul(vn, R(0.5)));
                        __stcg(V + e, f);
                        rprod[row] = RT::mul(pr, f);
                    } else {
                        __stcg(V + e, vn);
                        const R a = CT::mag(vold), an = CT::mag(vn);
                        R m1 = CT::mag(sm.x), m2 = sm.y;
                        const uint32_t par = sign_bit<R>(sm.x) ^ (vold <= R(0) ? 1u : 0u) ^ (vn <= R(0) ? 1u : 0u);
                        if (a == m1) {                        // this edge held the row minimum (or tied with it)
                            if (an <= m2) m1 = an;
                            else rescan = true;
                        } else if (a == m2) {                 // ... the second minimum
                            if (an < m1) { m2 = m1; m1 = an; }
                            else if (an <= m2) m2 = an;
                            else rescan = true;
                        } else if (an < m1) { m2 = m1; m1 = an; }
                        else if (an < m2) m2 = an;
                        rsum[row] = RT::mk(CT::signed_by(m1, par), m2);      // a row to re-scan keeps its new parity here
                    }
                    if (llr <= R(0)) atomicXor(&cand[row >> 5], 1u << (row & 31u));
                }
                if (active && q == 0) {
                    if (llr <= R(0)) atomicOr(&ebits[j >> 5], 1u << (j & 31));
                    if (last || b.write_llr_always) llr_all[static_cast<size_t>(shot) * b.llr_stride + j] = llr;
                    if (last) atomicAdd(&hist[llr_bin<R>(llr, static_cast<R>(w.bin_scale))], 1u);
                }
                __syncwarp();
                if (!PS) {
                    uint32_t mask = __ballot_sync(kFullMask, rescan);
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint32_t rr = __shfl_sync(kFullMask, row, src);
                        const int beg = __ldg(w.rptr + rr), end = __ldg(w.rptr + rr + 1);
                        slab_rescan_row<R>(V + beg, end - beg, rsum + rr, lane);
                    }
                }
            }
            // ---- stop test H e == s (after the full sweep, as the oracle does)
            __syncthreads();
            int mismatch = 0;
            for (int i = lane; i < w.rowsW32; i += 32) mismatch |= cand[i] != syn[i];
            if (!__syncthreads_or(mismatch)) { conv = true; break; }
            if (PS && !last) {
                // the next sweep's "old factors after this edge": suffix products of every row, and the running products start over
                for (int r = 0; r < rows; ++r) {
                    const int beg = __ldg(w.rptr + r), end = __ldg(w.rptr + r + 1);
                    slab_suffix_row<R>(V + beg, S + beg, end - beg, lane);
                }
                for (int i = lane; i < rows; i += 32) rprod[i] = R(1);
                __syncwarp();
            }
        }
        if (it > p.max_iter) it = p.max_iter;
        finish_shot<R, 32, false>(w, b, shot, lane, conv, it, 0u, syn, accs, car, hist, ebits);
    }
}

using KernelPtr = void (*)(const WinDev, const BatchDev, const BpParams);

template <typename R, bool PS>
KernelPtr pick_lpc(int lpc) {
    if (lpc <= 6) return bp_kernel_serial_slab<R, PS, 6>;
    if (lpc <= 8) return bp_kernel_serial_slab<R, PS, 8>;
    return bp_kernel_serial_slab<R, PS, 16>;
}

KernelPtr slab_kernel(int precision, int method, int lpc) {
    if (precision == 32) return method ? pick_lpc<float, true>(lpc) : pick_lpc<float, false>(lpc);
    return method ? pick_lpc<double, true>(lpc) : pick_lpc<double, false>(lpc);
}

// the setup kernel of the product-sum tables: v0[e] = tanh(l0 / 2) of the edge's column, s0 = suffix products of v0 along each row
template <typename R>
__global__ void serial_slab_ps_tables(const int rows, const int32_t* rptr, const uint16_t* rcol, const R* llr0, R* v0, R* s0) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < rows; r += nwarps) {
        const int beg = rptr[r], end = rptr[r + 1];
        for (int k = beg + lane; k < end; k += 32) v0[k] = Trans<R>::th(Real<R>::mul(llr0[rcol[k]], R(0.5)));
        __syncwarp();
        __threadfence_block();
        slab_suffix_row<R>(v0 + beg, s0 + beg, end - beg, lane);
    }
}

}  // namespace

size_t bp_serial_slab_smem_bytes(const WinDev& w, int precision, int method) {
    size_t off[7];
    return bpss_layout(w, precision == 32 ? 4 : 8, method != 0, off);
}

size_t bp_serial_slab_bytes(const WinDev& w, int precision, int method) {
    return (static_cast<size_t>(w.nnz) + 32) * (precision == 32 ? 4 : 8) * (method ? 2 : 1);
}

bool bp_serial_slab_supported(const WinDev& w, int precision, int method) {
    return w.ss_rec != nullptr && w.rowsW32 <= 2048 && bp_serial_slab_smem_bytes(w, precision, method) <= 200 * 1024;
}

cudaError_t bp_serial_slab_configure(const WinDev& w, int precision, int method) {
    const size_t smem = bp_serial_slab_smem_bytes(w, precision, method);
    static size_t have[kMaxDevices][2][2][3] = {};
    size_t& h = have[device_slot()][precision == 32 ? 0 : 1][method ? 1 : 0][w.ss_lpc <= 6 ? 0 : (w.ss_lpc <= 8 ? 1 : 2)];
    if (smem <= h) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(slab_kernel(precision, method, w.ss_lpc), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) h = smem;
    return e;
}

// resident CTAs per SM of the kernel this window runs (registers and shared memory): the persistent grid is sized to it
int bp_serial_slab_ctas_per_sm(const WinDev& w, int precision, int method) {
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, slab_kernel(precision, method, w.ss_lpc), 32,
                                                      bp_serial_slab_smem_bytes(w, precision, method)) != cudaSuccess) return 1;
    static const int cap = [] { const char* e = getenv("QB_SERIAL_CTAS"); return e ? atoi(e) : 0; }();      // experiments
    if (cap > 0 && n > cap) n = cap;
    return n > 0 ? n : 1;
}

cudaError_t launch_bp_serial_slab(const WinDev& w, const BatchDev& b, const BpParams& p, int precision, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    slab_kernel(precision, p.method, w.ss_lpc)<<<grid, 32, bp_serial_slab_smem_bytes(w, precision, p.method), st>>>(w, b, p);
    return cudaGetLastError();
}

cudaError_t launch_serial_slab_ps_tables(const WinDev& w, int precision, void* v0, void* s0, cudaStream_t st) {
    if (precision == 32) serial_slab_ps_tables<float><<<64, 128, 0, st>>>(w.rows, w.rptr, w.rcol, w.llr0f, static_cast<float*>(v0), static_cast<float*>(s0));
    else serial_slab_ps_tables<double><<<64, 128, 0, st>>>(w.rows, w.rptr, w.rcol, w.llr0d, static_cast<double*>(v0), static_cast<double*>(s0));
    return cudaGetLastError();
}

}  // namespace qb
