// quits_b200/csrc/bp.cu -- K3 (+K2/K5 fused): flooding min-sum BP of one sliding window, one shot per CTA (sm_100a).
//
// Replaces the BP stage of ldpc.BpOsdDecoder.decode() as the reference calls it once per shot and window
// (reference src/quits/decoder/sliding_window.py:171,182) together with the glue around it:
// window syndrome slice + carry XOR (:168-169), stop test H e == s, and on convergence the commit
// L_k e / U_k e (:172-175).
//
// Data layout.  The bit->check messages of the shot live in shared memory in padded row-major order
// V[row * RS + slot] (RS odd: a thread-per-row sweep is bank-conflict free, and consecutive columns that share
// a row hit consecutive banks).  The check->bit messages are never materialised: the check sweep reduces each
// row to (min1, min2, argmin slot, sign parity) and the bit sweep rebuilds its <= CW incoming messages from
// those summaries, which gives exactly the value of the forward/backward running-min formulation
// (min over the other edges, sign from syndrome + #{v <= 0}).  All floating-point operations are done in the
// same order as the CPU oracle with explicit round-to-nearest intrinsics (no FMA contraction), so hard
// decisions and posteriors agree bit for bit with the oracle of the same precision:
//   R = double  what ldpc computes in (default; LLR ties -- exact cancellations are common with 9 distinct
//               priors -- resolve as on the CPU)
//   R = float   half the shared memory, twice the resident shots; bit-exact with the fp32 oracle
// The window's column structure (<= CW packed (row, slot) entries per column, stored entry-major so a warp
// reads 128 contiguous bytes) and the prior LLRs are shared by all shots and stream from L2.
// Windows whose messages do not fit in shared memory run the same code with V in a per-CTA global scratch slab
// that stays L2 resident (VGLOBAL).
#include <algorithm>
#include <cfloat>
#include <cstdlib>

#include "qb_device.h"

namespace qb {

namespace {

template <typename R> struct Real;
template <> struct Real<float> {
    using pair = float2;
    static __device__ __forceinline__ float prior(const WinDev& w, int j) { return __ldg(w.llr0f + j); }
    static __device__ __forceinline__ float big() { return FLT_MAX; }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ pair mk(float a, float b) { return make_float2(a, b); }
};
template <> struct Real<double> {
    using pair = double2;
    static __device__ __forceinline__ double prior(const WinDev& w, int j) { return __ldg(w.llr0d + j); }
    static __device__ __forceinline__ double big() { return DBL_MAX; }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ pair mk(double a, double b) { return make_double2(a, b); }
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- hand-off to OSD: the posteriors' 32-bin histogram (built while the last iteration stores them) selects the least
// reliable columns -- tier 1: the smallest bin prefix with >= kOsdSelTarget columns; tier 2: the further bins that still fit
// kOsdSelCap columns in total -- and the CTA writes their
// (order key, column) pairs to HBM, so the OSD warp never scans the full posterior vector (osd.cu, fast path).
constexpr int kSelWords = 36;        // 32 bins, tier-1 count, tier-2 count, the two boundary bins

template <typename R>
__device__ __forceinline__ int llr_bin(const R v, const R scale) {
    return v > R(0) ? 1 + static_cast<int>(fmin(static_cast<double>(v * scale), 30.0)) : 0;
}
__device__ __forceinline__ uint32_t order_key_of(float f) {
    const uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t order_key_of(double f) {
    const uint64_t u = static_cast<uint64_t>(__double_as_longlong(f + 0.0));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

template <typename R, int NT>
__device__ __forceinline__ void select_for_osd(const WinDev& w, const BatchDev& b, int shot, int tid, uint32_t* hist) {
    using KeyT = decltype(order_key_of(R(0)));
    if (tid < 32) {
        uint32_t cum = hist[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, cum, o);
            if (tid >= o) cum += t;
        }
        // tier 1: the smallest bin prefix with >= kOsdSelTarget columns; tier 2: everything else that still fits the buffer
        const uint32_t fits = __ballot_sync(0xFFFFFFFFu, cum <= static_cast<uint32_t>(kOsdSelCap));
        const int b2 = fits ? 31 - __clz(fits) : -1;
        const uint32_t enough = __ballot_sync(0xFFFFFFFFu, cum >= static_cast<uint32_t>(kOsdSelTarget));
        int b1 = enough ? __ffs(enough) - 1 : 31;
        if (b1 > b2) b1 = b2;
        if (tid == 0) { hist[32] = 0; hist[33] = 0; hist[34] = static_cast<uint32_t>(b1 + 1); hist[35] = static_cast<uint32_t>(b2 + 1); }
    }
    __syncthreads();
    const int b1 = static_cast<int>(hist[34]) - 1, b2 = static_cast<int>(hist[35]) - 1;
    const R scale = static_cast<R>(w.bin_scale);
    const R* llr = reinterpret_cast<const R*>(b.llr_buf) + static_cast<size_t>(shot) * b.llr_stride;
    KeyT* gkey = reinterpret_cast<KeyT*>(b.sel_key) + static_cast<size_t>(shot) * kOsdSelCap;
    uint16_t* gidx = b.sel_idx + static_cast<size_t>(shot) * kOsdSelCap;
    for (int j = tid; j < w.ncols; j += NT) {
        const R v = llr[j];
        const int bin = llr_bin<R>(v, scale);
        if (bin <= b2) {                               // tier 1 fills the buffer from the front, tier 2 from the back
            const uint32_t pos = bin <= b1 ? atomicAdd(&hist[32], 1u) : static_cast<uint32_t>(kOsdSelCap - 1) - atomicAdd(&hist[33], 1u);
            gkey[pos] = order_key_of(v);
            gidx[pos] = static_cast<uint16_t>(j);
        }
    }
    __syncthreads();
    if (tid == 0) b.sel_cnt[shot] = static_cast<int>(hist[32] | (hist[33] << 16));
}

// shared-memory layout; returns the total.  off: V, rsum, rmeta, syn, cand, accs, car, hist, ebits (hard decisions as a bit array:
// windows wider than 32 columns per thread, where the per-thread mask runs out)
__host__ __device__ inline size_t bp_layout(const WinDev& w, int rsize, bool vglobal, size_t* off /*[9]*/) {
    size_t o = 0;
    off[0] = o; o += vglobal ? 0 : align_up(static_cast<size_t>(w.rows) * w.RS * rsize, 16);
    off[1] = o; o += align_up(static_cast<size_t>(w.rows) * 2 * rsize, 16);
    off[2] = o; o += align_up(static_cast<size_t>(w.rows) * 4, 16);
    off[3] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[4] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[5] = o; o += align_up(static_cast<size_t>(w.KW) * 8, 16);
    off[6] = o; o += align_up(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4, 16);
    off[7] = o; o += kSelWords * 4;
    off[8] = o; o += align_up(static_cast<size_t>(w.nW32) * 4, 16);
    return o;
}

// ---- window syndrome: detector bits [row0, row0+rows) of this shot, first rows XOR the carry (sliding_window.py:168-169)
__device__ __forceinline__ void load_syndrome(const WinDev& w, const BatchDev& b, int shot, int tid, uint32_t* syn, uint32_t* accs,
                                              uint32_t* car) {
    const int carryW = (w.carry_rows + 31) / 32;
    if (tid < w.rowsW32) {
        const uint32_t* d = b.det32 + static_cast<size_t>(shot) * b.det_stride32;
        const int bit = w.row0 + 32 * tid;
        const int wd = bit >> 5, sh = bit & 31;
        uint32_t v = __ldg(d + wd) >> sh;
        if (sh) v |= __ldg(d + wd + 1) << (32 - sh);
        const int left = w.rows - 32 * tid;
        if (left < 32) v &= (1u << left) - 1u;
        if (32 * tid < b.in_carry_rows) v ^= b.carry[static_cast<size_t>(shot) * b.carry_stride32 + tid];
        syn[tid] = v;
    }
    if (tid < 2 * w.KW) accs[tid] = 0;
    if (tid <= carryW) car[tid] = 0;
}

// ---- after BP: commit acc ^= L e[:ncommit], carry = U e[:ncommit] (sliding_window.py:172-175), or hand the shot to OSD
// hard decisions come either as a per-thread mask (bit k <-> column / record tid + k*NT) or, when `ebits` is given, as a bit
// array over the columns in shared memory (serial kernel)
template <typename R, int NT, bool RECORDS = false>
__device__ __forceinline__ void finish_shot(const WinDev& w, const BatchDev& b, int shot, int tid, bool conv, int it, uint32_t hmask,
                                            const uint32_t* syn, uint32_t* accs, uint32_t* car, uint32_t* hist,
                                            const uint32_t* ebits = nullptr) {
    const int carryW = (w.carry_rows + 31) / 32;
    if (conv) {
        int wd0 = ebits ? tid : 0;
        uint32_t hm = ebits ? (wd0 < w.nW32 ? ebits[wd0] : 0u) : hmask;
        for (;;) {
            if (!hm) {
                if (!ebits) break;
                wd0 += NT;
                if (wd0 >= w.nW32) break;
                hm = ebits[wd0];
                continue;
            }
            const int kk = __ffs(hm) - 1;
            hm &= hm - 1;
            int j = ebits ? 32 * wd0 + kk : tid + kk * NT;
            if (RECORDS) j = static_cast<int>(__ldg(&w.colrec[j].w) & 0xFFFFu);      // record -> original column
            if (b.ehat_out) atomicOr(&b.ehat_out[static_cast<size_t>(shot) * b.ehat_stride32 + (j >> 5)], 1u << (j & 31));
            if (j < w.ncommit) {
                for (int wd = 0; wd < w.KW; ++wd) {
                    const uint64_t lm = __ldg(w.lmask + static_cast<size_t>(j) * w.KW + wd);
                    if (static_cast<uint32_t>(lm)) atomicXor(&accs[2 * wd], static_cast<uint32_t>(lm));
                    if (static_cast<uint32_t>(lm >> 32)) atomicXor(&accs[2 * wd + 1], static_cast<uint32_t>(lm >> 32));
                }
                if (w.carry_rows) {
                    for (int q = __ldg(w.uptr + j); q < __ldg(w.uptr + j + 1); ++q) {
                        const uint32_t r = __ldg(w.uidx + q);
                        atomicXor(&car[r >> 5], 1u << (r & 31));
                    }
                }
            }
        }
        __syncthreads();
        if (tid < w.KW) {
            const uint64_t v = (static_cast<uint64_t>(accs[2 * tid + 1]) << 32) | accs[2 * tid];
            b.acc[static_cast<size_t>(shot) * w.KW + tid] ^= v;
        }
        if (tid < carryW) b.carry[static_cast<size_t>(shot) * b.carry_stride32 + tid] = car[tid];
    } else {
        // post-carry syndrome for OSD; the posteriors are already in llr_buf
        if (tid < w.rowsW32) b.syn_buf[static_cast<size_t>(shot) * b.syn_stride32 + tid] = syn[tid];
        if (tid == 0) {
            const int slot = atomicAdd(b.fail_count, 1);
            b.fail_list[slot] = shot;
        }
        if (b.sel_cnt) {
            __threadfence_block();
            __syncthreads();                                  // the posteriors of this shot are complete
            select_for_osd<R, NT>(w, b, shot, tid, hist);
        }
    }
    if (tid == 0) {
        if (conv) atomicAdd(&b.stats[0], 1ull);
        atomicAdd(&b.stats[1], static_cast<unsigned long long>(it));
        if (b.iters_out) b.iters_out[shot] = it;
        if (b.conv_out) b.conv_out[shot] = conv ? 1 : 0;
    }
}

template <typename R, int CW, int NT, int MINB, bool VGLOBAL>
__global__ void __launch_bounds__(NT, MINB) bp_kernel(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off[9];
    bp_layout(w, sizeof(R), VGLOBAL, off);
    R* V = VGLOBAL ? reinterpret_cast<R*>(b.vscratch) + static_cast<size_t>(blockIdx.x) * (static_cast<size_t>(w.rows) * w.RS)
                   : reinterpret_cast<R*>(smem_raw + off[0]);
    typename RT::pair* rsum = reinterpret_cast<typename RT::pair*>(smem_raw + off[1]);
    uint32_t* rmeta = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
    uint32_t* syn = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
    uint32_t* cand = reinterpret_cast<uint32_t*>(smem_raw + off[4]);
    uint32_t* accs = reinterpret_cast<uint32_t*>(smem_raw + off[5]);
    uint32_t* car = reinterpret_cast<uint32_t*>(smem_raw + off[6]);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off[7]);
    uint32_t* ebits = reinterpret_cast<uint32_t*>(smem_raw + off[8]);

    const int tid = threadIdx.x;
    const int rows = w.rows, ncols = w.ncols, RS = w.RS, npad = w.ncols_pad;
    const bool wide = ncols > 32 * NT;          // more than 32 columns per thread: hard decisions go to the bit array
    R* const llr_all = reinterpret_cast<R*>(b.llr_buf);

    for (int shot = blockIdx.x; shot < b.n_shots; shot += gridDim.x) {
        __syncthreads();
        load_syndrome(w, b, shot, tid, syn, accs, car);
        for (int i = tid; i < rows * RS; i += NT) V[i] = RT::big();       // padding slots: never the minimum, never negative
        __syncthreads();
        for (int j = tid; j < ncols; j += NT) {
            const R l0 = RT::prior(w, j);
#pragma unroll
            for (int q = 0; q < CW; ++q) {
                const uint32_t e = __ldg(w.colE + static_cast<size_t>(q) * npad + j);
                if (e != kNoEdge) V[(e >> 8) * RS + (e & 255u)] = l0;
            }
        }
        __syncthreads();

        uint32_t hmask = 0;          // hard decisions of this thread's columns (bit k <-> column tid + k*NT)
        bool conv = false;
        int it = 1;
        for (; it <= p.max_iter; ++it) {
            const R alpha = static_cast<R>(__ldg(p.alpha + it));
            // ---- check sweep: one thread per row
            for (int i = tid; i < rows; i += NT) {
                const R* vr = V + i * RS;
                R m1 = RT::big(), m2 = RT::big();
                uint32_t arg = 0, neg = (syn[i >> 5] >> (i & 31)) & 1u;
#pragma unroll 5
                for (int s = 0; s < RS; ++s) {
                    const R v = vr[s];
                    const R a = RT::abs(v);
                    neg += v <= R(0) ? 1u : 0u;
                    if (a < m1) { m2 = m1; m1 = a; arg = static_cast<uint32_t>(s); }
                    else if (a < m2) { m2 = a; }
                }
                rsum[i] = RT::mk(m1, m2);
                rmeta[i] = arg | (neg << 31);
            }
            if (tid < w.rowsW32) cand[tid] = 0;
            const bool last = it == p.max_iter;
            if (last && tid < 32) hist[tid] = 0;
            if (wide) for (int i = tid; i < w.nW32; i += NT) ebits[i] = 0;
            __syncthreads();
            // ---- bit sweep: one thread per column
            hmask = 0;
            int k = 0;
            for (int j = tid; j < ncols; j += NT, ++k) {
                uint32_t e[CW];
#pragma unroll
                for (int q = 0; q < CW; ++q) e[q] = __ldg(w.colE + static_cast<size_t>(q) * npad + j);
                const R l0 = RT::prior(w, j);
                R c[CW], vn[CW];
                int addr[CW];
#pragma unroll
                for (int q = 0; q < CW; ++q) {
                    c[q] = R(0);
                    addr[q] = 0;
                    if (e[q] != kNoEdge) {
                        const uint32_t row = e[q] >> 8, slot = e[q] & 255u;
                        addr[q] = static_cast<int>(row) * RS + static_cast<int>(slot);
                        const R v = V[addr[q]];
                        const typename RT::pair s = rsum[row];
                        const uint32_t meta = rmeta[row];
                        const R mag = slot == (meta & 0x7FFFFFFFu) ? s.y : s.x;
                        const uint32_t odd = (meta >> 31) ^ (v <= R(0) ? 1u : 0u);
                        c[q] = RT::mul(mag, odd ? -alpha : alpha);
                    }
                }
                R t = l0;
#pragma unroll
                for (int q = 0; q < CW; ++q) { vn[q] = t; t = RT::add(t, c[q]); }
                const R llr = t;
                t = R(0);
#pragma unroll
                for (int q = CW - 1; q >= 0; --q) { vn[q] = RT::add(vn[q], t); t = RT::add(t, c[q]); }
#pragma unroll
                for (int q = 0; q < CW; ++q)
                    if (e[q] != kNoEdge) V[addr[q]] = vn[q];
                if (llr <= R(0)) {
                    if (wide) atomicOr(&ebits[j >> 5], 1u << (j & 31));
                    else hmask |= 1u << k;
#pragma unroll
                    for (int q = 0; q < CW; ++q)
                        if (e[q] != kNoEdge) atomicXor(&cand[e[q] >> 13], 1u << ((e[q] >> 8) & 31u));
                }
                if (last || b.write_llr_always) llr_all[static_cast<size_t>(shot) * b.llr_stride + j] = llr;
                if (last) atomicAdd(&hist[llr_bin<R>(llr, static_cast<R>(w.bin_scale))], 1u);
            }
            // ---- stop test H e == s
            __syncthreads();
            const int mismatch = tid < w.rowsW32 ? (cand[tid] != syn[tid]) : 0;
            if (!__syncthreads_or(mismatch)) { conv = true; break; }
        }
        if (it > p.max_iter) it = p.max_iter;

        finish_shot<R, NT, false>(w, b, shot, tid, conv, it, hmask, syn, accs, car, hist, wide ? ebits : nullptr);
    }
}

// =====================================================================================================================
// Compact variant (the one the BB / HGP windows use): same arithmetic, fewer instructions and bytes.
//   * one 16-byte record per column (6 u16 message addresses + prior index) -> a single LDG.128 per column and sweep,
//     prefetched one record ahead;
//   * records are sorted by column weight and every warp runs the code path of its heaviest column (the weight field of
//     a record holds the warp's maximum).  Lighter columns are padded with DUMMY edges that point at a private dummy slot
//     of the handling thread in a dummy row (index >= `rows`) whose summary is (0, 0): their check->bit message is +-0, and
//     x + (+-0) == x exactly, so the padding needs no predication at all (and, the slots being private, races with nobody);
//   * the row summary is (min1 with the row's sign parity in its sign bit, min2): the edge that holds the minimum is
//     recognised by |v| == min1 (if two edges tie, min2 == min1 and either choice gives the same value), so no argmin;
//     the sign of a message is applied by XOR on the sign bit (mul(m, -a) == -mul(m, a) exactly in round-to-nearest);
//   * iteration 1 needs no message array at all: every bit->check message is the column's prior, so the row summaries
//     are precomputed per window (rsum0) and only the syndrome bit is folded in;
//   * rows are swept to their true length (rlen) with two independent (min1, min2) chains and explicit compare/select
//     (fmin/fmax on doubles expand to NaN-aware sequences several times longer).
// =====================================================================================================================
template <typename R> struct Compact;
template <> struct Compact<float> {
    static __device__ __forceinline__ float2 sum0(const WinDev& w, int i) { return __ldg(w.rsum0f + i); }
    static __device__ __forceinline__ const float* ptab(const WinDev& w) { return w.ptabf; }
    static __device__ __forceinline__ float signed_by(float m, uint32_t neg) { return __uint_as_float(__float_as_uint(m) | (neg << 31)); }
    static __device__ __forceinline__ float mag(float m) { return __uint_as_float(__float_as_uint(m) & 0x7FFFFFFFu); }
    // x with its sign flipped when (sign bit of s) xor neg
    static __device__ __forceinline__ float flip(float x, float s, bool neg) {
        return __uint_as_float(__float_as_uint(x) ^ ((__float_as_uint(s) & 0x80000000u) ^ (neg ? 0x80000000u : 0u)));
    }
};
template <> struct Compact<double> {
    static __device__ __forceinline__ double2 sum0(const WinDev& w, int i) { return __ldg(w.rsum0d + i); }
    static __device__ __forceinline__ const double* ptab(const WinDev& w) { return w.ptabd; }
    static __device__ __forceinline__ double signed_by(double m, uint32_t neg) {
        return __hiloint2double(__double2hiint(m) | static_cast<int>(neg << 31), __double2loint(m));
    }
    static __device__ __forceinline__ double mag(double m) { return __hiloint2double(__double2hiint(m) & 0x7FFFFFFF, __double2loint(m)); }
    static __device__ __forceinline__ double flip(double x, double s, bool neg) {
        const uint32_t f = (static_cast<uint32_t>(__double2hiint(s)) & 0x80000000u) ^ (neg ? 0x80000000u : 0u);
        return __hiloint2double(static_cast<int>(static_cast<uint32_t>(__double2hiint(x)) ^ f), __double2loint(x));
    }
};

// off: V, rsum, syn, cand, accs, car, ptab.  Behind the rows*RS real message slots V carries one private dummy slot per thread
// (dummy edges of a record handled by thread t point at slot rows*RS + t: nothing is shared, so nothing races), and rsum carries
// the dummy rows those addresses divide down to.
__host__ __device__ inline int bp_dummy_slots(int rsize) { return rsize == 4 ? 256 : 512; }      // = threads of the compact kernel
__host__ __device__ inline int bp_dummy_rows(const WinDev& w, int rsize) { return bp_dummy_slots(rsize) / w.RS + 2; }
__host__ __device__ inline size_t bpc_layout(const WinDev& w, int rsize, size_t* off /*[8]*/) {
    size_t o = 0;
    off[0] = o; o += align_up((static_cast<size_t>(w.rows) * w.RS + bp_dummy_slots(rsize)) * rsize, 16);
    off[1] = o; o += align_up(static_cast<size_t>(w.rows + bp_dummy_rows(w, rsize)) * 2 * rsize, 16);
    off[2] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[3] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[4] = o; o += align_up(static_cast<size_t>(w.KW) * 8, 16);
    off[5] = o; o += align_up(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4, 16);
    off[6] = o; o += align_up(static_cast<size_t>(w.n_ptab) * rsize, 16);
    off[7] = o; o += kSelWords * 4;
    return o;
}

// running two smallest magnitudes of a row
template <typename R>
__device__ __forceinline__ void min2_step(const R v, R& m1, R& m2, uint32_t& neg) {
    const R a = Compact<R>::mag(v);
    neg += v <= R(0) ? 1u : 0u;
    const bool p = a < m1, q = a < m2;
    m2 = p ? m1 : (q ? a : m2);
    m1 = p ? a : m1;
}

// one column of a warp whose heaviest column has W edges (dummy edges included, no predication)
template <typename R, int W, bool FIRST>
__device__ __forceinline__ R column_update(const uint4 rec, const R l0, const R alpha, R* V, const typename Real<R>::pair* rsum,
                                           uint32_t* cand, const uint32_t magic, const uint32_t rows) {
    using RT = Real<R>;
    using CT = Compact<R>;
    const uint32_t e[6] = {rec.x & 0xFFFFu, rec.x >> 16, rec.y & 0xFFFFu, rec.y >> 16, rec.z & 0xFFFFu, rec.z >> 16};
    R c[W], vn[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const R v = FIRST ? l0 : V[e[q]];
        const typename RT::pair s = rsum[__umulhi(e[q], magic)];
        const R m1 = CT::mag(s.x);
        const R m = CT::mag(v) == m1 ? s.y : m1;
        c[q] = CT::flip(RT::mul(m, alpha), s.x, v <= R(0));
    }
    R t = l0;
#pragma unroll
    for (int q = 0; q < W; ++q) { vn[q] = t; t = RT::add(t, c[q]); }
    const R llr = t;
    t = R(0);
#pragma unroll
    for (int q = W - 1; q >= 0; --q) { vn[q] = RT::add(vn[q], t); t = RT::add(t, c[q]); }
#pragma unroll
    for (int q = 0; q < W; ++q) V[e[q]] = vn[q];
    if (llr <= R(0)) {
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const uint32_t row = __umulhi(e[q], magic);
            if (row < rows) atomicXor(&cand[row >> 5], 1u << (row & 31u));
        }
    }
    return llr;
}

// ---- product-sum (ldpc bp_method 'product_sum'): check->bit message = 2 atanh( prod_{others} tanh(v/2) ), sign from the syndrome.
// The message array holds tanh(v/2) instead of v (one tanh per edge and iteration, taken when the message is written).
// The row summary is (s * prod of the non-zero tanh(v/2), number of zero factors); "the others" is obtained by dividing the
// row product by the edge's own factor, which differs from ldpc's prefix/suffix products by rounding only (x is formed
// explicitly, as ldpc does, so that saturation -- x rounding to exactly 1, message +inf -- happens at the same places) (no bit parity with
// the CPU here anyway: libm and CUDA tanh/log differ in the last ulps).
template <typename R> struct Trans;
template <> struct Trans<float> {
    static __device__ __forceinline__ float th(float x) { return tanhf(x); }
    static __device__ __forceinline__ float lg(float x) { return logf(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Trans<double> {
    static __device__ __forceinline__ double th(double x) { return tanh(x); }
    static __device__ __forceinline__ double lg(double x) { return log(x); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

template <typename R, int W>
__device__ __forceinline__ R column_update_ps(const uint4 rec, const R l0, R* V, const typename Real<R>::pair* rsum,
                                              uint32_t* cand, const uint32_t magic, const uint32_t rows) {
    using RT = Real<R>;
    using TT = Trans<R>;
    const uint32_t e[6] = {rec.x & 0xFFFFu, rec.x >> 16, rec.y & 0xFFFFu, rec.y >> 16, rec.z & 0xFFFFu, rec.z >> 16};
    R c[W], vn[W];
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const R t = V[e[q]];                                   // tanh(v/2) of the message, stored by the previous sweep
        const typename RT::pair s = rsum[__umulhi(e[q], magic)];
        R x = R(0);
        if (s.y == R(0)) x = TT::div(s.x, t);
        else if (s.y == R(1) && t == R(0)) x = s.x;
        c[q] = TT::lg(TT::div(RT::add(R(1), x), RT::add(R(1), -x)));
    }
    R t = l0;
#pragma unroll
    for (int q = 0; q < W; ++q) { vn[q] = t; t = RT::add(t, c[q]); }
    const R llr = t;
    t = R(0);
#pragma unroll
    for (int q = W - 1; q >= 0; --q) { vn[q] = RT::add(vn[q], t); t = RT::add(t, c[q]); }
#pragma unroll
    for (int q = 0; q < W; ++q) V[e[q]] = TT::th(RT::mul(vn[q], R(0.5)));
    if (llr <= R(0)) {
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const uint32_t row = __umulhi(e[q], magic);
            if (row < rows) atomicXor(&cand[row >> 5], 1u << (row & 31u));
        }
    }
    return llr;
}

template <typename R>
__device__ __forceinline__ R column_dispatch_ps(const uint4 rec, const R l0, R* V, const typename Real<R>::pair* rsum, uint32_t* cand,
                                                const uint32_t magic, const uint32_t rows) {
    switch (rec.w >> 28) {
    case 0: return l0;
    case 1: return column_update_ps<R, 1>(rec, l0, V, rsum, cand, magic, rows);
    case 2: return column_update_ps<R, 2>(rec, l0, V, rsum, cand, magic, rows);
    case 3: return column_update_ps<R, 3>(rec, l0, V, rsum, cand, magic, rows);
    case 4: return column_update_ps<R, 4>(rec, l0, V, rsum, cand, magic, rows);
    case 5: return column_update_ps<R, 5>(rec, l0, V, rsum, cand, magic, rows);
    default: return column_update_ps<R, 6>(rec, l0, V, rsum, cand, magic, rows);
    }
}

template <typename R, bool FIRST>
__device__ __forceinline__ R column_dispatch(const uint4 rec, const R l0, const R alpha, R* V, const typename Real<R>::pair* rsum,
                                             uint32_t* cand, const uint32_t magic, const uint32_t rows) {
    switch (rec.w >> 28) {                  // warp-uniform: the weight of the warp's heaviest column
    case 0: return l0;
    case 1: return column_update<R, 1, FIRST>(rec, l0, alpha, V, rsum, cand, magic, rows);
    case 2: return column_update<R, 2, FIRST>(rec, l0, alpha, V, rsum, cand, magic, rows);
    case 3: return column_update<R, 3, FIRST>(rec, l0, alpha, V, rsum, cand, magic, rows);
    case 4: return column_update<R, 4, FIRST>(rec, l0, alpha, V, rsum, cand, magic, rows);
    case 5: return column_update<R, 5, FIRST>(rec, l0, alpha, V, rsum, cand, magic, rows);
    default: return column_update<R, 6, FIRST>(rec, l0, alpha, V, rsum, cand, magic, rows);
    }
}

template <typename R, int NT, int MINB, bool PS>
__global__ void __launch_bounds__(NT, MINB) bp_kernel_compact(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    using CT = Compact<R>;
    using Pair = typename RT::pair;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off[8];
    bpc_layout(w, sizeof(R), off);
    R* V = reinterpret_cast<R*>(smem_raw + off[0]);
    Pair* rsum = reinterpret_cast<Pair*>(smem_raw + off[1]);
    uint32_t* syn = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
    uint32_t* cand = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
    uint32_t* accs = reinterpret_cast<uint32_t*>(smem_raw + off[4]);
    uint32_t* car = reinterpret_cast<uint32_t*>(smem_raw + off[5]);
    R* ptab = reinterpret_cast<R*>(smem_raw + off[6]);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off[7]);

    const int tid = threadIdx.x;
    const int rows = w.rows, npad = w.ncols_pad, RS = w.RS;
    const uint32_t magic = w.rs_magic;
    R* const llr_all = reinterpret_cast<R*>(b.llr_buf);
    for (int i = tid; i < w.n_ptab; i += NT) ptab[i] = CT::ptab(w)[i];
    // the dummy rows: their messages are +-0 (min-sum: minima 0; product-sum: two zero factors); one dummy slot per thread
    if (tid < bp_dummy_rows(w, sizeof(R))) rsum[rows + tid] = RT::mk(R(0), PS ? R(2) : R(0));
    V[rows * RS + tid] = R(0);

    for (int shot = blockIdx.x; shot < b.n_shots; shot += gridDim.x) {
        __syncthreads();
        load_syndrome(w, b, shot, tid, syn, accs, car);
        __syncthreads();
        uint32_t hmask = 0;          // bit k <-> record tid + k*NT (records are the columns sorted by weight)
        bool conv = false;
        int it = 1;
        if (PS) {                    // product-sum has no closed form for iteration 1: start from the priors in the message array
            for (int r = tid; r < npad; r += NT) {
                const uint4 rec = __ldg(w.colrec + r);
                const R t0 = Trans<R>::th(RT::mul(ptab[(rec.w >> 16) & 0xFFFu], R(0.5)));
                const uint32_t e[6] = {rec.x & 0xFFFFu, rec.x >> 16, rec.y & 0xFFFFu, rec.y >> 16, rec.z & 0xFFFFu, rec.z >> 16};
#pragma unroll
                for (int q = 0; q < 6; ++q)
                    if (e[q] < static_cast<uint32_t>(rows * RS)) V[e[q]] = t0;
            }
            __syncthreads();
        }
        for (; it <= p.max_iter; ++it) {
            const R alpha = static_cast<R>(__ldg(p.alpha + it));
            const bool first = !PS && it == 1;
            uint4 rec = __ldg(w.colrec + tid);               // NT <= npad is not guaranteed: colrec is padded to a multiple of NT
            // ---- check sweep: one thread per row -> (min1 | parity sign, min2)
            for (int i = tid; i < rows; i += NT) {
                uint32_t neg = (syn[i >> 5] >> (i & 31)) & 1u;
                Pair s;
                if (PS) {
                    const R* vr = V + i * RS;
                    const int len = __ldg(w.rlen + i);
                    R prod = (neg & 1u) ? R(-1) : R(1);
                    int zc = 0;
                    for (int q = 0; q < len; ++q) {
                        const R t = vr[q];
                        if (t == R(0)) ++zc;
                        else prod = RT::mul(prod, t);
                    }
                    rsum[i] = RT::mk(prod, static_cast<R>(zc));
                    continue;
                }
                if (first) {
                    s = CT::sum0(w, i);
                    neg += __ldg(w.neg0 + i);
                } else {
                    const R* vr = V + i * RS;
                    const int len = __ldg(w.rlen + i);
                    R m1a = RT::big(), m2a = RT::big(), m1b = RT::big(), m2b = RT::big();
                    int q = 0;
#pragma unroll 2
                    for (; q + 1 < len; q += 2) {
                        min2_step<R>(vr[q], m1a, m2a, neg);
                        min2_step<R>(vr[q + 1], m1b, m2b, neg);
                    }
                    if (q < len) min2_step<R>(vr[q], m1a, m2a, neg);
                    const bool lo = m1b < m1a;
                    const R m1 = lo ? m1b : m1a, mo = lo ? m1a : m1b;       // smaller / larger of the two chain minima
                    const R m2c = m2b < m2a ? m2b : m2a;
                    s = RT::mk(m1, m2c < mo ? m2c : mo);
                }
                s.x = CT::signed_by(s.x, neg & 1u);
                rsum[i] = s;
            }
            if (tid < w.rowsW32) cand[tid] = 0;
            const bool last = it == p.max_iter;
            if (last && tid < 32) hist[tid] = 0;
            __syncthreads();
            // ---- bit sweep: one thread per column record
            hmask = 0;
            int k = 0;
            for (int r = tid; r < npad; r += NT, ++k) {
                const uint4 cur = rec;
                if (r + NT < npad) rec = __ldg(w.colrec + r + NT);
                const R l0 = ptab[(cur.w >> 16) & 0xFFFu];
                R llr;
                const uint32_t urows = static_cast<uint32_t>(rows);
                if (PS) llr = column_dispatch_ps<R>(cur, l0, V, rsum, cand, magic, urows);
                else llr = first ? column_dispatch<R, true>(cur, l0, alpha, V, rsum, cand, magic, urows)
                                 : column_dispatch<R, false>(cur, l0, alpha, V, rsum, cand, magic, urows);
                const uint32_t j = cur.w & 0xFFFFu;                    // original column; 0xFFFF marks a padding record
                if (j != 0xFFFFu) {
                    if (llr <= R(0)) hmask |= 1u << k;
                    if (last || b.write_llr_always) llr_all[static_cast<size_t>(shot) * b.llr_stride + j] = llr;
                    if (last) atomicAdd(&hist[llr_bin<R>(llr, static_cast<R>(w.bin_scale))], 1u);
                }
            }
            // ---- stop test H e == s
            __syncthreads();
            const int mismatch = tid < w.rowsW32 ? (cand[tid] != syn[tid]) : 0;
            if (!__syncthreads_or(mismatch)) { conv = true; break; }
        }
        if (it > p.max_iter) it = p.max_iter;
        // map record bits back to columns for the commit
        finish_shot<R, NT, true>(w, b, shot, tid, conv, it, hmask, syn, accs, car, hist);
    }
}

// =====================================================================================================================
// Flooding min-sum, second form (bp_kernel_ms2; the headline path).  Same arithmetic and data as bp_kernel_compact<.., false>,
// half the instructions per edge and iteration and fewer shared-memory wavefronts:
//   * messages are stored with a CANONICAL SIGN: the sign bit of a stored message is set exactly when v <= 0 (a message that
//     is exactly +0 is stored as -0; |v| and every comparison are unchanged).  The row parity is then the XOR of the high words
//     (one LOP3 per edge instead of a compare and an add) and the sign of a check->bit message is one LOP3 on the high word;
//   * |.| comparisons use the compare instruction's operand modifiers (DSETP |a|, |b|) and the running minima of a row are kept
//     as SIGNED values (the one with the smallest magnitude): no per-edge integer masking, no register-pair copies;
//   * the row summary is two arrays, min1[row] and min2[row], both carrying the row parity in their sign bit: the bit sweep loads
//     min1 (8 bytes instead of a 16-byte pair) and only the edge that holds the minimum -- one in ~30 -- loads min2, predicated;
//   * ms_scaling_factor == 1 (ldpc's and the reference's default; the wrappers do not plumb the option) skips the multiply
//     (x * 1.0 == x exactly);
//   * the backward pass starts from t = c[W-1] instead of 0 + c[W-1] and the last edge's message is the prefix itself: the two
//     forms differ only in the sign of an exact zero, which the canonical sign absorbs;
//   * column records are walked in warp chunks (32 records) by weight segment -- `chunk_end[W]` -- so the code path of a chunk
//     is selected by loop structure, not by a per-record switch; hard decisions are one ballot word per chunk.
// =====================================================================================================================
template <typename R> struct Bits;
template <> struct Bits<float> {
    static __device__ __forceinline__ uint32_t hi(float x) { return __float_as_uint(x); }
    static __device__ __forceinline__ float with_hi(float, uint32_t h) { return __uint_as_float(h); }
};
template <> struct Bits<double> {
    static __device__ __forceinline__ uint32_t hi(double x) { return static_cast<uint32_t>(__double2hiint(x)); }
    static __device__ __forceinline__ double with_hi(double x, uint32_t h) { return __hiloint2double(static_cast<int>(h), __double2loint(x)); }
};

// sign bit set <=> x <= 0  (x == +0 becomes -0)
template <typename R>
__device__ __forceinline__ R canonical_sign(const R x) {
    return Bits<R>::with_hi(x, x == R(0) ? 0x80000000u : Bits<R>::hi(x));
}

// off: rm1, rm2, V, syn, cand, accs, car, ptab, hist, ebits
__host__ __device__ inline size_t bpm_layout(const WinDev& w, int rsize, size_t* off /*[10]*/) {
    size_t o = 0;
    const size_t rrows = static_cast<size_t>(w.rows + bp_dummy_rows(w, rsize));
    off[0] = o; o += align_up(rrows * rsize, 16);
    off[1] = o; o += align_up(rrows * rsize, 16);
    off[2] = o; o += align_up((static_cast<size_t>(w.rows) * w.RS + bp_dummy_slots(rsize)) * rsize, 16);
    off[3] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[4] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 2 * 4, 16);          // candidate syndrome, double-buffered
    off[5] = o; o += align_up(static_cast<size_t>(w.KW) * 8, 16);
    off[6] = o; o += align_up(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4, 16);
    off[7] = o; o += align_up(static_cast<size_t>(w.n_ptab) * rsize, 16);
    off[8] = o; o += kSelWords * 4;
    off[9] = o; o += align_up(static_cast<size_t>(w.nW32) * 4, 16);
    return o;
}

// one element of a row scan: (m1, m2) = the two entries of smallest magnitude seen so far, signed; par ^= sign word
template <typename R>
__device__ __forceinline__ void min2_signed(const R v, R& m1, R& m2, uint32_t& par) {
    par ^= Bits<R>::hi(v);
    const bool p = fabs(v) < fabs(m1), q = fabs(v) < fabs(m2);
    m2 = p ? m1 : (q ? v : m2);
    m1 = p ? v : m1;
}

// shared-memory accesses by 32-bit shared-window address (the bit sweep does its own address arithmetic on byte offsets)
template <typename R> struct Sh;
template <> struct Sh<double> {
    static constexpr uint32_t kOffMask = 0x7FFF8u;      // byte offset of a u16 slot index
    static __device__ __forceinline__ uint32_t off_lo(uint32_t w) { return (w << 3) & kOffMask; }
    static __device__ __forceinline__ uint32_t off_hi(uint32_t w) { return (w >> 13) & kOffMask; }
    static __device__ __forceinline__ uint32_t ptab_off(uint32_t w) { return (w >> 13) & 0x7FF8u; }
    static __device__ __forceinline__ double ld(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a)); return v; }
    static __device__ __forceinline__ void st(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
    // m = [a1]; if |v| == |m| then m = [a2]   (the second load is predicated: one edge in ~30 takes it)
    static __device__ __forceinline__ double row_msg(uint32_t a1, uint32_t a2, double v) {
        double m;
        asm volatile("{\n .reg .pred p;\n .reg .f64 av, am;\n ld.shared.f64 %0, [%1];\n abs.f64 av, %3;\n abs.f64 am, %0;\n"
                     " setp.eq.f64 p, av, am;\n @p ld.shared.f64 %0, [%2];\n}"
                     : "=&d"(m) : "r"(a1), "r"(a2), "d"(v));
        return m;
    }
};
template <> struct Sh<float> {
    static constexpr uint32_t kOffMask = 0x3FFFCu;
    static __device__ __forceinline__ uint32_t off_lo(uint32_t w) { return (w << 2) & kOffMask; }
    static __device__ __forceinline__ uint32_t off_hi(uint32_t w) { return (w >> 14) & kOffMask; }
    static __device__ __forceinline__ uint32_t ptab_off(uint32_t w) { return (w >> 14) & 0x3FFCu; }
    static __device__ __forceinline__ float ld(uint32_t a) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); return v; }
    static __device__ __forceinline__ void st(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
    static __device__ __forceinline__ float row_msg(uint32_t a1, uint32_t a2, float v) {
        float m;
        asm volatile("{\n .reg .pred p;\n .reg .f32 av, am;\n ld.shared.f32 %0, [%1];\n abs.f32 av, %3;\n abs.f32 am, %0;\n"
                     " setp.eq.f32 p, av, am;\n @p ld.shared.f32 %0, [%2];\n}"
                     : "=&f"(m) : "r"(a1), "r"(a2), "f"(v));
        return m;
    }
};

struct Ms2Ctx {
    uint32_t vbase, r1base, r2base;      // shared-window addresses of V, min1, min2
    uint32_t magic;                      // umulhi(byte offset of a slot, magic) & ~(sizeof(R)-1) = byte offset of its row in min1 / min2
};

// MODE 0: first iteration (every bit->check message is the column's prior, V is only written)   1: later iterations
template <typename R, int W, int MODE, bool UNIT>
__device__ __forceinline__ R column_ms2(const uint4 rec, const R l0, const R alpha, const Ms2Ctx& x) {
    using RT = Real<R>;
    using SH = Sh<R>;
    const uint32_t vo[6] = {SH::off_lo(rec.x), SH::off_hi(rec.x), SH::off_lo(rec.y), SH::off_hi(rec.y), SH::off_lo(rec.z), SH::off_hi(rec.z)};
    R c[W], vn[W];
    const R l0c = MODE == 0 ? canonical_sign<R>(l0) : l0;
#pragma unroll
    for (int q = 0; q < W; ++q) {
        const R v = MODE == 0 ? l0c : SH::ld(x.vbase + vo[q]);
        const uint32_t ro = __umulhi(vo[q], x.magic) & ~static_cast<uint32_t>(sizeof(R) - 1);
        R m = SH::row_msg(x.r1base + ro, x.r2base + ro, v);
        if (!UNIT) m = RT::mul(m, alpha);
        c[q] = Bits<R>::with_hi(m, Bits<R>::hi(m) ^ (Bits<R>::hi(v) & 0x80000000u));
    }
    R t = l0;
#pragma unroll
    for (int q = 0; q < W; ++q) { vn[q] = t; t = RT::add(t, c[q]); }
    const R llr = t;
    t = c[W - 1];
#pragma unroll
    for (int q = W - 2; q >= 0; --q) {
        vn[q] = RT::add(vn[q], t);
        if (q > 0) t = RT::add(t, c[q]);
    }
#pragma unroll
    for (int q = 0; q < W; ++q) SH::st(x.vbase + vo[q], canonical_sign<R>(vn[q]));
    return llr;
}

// one bit sweep: warp chunks of 32 records, heaviest first; chunk c is handled by warp c % NWARPS.
// CM 0: first iteration, 1: later iterations.  WRITE: posteriors are written out (last iteration, or the caller wants them every iteration)
// One bit sweep: warp chunks of 32 records, heaviest first; chunk c is handled by warp c % NWARPS.  Padding records carry a
// positive prior (an extra entry of the prior table) and dummy edges only, so their posterior is never <= 0.
template <typename R, int NWARPS, int CM, bool WRITE, bool UNIT>
__device__ __forceinline__ void sweep_ms2(const WinDev& w, const Ms2Ctx& x, const R alpha, const uint32_t ptab_s, uint32_t* ebits,
                                          uint32_t* cand, uint32_t* hist, R* llr_row, const bool last, const int warp, const int lane) {
    using SH = Sh<R>;
    constexpr int S = NWARPS * 32;
    int c = warp;
    const uint4* rp = w.colrec + warp * 32 + lane;          // colrec is padded by two chunks per warp past the last chunk
    uint4 ra = __ldg(rp), rb;
#define QB_CHUNK(WT, REC)                                                                                                \
    {                                                                                                                    \
        const R l0 = SH::ld(ptab_s + SH::ptab_off(REC.w));                                                               \
        const R llr = column_ms2<R, WT, CM, UNIT>(REC, l0, alpha, x);                                                    \
        if (llr <= R(0)) {                                    /* hard decision 1: record bit, candidate syndrome */      \
            atomicOr(&ebits[c], 1u << lane);                                                                             \
            const uint32_t e[6] = {REC.x & 0xFFFFu, REC.x >> 16, REC.y & 0xFFFFu, REC.y >> 16, REC.z & 0xFFFFu, REC.z >> 16}; \
            _Pragma("unroll") for (int q = 0; q < WT; ++q) {                                                             \
                const uint32_t row = __umulhi(e[q], w.rs_magic);                                                         \
                if (row < static_cast<uint32_t>(w.rows)) atomicXor(&cand[row >> 5], 1u << (row & 31u));                  \
            }                                                                                                            \
        }                                                                                                                \
        if (WRITE) {                                                                                                     \
            const uint32_t j = REC.w & 0xFFFFu;               /* original column; 0xFFFF marks a padding record */       \
            if (j != 0xFFFFu) {                                                                                          \
                llr_row[j] = llr;                                                                                        \
                if (last) atomicAdd(&hist[llr_bin<R>(llr, static_cast<R>(w.bin_scale))], 1u);                            \
            }                                                                                                            \
        }                                                                                                                \
    }
    // two records in flight (ra: chunk c, rb: chunk c + NWARPS) so that the prefetch needs no register copies in steady state
#define QB_SEGMENT(WT)                                                                                                   \
    while (c < w.chunk_end[WT]) {                                                                                        \
        rb = __ldg(rp + S);                                                                                              \
        QB_CHUNK(WT, ra)                                                                                                 \
        c += NWARPS; rp += S;                                                                                            \
        if (c >= w.chunk_end[WT]) { ra = rb; break; }                                                                    \
        ra = __ldg(rp + S);                                                                                              \
        QB_CHUNK(WT, rb)                                                                                                 \
        c += NWARPS; rp += S;                                                                                            \
    }
    QB_SEGMENT(6) QB_SEGMENT(5) QB_SEGMENT(4) QB_SEGMENT(3) QB_SEGMENT(2) QB_SEGMENT(1)
#undef QB_SEGMENT
#undef QB_CHUNK
    if (WRITE) {
        for (; c < w.nW32; c += NWARPS, rp += S) {            // weight-0 chunks: columns without a check (posterior = prior)
            const uint4 cur = __ldg(rp);
            const R l0 = SH::ld(ptab_s + SH::ptab_off(cur.w));
            const uint32_t j = cur.w & 0xFFFFu;
            if (j != 0xFFFFu) {
                if (l0 <= R(0)) atomicOr(&ebits[c], 1u << lane);
                llr_row[j] = l0;
                if (last) atomicAdd(&hist[llr_bin<R>(l0, static_cast<R>(w.bin_scale))], 1u);
            }
        }
    } else {
        for (; c < w.nW32; c += NWARPS, rp += S) {
            const uint4 cur = __ldg(rp);
            if ((cur.w & 0xFFFFu) != 0xFFFFu && SH::ld(ptab_s + SH::ptab_off(cur.w)) <= R(0)) atomicOr(&ebits[c], 1u << lane);
        }
    }
}

template <typename R, int NT, int MINB, bool UNIT>
__global__ void __launch_bounds__(NT, MINB) bp_kernel_ms2(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    using CT = Compact<R>;
    using BT = Bits<R>;
    constexpr int NWARPS = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off[10];
    bpm_layout(w, sizeof(R), off);
    R* rm1 = reinterpret_cast<R*>(smem_raw + off[0]);
    R* rm2 = reinterpret_cast<R*>(smem_raw + off[1]);
    R* V = reinterpret_cast<R*>(smem_raw + off[2]);
    uint32_t* syn = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
    uint32_t* cand = reinterpret_cast<uint32_t*>(smem_raw + off[4]);
    uint32_t* accs = reinterpret_cast<uint32_t*>(smem_raw + off[5]);
    uint32_t* car = reinterpret_cast<uint32_t*>(smem_raw + off[6]);
    R* ptab = reinterpret_cast<R*>(smem_raw + off[7]);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off[8]);
    uint32_t* ebits = reinterpret_cast<uint32_t*>(smem_raw + off[9]);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = w.rows, RS = w.RS;
    const int nchunks = w.nW32;
    Ms2Ctx x;
    x.r1base = static_cast<uint32_t>(__cvta_generic_to_shared(rm1));
    x.r2base = static_cast<uint32_t>(__cvta_generic_to_shared(rm2));
    x.vbase = static_cast<uint32_t>(__cvta_generic_to_shared(V));
    x.magic = w.rs_magic;
    const uint32_t ptab_s = static_cast<uint32_t>(__cvta_generic_to_shared(ptab));
    for (int i = tid; i < w.n_ptab; i += NT) ptab[i] = CT::ptab(w)[i];
    if (tid < bp_dummy_rows(w, sizeof(R))) { rm1[rows + tid] = R(0); rm2[rows + tid] = R(0); }
    V[rows * RS + tid] = R(0);

    for (int shot = blockIdx.x; shot < b.n_shots; shot += gridDim.x) {
        __syncthreads();
        load_syndrome(w, b, shot, tid, syn, accs, car);
        __syncthreads();
        R* const llr_row = reinterpret_cast<R*>(b.llr_buf) + static_cast<size_t>(shot) * b.llr_stride;
        bool conv = false;
        int it = 1;
        for (; it <= p.max_iter; ++it) {
            const R alpha = static_cast<R>(__ldg(p.alpha + it));
            const bool first = it == 1;
            const bool last = it == p.max_iter;
            // ---- check sweep: one thread per row -> min1[row], min2[row], both signed by the row parity
            for (int i = tid; i < rows; i += NT) {
                uint32_t par = ((syn[i >> 5] >> (i & 31)) & 1u) << 31;
                R m1, m2;
                if (first) {
                    const typename RT::pair s0 = CT::sum0(w, i);
                    m1 = s0.x; m2 = s0.y;
                    par ^= static_cast<uint32_t>(__ldg(w.neg0 + i)) << 31;
                } else {
                    const R* vr = V + i * RS;
                    const int len = __ldg(w.rlen + i);
                    R m1a = RT::big(), m2a = RT::big(), m1b = RT::big(), m2b = RT::big();
                    int q = 0;
                    for (; q + 7 < len; q += 8) {
#pragma unroll
                        for (int u = 0; u < 8; u += 2) {
                            min2_signed<R>(vr[q + u], m1a, m2a, par);
                            min2_signed<R>(vr[q + u + 1], m1b, m2b, par);
                        }
                    }
                    for (; q + 1 < len; q += 2) {
                        min2_signed<R>(vr[q], m1a, m2a, par);
                        min2_signed<R>(vr[q + 1], m1b, m2b, par);
                    }
                    if (q < len) min2_signed<R>(vr[q], m1a, m2a, par);
                    const bool lo = fabs(m1b) < fabs(m1a);
                    m1 = lo ? m1b : m1a;
                    const R mo = lo ? m1a : m1b;                                  // the larger of the two chain minima
                    const R m2c = fabs(m2b) < fabs(m2a) ? m2b : m2a;
                    m2 = fabs(m2c) < fabs(mo) ? m2c : mo;
                }
                par &= 0x80000000u;
                rm1[i] = BT::with_hi(m1, (BT::hi(m1) & 0x7FFFFFFFu) | par);
                rm2[i] = BT::with_hi(m2, (BT::hi(m2) & 0x7FFFFFFFu) | par);
            }
            // the candidate syndrome is double-buffered by iteration parity: every warp evaluates the stop test itself right after
            // the barrier that ends the bit sweep, and a fast warp may already be clearing the other buffer for the next iteration
            uint32_t* const cnd = cand + (it & 1) * w.rowsW32;
            for (int i = tid; i < w.rowsW32; i += NT) cnd[i] = 0;
            for (int i = tid; i < nchunks; i += NT) ebits[i] = 0;
            if (last && tid < 32) hist[tid] = 0;
            __syncthreads();
            // ---- bit sweep
            const bool wr = last || b.write_llr_always;
            if (first && !wr) sweep_ms2<R, NWARPS, 0, false, UNIT>(w, x, alpha, ptab_s, ebits, cnd, hist, llr_row, last, warp, lane);
            else if (first) sweep_ms2<R, NWARPS, 0, true, UNIT>(w, x, alpha, ptab_s, ebits, cnd, hist, llr_row, last, warp, lane);
            else if (!wr) sweep_ms2<R, NWARPS, 1, false, UNIT>(w, x, alpha, ptab_s, ebits, cnd, hist, llr_row, last, warp, lane);
            else sweep_ms2<R, NWARPS, 1, true, UNIT>(w, x, alpha, ptab_s, ebits, cnd, hist, llr_row, last, warp, lane);
            // ---- stop test H e == s, by every warp for itself (no second barrier)
            __syncthreads();
            int mismatch = 0;
            for (int i = lane; i < w.rowsW32; i += 32) mismatch |= cnd[i] != syn[i];
            if (!__any_sync(0xFFFFFFFFu, mismatch)) { conv = true; break; }
        }
        if (it > p.max_iter) it = p.max_iter;
        finish_shot<R, NT, true>(w, b, shot, tid, conv, it, 0u, syn, accs, car, hist, ebits);
    }
}

// =====================================================================================================================
// Serial schedule (ldpc schedule='serial', the reference wrappers' default, decoder/bposd.py:54): the columns are updated one
// after the other in index order, each from the CURRENT messages of its rows (oracle/bp_impl.inc, serial branch).  Columns
// that share no row commute, so the host cuts the column sequence into dependency levels (api.cu, serial tables) and the
// kernel walks "steps" of up to NT/8 independent (column, row) pairs: eight lanes scan one row for the minimum / parity /
// tanh product over the OTHER edges, the pair leader forms the check->bit message, then one thread per column does the
// prefix / suffix sums in the oracle's order and writes the column's new messages.  Two block barriers per step, ~1100 steps
// per iteration on the gross-code windows: latency bound, an order of magnitude slower than flooding -- as it is on the CPU.
// =====================================================================================================================
constexpr int kSerialThreads = 128;
constexpr int kSerialPairs = kSerialThreads / 8;

// off: V, syn, cand, accs, car, ptab, hist, ebits, cbuf, pbuf
__host__ __device__ inline size_t bps_layout(const WinDev& w, int rsize, size_t* off /*[10]*/) {
    size_t o = 0;
    off[0] = o; o += align_up((static_cast<size_t>(w.rows) * w.RS + 1) * rsize, 16);
    off[1] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[2] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[3] = o; o += align_up(static_cast<size_t>(w.KW) * 8, 16);
    off[4] = o; o += align_up(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4, 16);
    off[5] = o; o += align_up(static_cast<size_t>(w.n_ptab) * rsize, 16);
    off[6] = o; o += kSelWords * 4;
    off[7] = o; o += align_up(static_cast<size_t>(w.nW32) * 4, 16);
    off[8] = o; o += align_up(static_cast<size_t>(kSerialPairs) * rsize, 16);
    off[9] = o; o += align_up(static_cast<size_t>(kSerialPairs) * 4, 16);
    return o;
}

template <typename R, bool PS>
__global__ void __launch_bounds__(kSerialThreads) bp_kernel_serial(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    using CT = Compact<R>;
    constexpr int NT = kSerialThreads;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off[10];
    bps_layout(w, sizeof(R), off);
    R* V = reinterpret_cast<R*>(smem_raw + off[0]);
    uint32_t* syn = reinterpret_cast<uint32_t*>(smem_raw + off[1]);
    uint32_t* cand = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
    uint32_t* accs = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
    uint32_t* car = reinterpret_cast<uint32_t*>(smem_raw + off[4]);
    R* ptab = reinterpret_cast<R*>(smem_raw + off[5]);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off[6]);
    uint32_t* ebits = reinterpret_cast<uint32_t*>(smem_raw + off[7]);
    R* cbuf = reinterpret_cast<R*>(smem_raw + off[8]);
    uint32_t* pbuf = reinterpret_cast<uint32_t*>(smem_raw + off[9]);

    const int tid = threadIdx.x, grp = tid >> 3, l8 = tid & 7;
    const int rows = w.rows, RS = w.RS, npad = w.ncols_pad;
    R* const llr_all = reinterpret_cast<R*>(b.llr_buf);
    for (int i = tid; i < w.n_ptab; i += NT) ptab[i] = CT::ptab(w)[i];

    for (int shot = blockIdx.x; shot < b.n_shots; shot += gridDim.x) {
        __syncthreads();
        load_syndrome(w, b, shot, tid, syn, accs, car);
        for (int r = tid; r < npad; r += NT) {                  // every message starts at its column's prior
            const uint4 rec = __ldg(w.colrec + r);
            const R l0 = ptab[(rec.w >> 16) & 0xFFFu];
            const R m0 = PS ? Trans<R>::th(RT::mul(l0, R(0.5))) : l0;        // product-sum keeps tanh(v/2) in the message array
            const uint32_t e[6] = {rec.x & 0xFFFFu, rec.x >> 16, rec.y & 0xFFFFu, rec.y >> 16, rec.z & 0xFFFFu, rec.z >> 16};
#pragma unroll
            for (int q = 0; q < 6; ++q)
                if (e[q] < static_cast<uint32_t>(rows * RS)) V[e[q]] = m0;
        }
        __syncthreads();
        bool conv = false;
        int it = 1;
        for (; it <= p.max_iter; ++it) {
            const R alpha = static_cast<R>(__ldg(p.alpha + it));
            const bool last = it == p.max_iter;
            for (int i = tid; i < w.rowsW32; i += NT) cand[i] = 0;
            for (int i = tid; i < w.nW32; i += NT) ebits[i] = 0;
            if (last && tid < 32) hist[tid] = 0;
            __syncthreads();
            // records of step 0; inside the loop the records of step s+1 are fetched before the barriers of step s
            uint32_t hdr = __ldg(w.ser_steps);
            uint2 pr = __ldg(w.ser_pairs + grp);
            uint2 cr = tid < kSerialPairs ? __ldg(w.ser_cols + tid) : make_uint2(0u, 0u);
            for (int s = 0; s < w.ser_nsteps; ++s) {
                const int n_pairs = static_cast<int>(hdr & 0xFFu), n_cols = static_cast<int>(hdr >> 8);
                const uint2 mypair = pr, mycol = cr;
                hdr = __ldg(w.ser_steps + s + 1);
                pr = __ldg(w.ser_pairs + static_cast<size_t>(s + 1) * kSerialPairs + grp);
                if (tid < kSerialPairs) cr = __ldg(w.ser_cols + static_cast<size_t>(s + 1) * kSerialPairs + tid);
                // ---- eight lanes per (column, row) pair: message of the row to the column from the row's other edges
                {
                    const bool act = grp < n_pairs;                        // idle groups run the shuffles too (full-warp mask)
                    const int addr = static_cast<int>(mypair.x & 0xFFFFu), row = static_cast<int>(mypair.x >> 16);
                    const int base = row * RS, own = addr - base, len = act ? static_cast<int>(mypair.y) : 0;
                    R acc = PS ? R(1) : RT::big();
                    uint32_t neg = 0;
                    for (int k = l8; k < len; k += 8) {
                        if (k == own) continue;
                        const R v = V[base + k];
                        if (PS) {
                            acc = RT::mul(acc, v);
                        } else {
                            const R a = CT::mag(v);
                            acc = a < acc ? a : acc;
                            neg += v <= R(0) ? 1u : 0u;
                        }
                    }
#pragma unroll
                    for (int o = 4; o > 0; o >>= 1) {
                        const R other = __shfl_xor_sync(0xFFFFFFFFu, acc, o, 8);
                        acc = PS ? RT::mul(acc, other) : (other < acc ? other : acc);
                        neg += __shfl_xor_sync(0xFFFFFFFFu, neg, o, 8);
                    }
                    if (act && l8 == 0) {
                        const uint32_t sbit = (syn[row >> 5] >> (row & 31)) & 1u;
                        R c;
                        if (PS) {
                            const R x = Trans<R>::lg(Trans<R>::div(RT::add(R(1), acc), RT::add(R(1), -acc)));
                            c = sbit ? -x : x;
                        } else {
                            c = RT::mul(acc, ((sbit + neg) & 1u) ? -alpha : alpha);
                        }
                        cbuf[grp] = c;
                        pbuf[grp] = mypair.x;
                    }
                }
                __syncthreads();
                // ---- one thread per column of the step: prefix / suffix sums in the oracle's order, new messages, hard decision
                if (tid < n_cols) {
                    const int j = static_cast<int>(mycol.x & 0xFFFFu), first = static_cast<int>(mycol.y & 0xFFu), wt = static_cast<int>((mycol.y >> 8) & 0xFFu);
                    const R l0 = ptab[(mycol.x >> 16) & 0xFFFu];
                    R c[6], vn[6];
                    uint32_t pa[6];
#pragma unroll
                    for (int q = 0; q < 6; ++q) {
                        c[q] = q < wt ? cbuf[first + q] : R(0);
                        pa[q] = q < wt ? pbuf[first + q] : 0u;
                    }
                    R t = l0;
#pragma unroll
                    for (int q = 0; q < 6; ++q) if (q < wt) { vn[q] = t; t = RT::add(t, c[q]); }
                    const R llr = t;
                    t = R(0);
#pragma unroll
                    for (int q = 5; q >= 0; --q) if (q < wt) { vn[q] = RT::add(vn[q], t); t = RT::add(t, c[q]); }
#pragma unroll
                    for (int q = 0; q < 6; ++q) if (q < wt) V[pa[q] & 0xFFFFu] = PS ? Trans<R>::th(RT::mul(vn[q], R(0.5))) : vn[q];
                    if (llr <= R(0)) {
                        atomicOr(&ebits[j >> 5], 1u << (j & 31));
#pragma unroll
                        for (int q = 0; q < 6; ++q) if (q < wt) atomicXor(&cand[(pa[q] >> 16) >> 5], 1u << ((pa[q] >> 16) & 31u));
                    }
                    if (last || b.write_llr_always) llr_all[static_cast<size_t>(shot) * b.llr_stride + j] = llr;
                    if (last) atomicAdd(&hist[llr_bin<R>(llr, static_cast<R>(w.bin_scale))], 1u);
                }
                __syncthreads();
            }
            // ---- stop test H e == s (after the full sweep, as the oracle does)
            int mismatch = 0;
            for (int i = tid; i < w.rowsW32; i += NT) mismatch |= cand[i] != syn[i];
            if (!__syncthreads_or(mismatch)) { conv = true; break; }
        }
        if (it > p.max_iter) it = p.max_iter;
        finish_shot<R, NT, false>(w, b, shot, tid, conv, it, 0u, syn, accs, car, hist, ebits);
    }
}

// =====================================================================================================================
// Serial schedule, second form: ONE WARP per shot, one LANE per edge, incrementally maintained row summaries
// (bp_kernel_serial_warp).  The oracle's serial sweep forms the message of row i to column j from the row's OTHER current
// messages -- a scan of ~35 entries per edge in the kernel above.  Here every row keeps its summary up to date as the sweep goes:
//   min-sum      (min1 with the parity of the syndrome bit and of #{v <= 0} in its sign bit, min2): "the others" is min2 when the
//                edge's own |v| equals min1, else min1 (exact: a minimum does not depend on the order it is taken in), and after the
//                column's new messages are written the summary is patched in O(1) -- except when an edge that held min1 or
//                min2 grows past min2: then the third smallest is needed and the warp re-scans that row together (32 lanes,
//                redux.sync minima on the order-preserving bit patterns)
//   product-sum  no summary: dividing the own factor out of a running row product drifts and, near saturation (tanh(v/2) == 1,
//                x -> 1), turns a last-ulp excess into log(negative) = NaN that an incremental product never forgets; each lane
//                multiplies the other factors of its row itself, in two chains (the tolerance path of the kernel above)
// so a min-sum edge costs O(1) instead of O(row length), and there is no block barrier in either variant.  A step is up to five independent columns
// of one dependency level; lanes 6g .. 6g+5 hold the (at most six) edges of column g, exchange their check->bit messages by
// shuffle, and each forms the column's prefix / suffix sums in the oracle's order; one 128-byte record row per step, the rows
// of the next four steps already in registers.  ~1140 steps x a few hundred cycles per iteration and shot-window; the number
// of concurrent shots is still set by the message array (two per SM in fp64), so the kernel stays latency bound -- on a
// several times shorter chain.
// =====================================================================================================================
constexpr int kSerCols = 5;          // columns per step
constexpr int kSerEdges = 6;         // lanes per column (the compact layout holds columns of weight <= 6)

// off: V, rsum, syn, cand, accs, car, ptab, hist, ebits
__host__ __device__ inline size_t bpsw_layout(const WinDev& w, int rsize, size_t* off /*[9]*/) {
    size_t o = 0;
    off[0] = o; o += align_up((static_cast<size_t>(w.rows) * w.RS + 32) * rsize, 16);          // + one private dummy slot per lane
    off[1] = o; o += align_up((static_cast<size_t>(w.rows) + 34) * 2 * rsize, 16);              // + the rows the dummy slots divide down to
    off[2] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[3] = o; o += align_up(static_cast<size_t>(w.rowsW32) * 4, 16);
    off[4] = o; o += align_up(static_cast<size_t>(w.KW) * 8, 16);
    off[5] = o; o += align_up(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4, 16);
    off[6] = o; o += align_up(static_cast<size_t>(w.n_ptab) * rsize, 16);
    off[7] = o; o += kSelWords * 4;
    off[8] = o; o += align_up(static_cast<size_t>(w.nW32) * 4, 16);
    return o;
}

// warp minimum of non-negative reals through their bit patterns (monotone for x >= 0)
__device__ __forceinline__ float warp_min_mag(float x) {
    return __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, __float_as_uint(x)));
}
__device__ __forceinline__ double warp_min_mag(double x) {
    const uint32_t hi = static_cast<uint32_t>(__double2hiint(x)), lo = static_cast<uint32_t>(__double2loint(x));
    const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
    const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, hi == mh ? lo : 0xFFFFFFFFu);
    return __hiloint2double(static_cast<int>(mh), static_cast<int>(ml));
}

template <typename R> __device__ __forceinline__ uint32_t sign_bit(R x);
template <> __device__ __forceinline__ uint32_t sign_bit<float>(float x) { return __float_as_uint(x) >> 31; }
template <> __device__ __forceinline__ uint32_t sign_bit<double>(double x) { return static_cast<uint32_t>(__double2hiint(x)) >> 31; }

// the two smallest |v| of a row, by the whole warp; the parity already in the summary's sign bit is kept
template <typename R>
__device__ __noinline__ void serial_rescan_row(const R* vr, const int len, typename Real<R>::pair* srow, const int lane) {
    using RT = Real<R>;
    using CT = Compact<R>;
    R l1 = RT::big(), l2 = RT::big();
    for (int k = lane; k < len; k += 32) {
        const R a = CT::mag(vr[k]);
        const bool p = a < l1, q = a < l2;
        l2 = p ? l1 : (q ? a : l2);
        l1 = p ? a : l1;
    }
    const R m1 = warp_min_mag(l1);
    const R m2 = warp_min_mag(l1 == m1 ? l2 : l1);
    const uint32_t holders = __ballot_sync(0xFFFFFFFFu, l1 == m1);
    if (lane == 0) *srow = RT::mk(CT::signed_by(m1, sign_bit<R>(srow->x)), __popc(holders) >= 2 ? m1 : m2);
    __syncwarp();
}

template <typename R> __device__ __forceinline__ R shfl_real(R x, int src);
template <> __device__ __forceinline__ float shfl_real<float>(float x, int src) { return __shfl_sync(0xFFFFFFFFu, x, src); }
template <> __device__ __forceinline__ double shfl_real<double>(double x, int src) { return __shfl_sync(0xFFFFFFFFu, x, src); }

template <typename R, bool PS>
__global__ void __launch_bounds__(32) bp_kernel_serial_warp(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    using CT = Compact<R>;
    using TT = Trans<R>;
    using Pair = typename RT::pair;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t off[9];
    bpsw_layout(w, sizeof(R), off);
    R* V = reinterpret_cast<R*>(smem_raw + off[0]);
    Pair* rsum = reinterpret_cast<Pair*>(smem_raw + off[1]);
    uint32_t* syn = reinterpret_cast<uint32_t*>(smem_raw + off[2]);
    uint32_t* cand = reinterpret_cast<uint32_t*>(smem_raw + off[3]);
    uint32_t* accs = reinterpret_cast<uint32_t*>(smem_raw + off[4]);
    uint32_t* car = reinterpret_cast<uint32_t*>(smem_raw + off[5]);
    R* ptab = reinterpret_cast<R*>(smem_raw + off[6]);
    uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw + off[7]);
    uint32_t* ebits = reinterpret_cast<uint32_t*>(smem_raw + off[8]);

    const int lane = threadIdx.x;
    const int grp = lane / kSerEdges, q = lane - grp * kSerEdges, gbase = grp * kSerEdges;      // lanes 30, 31: group 5, never active
    const int rows = w.rows, RS = w.RS, npad = w.ncols_pad;
    const uint32_t realN = static_cast<uint32_t>(rows) * static_cast<uint32_t>(RS);
    const uint32_t magic = w.rs_magic;
    R* const llr_all = reinterpret_cast<R*>(b.llr_buf);
    for (int i = lane; i < w.n_ptab; i += 32) ptab[i] = CT::ptab(w)[i];
    V[realN + lane] = R(0);
    for (int i = rows + lane; i < rows + 34; i += 32) rsum[i] = RT::mk(R(0), PS ? R(2) : R(0));
    const uint32_t* tab = w.ser32_rec;

    for (int shot = blockIdx.x; shot < b.n_shots; shot += gridDim.x) {
        __syncthreads();
        load_syndrome(w, b, shot, lane, syn, accs, car);
        // bit -> check messages start at the priors (product-sum keeps tanh(v/2))
        for (int r = lane; r < npad; r += 32) {
            const uint4 rec = __ldg(w.colrec + r);
            const R l0 = ptab[(rec.w >> 16) & 0xFFFu];
            const R v0 = PS ? TT::th(RT::mul(l0, R(0.5))) : l0;
            const uint32_t e[6] = {rec.x & 0xFFFFu, rec.x >> 16, rec.y & 0xFFFFu, rec.y >> 16, rec.z & 0xFFFFu, rec.z >> 16};
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (e[k] < realN) V[e[k]] = v0;
        }
        __syncthreads();
        for (int i = lane; i < rows; i += 32) {
            const uint32_t sbit = (syn[i >> 5] >> (i & 31)) & 1u;
            if (!PS) {
                const Pair s0 = CT::sum0(w, i);
                rsum[i] = RT::mk(CT::signed_by(s0.x, (sbit + __ldg(w.neg0 + i)) & 1u), s0.y);
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
            // record rows of the next four steps in registers (the table is padded by four rows)
            uint32_t r0 = __ldg(tab + lane), r1 = __ldg(tab + 32 + lane), r2 = __ldg(tab + 64 + lane), r3 = __ldg(tab + 96 + lane);
            const int ns = w.ser32_nsteps;
#pragma unroll 1
            for (int s = 0; s < ns; ++s) {
                const uint32_t rec = r0;
                r0 = r1; r1 = r2; r2 = r3;
                r3 = __ldg(tab + static_cast<size_t>(s + 4) * 32 + lane);
                // lane word: message address | (q == 0: column, q == 1: prior index) << 16
                const uint32_t e = rec & 0xFFFFu;
                const uint32_t j = __shfl_sync(0xFFFFFFFFu, rec, gbase) >> 16;
                const uint32_t pi = (__shfl_sync(0xFFFFFFFFu, rec, gbase + 1) >> 16) & 0xFFFu;
                const bool active = j != 0xFFFFu && grp < kSerCols;
                const bool real = active && e < realN;
                const uint32_t row = __umulhi(e, magic);
                const R vold = V[e];
                const Pair sm = rsum[row];
                R c;
                if (PS) {
                    c = R(0);
                    if (real) {
                        const R* vr = V + row * RS;
                        const int len = __ldg(w.rlen + row), own = static_cast<int>(e) - static_cast<int>(row) * RS;
                        R pa = R(1), pb = R(1);
                        int k = 0;
                        for (; k + 1 < len; k += 2) {
                            const R ta = vr[k], tb = vr[k + 1];
                            pa = k == own ? pa : RT::mul(pa, ta);
                            pb = k + 1 == own ? pb : RT::mul(pb, tb);
                        }
                        if (k < len && k != own) pa = RT::mul(pa, vr[k]);
                        const R x = RT::mul(pa, pb);
                        const R lx = TT::lg(TT::div(RT::add(R(1), x), RT::add(R(1), -x)));
                        c = ((syn[row >> 5] >> (row & 31u)) & 1u) ? -lx : lx;
                    }
                } else {
                    const R m1 = CT::mag(sm.x);
                    const R m = CT::mag(vold) == m1 ? sm.y : m1;
                    c = CT::flip(RT::mul(m, alpha), sm.x, vold <= R(0));
                }
                // the column's sums in the oracle's order: v_q = (l0 + c_0 + .. + c_{q-1}) + (c_5 + .. + c_{q+1}); dummy edges add +-0
                R ck[kSerEdges];
#pragma unroll
                for (int k = 0; k < kSerEdges; ++k) ck[k] = shfl_real<R>(c, gbase + k);
                R t = ptab[pi], pre = R(0), suf = R(0);
#pragma unroll
                for (int k = 0; k < kSerEdges; ++k) { pre = k == q ? t : pre; t = RT::add(t, ck[k]); }
                const R llr = t;
                t = R(0);
#pragma unroll
                for (int k = kSerEdges - 1; k >= 0; --k) { suf = k == q ? t : suf; t = RT::add(t, ck[k]); }
                const R vn = RT::add(pre, suf);
                bool rescan = false;
                if (real) {
                    if (PS) {
                        V[e] = TT::th(RT::mul(vn, R(0.5)));
                    } else {
                        V[e] = vn;
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
                    uint32_t mask = __ballot_sync(0xFFFFFFFFu, rescan);
                    while (mask) {
                        const int src = __ffs(mask) - 1;
                        mask &= mask - 1;
                        const uint32_t rr = __shfl_sync(0xFFFFFFFFu, row, src);
                        serial_rescan_row<R>(V + rr * RS, __ldg(w.rlen + rr), rsum + rr, lane);
                    }
                }
            }
            // ---- stop test H e == s (after the full sweep, as the oracle does)
            __syncthreads();
            int mismatch = 0;
            for (int i = lane; i < w.rowsW32; i += 32) mismatch |= cand[i] != syn[i];
            if (!__syncthreads_or(mismatch)) { conv = true; break; }
        }
        if (it > p.max_iter) it = p.max_iter;
        finish_shot<R, 32, false>(w, b, shot, lane, conv, it, 0u, syn, accs, car, hist, ebits);
    }
}

// kernel variants: (precision, column-weight template, V in shared or global)
using KernelPtr = void (*)(const WinDev, const BatchDev, const BpParams);

struct Variant {
    KernelPtr fn;
    int threads;
    size_t configured[kMaxDevices];
};

template <typename R, int CW, bool VG>
KernelPtr pick_kernel(int* threads) {
    // fp32: 256 threads x 4 CTAs/SM (64 regs); fp64: 512 threads x 2 CTAs/SM; wide columns get more registers
    constexpr int NT = sizeof(R) == 4 ? 256 : 512;
    constexpr int MINB = CW <= 8 ? (sizeof(R) == 4 ? 4 : 2) : 1;
    *threads = NT;
    return bp_kernel<R, CW, NT, MINB, VG>;
}

Variant& variant(int prec, int cw, bool vg) {
    static Variant table[2][3][2] = {};
    const int pi = prec == 32 ? 0 : 1, ci = cw <= 6 ? 0 : (cw <= 8 ? 1 : 2), gi = vg ? 1 : 0;
    Variant& v = table[pi][ci][gi];
    if (!v.fn) {
#define QB_PICK(R, CWV) (vg ? pick_kernel<R, CWV, true>(&v.threads) : pick_kernel<R, CWV, false>(&v.threads))
        if (pi == 0) v.fn = ci == 0 ? QB_PICK(float, 6) : (ci == 1 ? QB_PICK(float, 8) : QB_PICK(float, 16));
        else v.fn = ci == 0 ? QB_PICK(double, 6) : (ci == 1 ? QB_PICK(double, 8) : QB_PICK(double, 16));
#undef QB_PICK
    }
    return v;
}

Variant& compact_variant(int prec, int method) {
    static Variant table[2][2] = {};
    Variant& v = table[prec == 32 ? 0 : 1][method ? 1 : 0];
    if (!v.fn) {
        if (prec == 32) { v.fn = method ? bp_kernel_compact<float, 256, 4, true> : bp_kernel_compact<float, 256, 4, false>; v.threads = 256; }
        else { v.fn = method ? bp_kernel_compact<double, 512, 2, true> : bp_kernel_compact<double, 512, 2, false>; v.threads = 512; }
    }
    return v;
}

inline bool use_compact(const WinDev& w, bool vglobal) { return w.compact && !vglobal; }

// flooding min-sum on the compact layout: second form unless QB_BP_MS2=0 (A/B measurements)
inline bool ms2_enabled() {
    static const int on = [] { const char* e = getenv("QB_BP_MS2"); return e ? atoi(e) : 1; }();
    return on != 0;
}
inline bool use_ms2(const WinDev& w, bool vglobal, int method) { return use_compact(w, vglobal) && method == 0 && ms2_enabled(); }

Variant& ms2_variant(int prec, bool unit) {
    static Variant table[2][2] = {};
    Variant& v = table[prec == 32 ? 0 : 1][unit ? 1 : 0];
    if (!v.fn) {
        if (prec == 32) { v.fn = unit ? bp_kernel_ms2<float, 256, 4, true> : bp_kernel_ms2<float, 256, 4, false>; v.threads = 256; }
        else { v.fn = unit ? bp_kernel_ms2<double, 512, 2, true> : bp_kernel_ms2<double, 512, 2, false>; v.threads = 512; }
    }
    return v;
}

}  // namespace

// the warp-per-shot form needs everything its single warp indexes by lane to fit 32 lanes
static bool serial_warp_ok(const WinDev& w) {
    return w.ser32_rec != nullptr && w.rowsW32 <= 32 && 2 * w.KW <= 32 && (w.carry_rows + 31) / 32 + 1 <= 32;
}

size_t bp_serial_smem_bytes(const WinDev& w, int precision) {
    size_t off[10];
    if (serial_warp_ok(w)) return bpsw_layout(w, precision == 32 ? 4 : 8, off);
    return bps_layout(w, precision == 32 ? 4 : 8, off);
}

static KernelPtr serial_kernel(int precision, int method, bool warp) {
    if (warp) {
        if (precision == 32) return method ? bp_kernel_serial_warp<float, true> : bp_kernel_serial_warp<float, false>;
        return method ? bp_kernel_serial_warp<double, true> : bp_kernel_serial_warp<double, false>;
    }
    if (precision == 32) return method ? bp_kernel_serial<float, true> : bp_kernel_serial<float, false>;
    return method ? bp_kernel_serial<double, true> : bp_kernel_serial<double, false>;
}

cudaError_t bp_serial_configure(const WinDev& w, int precision, int method) {
    if (!w.compact || !w.ser_steps) return cudaErrorInvalidValue;
    const size_t smem = bp_serial_smem_bytes(w, precision);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    static size_t have[kMaxDevices][2][2][2] = {};
    const bool warp = serial_warp_ok(w);
    size_t& h = have[device_slot()][precision == 32 ? 0 : 1][method ? 1 : 0][warp ? 1 : 0];
    if (smem <= h) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(serial_kernel(precision, method, warp), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) h = smem;
    return e;
}

cudaError_t launch_bp_serial(const WinDev& w, const BatchDev& b, const BpParams& p, int precision, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const bool warp = serial_warp_ok(w);
    serial_kernel(precision, p.method, warp)<<<grid, warp ? 32 : kSerialThreads, bp_serial_smem_bytes(w, precision), st>>>(w, b, p);
    return cudaGetLastError();
}

size_t bp_smem_bytes(const WinDev& w, int precision, bool vglobal) {
    size_t off[10];
    if (use_compact(w, vglobal)) return std::max(bpc_layout(w, precision == 32 ? 4 : 8, off), bpm_layout(w, precision == 32 ? 4 : 8, off));
    return bp_layout(w, precision == 32 ? 4 : 8, vglobal, off);
}

int bp_threads(int precision) { return precision == 32 ? 256 : 512; }

bool bp_ms2_enabled() { return ms2_enabled(); }

bool bp_supports(const WinDev& w, int method, bool vglobal) { return method == 0 || use_compact(w, vglobal); }

cudaError_t bp_configure(const WinDev& w, int precision, bool vglobal, int method) {
    if (w.cw > 16) return cudaErrorInvalidValue;
    if (!bp_supports(w, method, vglobal)) return cudaErrorInvalidValue;
    const size_t smem = bp_smem_bytes(w, precision, vglobal);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    Variant& v = use_ms2(w, vglobal, method) ? ms2_variant(precision, w.unit_alpha != 0)
                 : (use_compact(w, vglobal) ? compact_variant(precision, method) : variant(precision, w.cw, vglobal));
    size_t& have = v.configured[device_slot()];
    if (smem <= have) return cudaSuccess;                   // the attribute only ever grows (decoders of different sizes coexist)
    cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) have = smem;
    return e;
}

cudaError_t launch_bp(const WinDev& w, const BatchDev& b, const BpParams& p, int precision, bool vglobal, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    Variant& v = use_ms2(w, vglobal, p.method) ? ms2_variant(precision, w.unit_alpha != 0)
                 : (use_compact(w, vglobal) ? compact_variant(precision, p.method) : variant(precision, w.cw, vglobal));
    const size_t smem = bp_smem_bytes(w, precision, vglobal);
    v.fn<<<grid, v.threads, smem, st>>>(w, b, p);
    return cudaGetLastError();
}

}  // namespace qb
