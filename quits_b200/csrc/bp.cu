// quits_b200/csrc/bp.cu -- K3 (+K2/K5 fused): flooding BP of one sliding window, one shot per CTA (sm_100a).
//
// Three kernels, one arithmetic:
//   bp_kernel_ms2        flooding min-sum on the compact layout (column weight <= 6, messages in shared memory): the headline path
//   bp_kernel_compact    the same layout, kept for product-sum (and as round 1's min-sum kernel for A/B runs, QB_BP_MS2=0)
//   bp_kernel            any window (column weight <= 16; messages in shared memory or, VGLOBAL, in a global slab per CTA), both methods
// (the serial schedule is bp_serial.cu).
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

#include "bp_common.cuh"
#include "qb_device.h"

namespace qb {

namespace {

// ---- bulk-asynchronous staging (TMA engine, 1-D form): cp.async.bulk global -> shared, completion on an mbarrier.
// Used by the global-slab flooding kernel: its check sweep walks every row of the message slab, one thread per row -- 32 lanes
// touching 32 different 32-byte sectors per load, four times the bytes the rows hold.  Staging tiles of whole rows through the
// copy engine reads each byte once, in 128-byte lines, while the threads work on the previous tile.
constexpr int kTileRows = 32;                   // rows per staged tile (16 lanes per row with 512 threads)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();        // a lost copy must not hang the device
    }
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// shared-memory layout; returns the total.  off: V, rsum, rmeta, syn, cand, accs, car, hist, ebits (hard decisions as a bit array:
// windows wider than 32 columns per thread, where the per-thread mask runs out)
__host__ __device__ inline size_t bp_tile_bytes(const WinDev& w, int rsize) { return static_cast<size_t>(kTileRows) * w.RS * rsize; }
__host__ __device__ inline size_t bp_slab_elems(const WinDev& w) {
    return static_cast<size_t>((w.rows + kTileRows - 1) / kTileRows * kTileRows) * w.RS;
}
__host__ __device__ inline size_t bp_layout(const WinDev& w, int rsize, bool vglobal, size_t* off /*[11]*/, bool staged = false) {
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
    off[9] = o; o += (vglobal && staged) ? 2 * align_up(bp_tile_bytes(w, rsize), 128) : 0;      // two row tiles in flight
    off[10] = o; o += (vglobal && staged) ? 16 : 0;                                               // their mbarriers
    return o;
}

// PS: product-sum (the message array holds tanh(v/2); the row summary is (signed product of the non-zero factors, number of
// zero factors) and "the others" is the row product divided by the edge's own factor -- as in bp_kernel_compact<.., true>)
template <typename R, int CW, int NT, int MINB, bool VGLOBAL, bool PS = false, bool STAGED = false>
__global__ void __launch_bounds__(NT, MINB) bp_kernel(const WinDev w, const BatchDev b, const BpParams p) {
    using RT = Real<R>;
    using TT = Trans<R>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    size_t off[11];
    bp_layout(w, sizeof(R), VGLOBAL, off, STAGED);
    // the slab of a CTA holds whole tiles of kTileRows rows (the staged check sweep copies the last tile in full)
    R* V = VGLOBAL ? reinterpret_cast<R*>(b.vscratch) + static_cast<size_t>(blockIdx.x) * bp_slab_elems(w)
                   : reinterpret_cast<R*>(smem_raw + off[0]);
    R* const tile = reinterpret_cast<R*>(smem_raw + off[9]);
    uint64_t* const tbar = reinterpret_cast<uint64_t*>(smem_raw + off[10]);
    const uint32_t tile_bytes = static_cast<uint32_t>(bp_tile_bytes(w, sizeof(R)));
    const size_t tile_stride = align_up(tile_bytes, 128) / sizeof(R);
    uint32_t tphase0 = 0, tphase1 = 0;           // parity of the next completion of each tile barrier
    if (STAGED && VGLOBAL) {
        if (threadIdx.x == 0) { mbar_init(&tbar[0], 1); mbar_init(&tbar[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
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
        for (int i = tid; i < rows * RS; i += NT) V[i] = PS ? R(1) : RT::big();       // padding slots: never the minimum, never negative (product-sum: a factor 1)
        __syncthreads();
        for (int j = tid; j < ncols; j += NT) {
            const R l0 = PS ? TT::th(RT::mul(RT::prior(w, j), R(0.5))) : RT::prior(w, j);
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
            // ---- check sweep
            if (STAGED && VGLOBAL) {
                // tiles of kTileRows whole rows come in through the copy engine (two in flight); NT / kTileRows lanes share a row
                constexpr int LPR = NT / kTileRows;
                const int ntiles = (rows + kTileRows - 1) / kTileRows;
                const int tr = tid / LPR, tl = tid - tr * LPR;
                fence_proxy_async();                                  // this CTA's message stores (bit sweep) before the engine reads them
                __syncthreads();
                if (tid == 0) {
                    mbar_expect_tx(&tbar[0], tile_bytes);
                    bulk_load(tile, V, tile_bytes, &tbar[0]);
                }
                for (int t = 0; t < ntiles; ++t) {
                    const int cur = t & 1;
                    if (tid == 0 && t + 1 < ntiles) {                 // next tile into the other buffer (released by the barrier below)
                        mbar_expect_tx(&tbar[cur ^ 1], tile_bytes);
                        bulk_load(tile + (cur ^ 1) * tile_stride, V + static_cast<size_t>(t + 1) * kTileRows * RS, tile_bytes, &tbar[cur ^ 1]);
                    }
                    mbar_wait(&tbar[cur], cur ? tphase1 : tphase0);
                    if (cur) tphase1 ^= 1u; else tphase0 ^= 1u;
                    const int i = t * kTileRows + tr;
                    const R* vr = tile + cur * tile_stride + tr * RS;
                    if (PS) {
                        R prod = R(1);
                        int zc = 0;
                        for (int sl = tl; sl < RS; sl += LPR) {
                            const R f = vr[sl];
                            if (f == R(0)) ++zc;
                            else prod = RT::mul(prod, f);
                        }
#pragma unroll
                        for (int o = LPR / 2; o > 0; o >>= 1) {
                            prod = RT::mul(prod, __shfl_xor_sync(0xFFFFFFFFu, prod, o, LPR));
                            zc += __shfl_xor_sync(0xFFFFFFFFu, zc, o, LPR);
                        }
                        if (tl == 0 && i < rows) {
                            const uint32_t sb = (syn[i >> 5] >> (i & 31)) & 1u;
                            rsum[i] = RT::mk(sb ? -prod : prod, static_cast<R>(zc));
                        }
                    } else {
                        R m1 = RT::big(), m2 = RT::big();
                        uint32_t arg = 0, neg = 0;
                        for (int sl = tl; sl < RS; sl += LPR) {
                            const R v = vr[sl];
                            const R a = RT::abs(v);
                            neg += v <= R(0) ? 1u : 0u;
                            if (a < m1) { m2 = m1; m1 = a; arg = static_cast<uint32_t>(sl); }
                            else if (a < m2) { m2 = a; }
                        }
#pragma unroll
                        for (int o = LPR / 2; o > 0; o >>= 1) {
                            const R o1 = __shfl_xor_sync(0xFFFFFFFFu, m1, o, LPR), o2 = __shfl_xor_sync(0xFFFFFFFFu, m2, o, LPR);
                            const uint32_t oa = __shfl_xor_sync(0xFFFFFFFFu, arg, o, LPR);
                            neg += __shfl_xor_sync(0xFFFFFFFFu, neg, o, LPR);
                            if (o1 < m1) { m2 = m1 < o2 ? m1 : o2; m1 = o1; arg = oa; }
                            else { m2 = o1 < m2 ? o1 : m2; }
                        }
                        if (tl == 0 && i < rows) {
                            neg += (syn[i >> 5] >> (i & 31)) & 1u;
                            rsum[i] = RT::mk(m1, m2);
                            rmeta[i] = arg | (neg << 31);
                        }
                    }
                    __syncthreads();                                  // the tile is consumed: its buffer may be refilled
                }
            } else
            for (int i = tid; i < rows; i += NT) {
                const R* vr = V + i * RS;
                R m1 = RT::big(), m2 = RT::big();
                uint32_t arg = 0, neg = (syn[i >> 5] >> (i & 31)) & 1u;
                if (PS) {
                    R prod = (neg & 1u) ? R(-1) : R(1);
                    int zc = 0;
                    for (int s = 0; s < RS; ++s) {
                        const R t = vr[s];
                        if (t == R(0)) ++zc;
                        else prod = RT::mul(prod, t);
                    }
                    rsum[i] = RT::mk(prod, static_cast<R>(zc));
                    continue;
                }
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
                        if (PS) {
                            R x = R(0);
                            if (s.y == R(0)) x = TT::div(s.x, v);
                            else if (s.y == R(1) && v == R(0)) x = s.x;
                            c[q] = TT::lg(TT::div(RT::add(R(1), x), RT::add(R(1), -x)));
                        } else {
                            const uint32_t meta = rmeta[row];
                            const R mag = slot == (meta & 0x7FFFFFFFu) ? s.y : s.x;
                            const uint32_t odd = (meta >> 31) ^ (v <= R(0) ? 1u : 0u);
                            c[q] = RT::mul(mag, odd ? -alpha : alpha);
                        }
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
                    if (e[q] != kNoEdge) V[addr[q]] = PS ? TT::th(RT::mul(vn[q], R(0.5))) : vn[q];
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

// sign bit set <=> x <= 0  (x == +0 becomes -0): one addition of -0 in round-down mode -- x + (-0) is x for every non-zero x in
// any rounding mode, (+0) + (-0) is -0 when rounding down, (-0) + (-0) is -0
__device__ __forceinline__ float canonical_sign(const float x) { return __fadd_rd(x, -0.0f); }
__device__ __forceinline__ double canonical_sign(const double x) { return __dadd_rd(x, -0.0); }

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
    const R l0c = MODE == 0 ? canonical_sign(l0) : l0;
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
    for (int q = 0; q < W; ++q) SH::st(x.vbase + vo[q], canonical_sign(vn[q]));
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
    for (int i = tid; i < bp_dummy_slots(sizeof(R)); i += NT) V[rows * RS + i] = R(0);
    bool s_first = true;

    // persistent grid: CTAs pull shots from a queue (shots run 1 to 10 iterations: a static stride would leave a tail, one CTA
    // per shot pays the CTA set-up -- prior table, dummy slots -- and the launch's drain for every shot)
    __shared__ int s_shot;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_shot = b.bp_next ? atomicAdd(b.bp_next, 1) : (s_first ? static_cast<int>(blockIdx.x) : b.n_shots);
        s_first = false;
        __syncthreads();
        const int shot = s_shot;
        if (shot >= b.n_shots) break;
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
#pragma unroll 2                     // (unrolling by 8 measured 2.5 % slower: registers)
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
            uint32_t* const cnd = cand;
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
            // ---- stop test H e == s (a per-warp test on a double-buffered candidate, without the second barrier, measured 2.7 % slower)
            __syncthreads();
            const int mismatch = tid < w.rowsW32 ? (cnd[tid] != syn[tid]) : 0;
            if (!__syncthreads_or(mismatch)) { conv = true; break; }
        }
        if (it > p.max_iter) it = p.max_iter;
        finish_shot<R, NT, true>(w, b, shot, tid, conv, it, 0u, syn, accs, car, hist, ebits);
    }
}

// kernel variants: (precision, column-weight template, V in shared or global)
using KernelPtr = void (*)(const WinDev, const BatchDev, const BpParams);

struct Variant {
    KernelPtr fn;
    int threads;
    size_t configured[kMaxDevices];
};

// QB_BP_STAGE=1: the global-slab variant stages its check sweep through cp.async.bulk (UBLKCP).  Measured on BASELINE config 5
// (2250 x 31500 windows, fp64, 16384 shots): BP 3141 ms staged vs 3095 ms direct -- the sweep that binds that kernel is the bit
// sweep's scattered 8-byte gathers over a 232 MB working set, not the row walk -- so it is off by default.
inline bool staged_enabled() {
    static const int on = [] { const char* e = getenv("QB_BP_STAGE"); return e ? atoi(e) : 0; }();
    return on != 0;
}

template <typename R, int CW, bool VG, bool PS>
KernelPtr pick_kernel(int* threads) {
    // fp32: 256 threads x 4 CTAs/SM (64 regs); fp64: 512 threads x 2 CTAs/SM; wide columns and product-sum get more registers
    constexpr int NT = sizeof(R) == 4 ? 256 : 512;
    constexpr int MINB = (CW <= 8 && !PS) ? (sizeof(R) == 4 ? 4 : 2) : 1;       // (two CTAs per SM at 64 registers: 27 % slower on config 5 -- spills, and twice the slabs competing for L2)
    *threads = NT;
    if (VG && staged_enabled()) return bp_kernel<R, CW, NT, MINB, VG, PS, true>;
    return bp_kernel<R, CW, NT, MINB, VG, PS, false>;
}

Variant& variant(int prec, int cw, bool vg, int method = 0) {
    static Variant table[2][3][2][2] = {};
    const int pi = prec == 32 ? 0 : 1, ci = cw <= 6 ? 0 : (cw <= 8 ? 1 : 2), gi = vg ? 1 : 0;
    Variant& v = table[pi][ci][gi][method ? 1 : 0];
    if (!v.fn) {
#define QB_PICK(R, CWV) (method ? (vg ? pick_kernel<R, CWV, true, true>(&v.threads) : pick_kernel<R, CWV, false, true>(&v.threads)) \
                                : (vg ? pick_kernel<R, CWV, true, false>(&v.threads) : pick_kernel<R, CWV, false, false>(&v.threads)))
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
        static const int nt = [] { const char* e = getenv("QB_BP_NT"); return e ? atoi(e) : 0; }();      // experiments
        if (prec == 32) { v.fn = unit ? bp_kernel_ms2<float, 256, 4, true> : bp_kernel_ms2<float, 256, 4, false>; v.threads = 256; }
        else if (nt == 384) { v.fn = unit ? bp_kernel_ms2<double, 384, 2, true> : bp_kernel_ms2<double, 384, 2, false>; v.threads = 384; }
        else if (nt == 256) { v.fn = unit ? bp_kernel_ms2<double, 256, 2, true> : bp_kernel_ms2<double, 256, 2, false>; v.threads = 256; }
        else { v.fn = unit ? bp_kernel_ms2<double, 512, 2, true> : bp_kernel_ms2<double, 512, 2, false>; v.threads = 512; }
    }
    return v;
}

}  // namespace

size_t bp_slab_bytes(const WinDev& w, int precision) { return bp_slab_elems(w) * (precision == 32 ? 4 : 8); }

size_t bp_smem_bytes(const WinDev& w, int precision, bool vglobal) {
    size_t off[11];
    if (use_compact(w, vglobal)) return std::max(bpc_layout(w, precision == 32 ? 4 : 8, off), bpm_layout(w, precision == 32 ? 4 : 8, off));
    return bp_layout(w, precision == 32 ? 4 : 8, vglobal, off, staged_enabled());
}

int bp_threads(int precision) { return precision == 32 ? 256 : 512; }

bool bp_ms2_enabled() { return ms2_enabled(); }

// CTAs of the persistent grid the flooding min-sum kernel is launched with (0: the window takes another kernel, one CTA per shot)
int bp_persistent_grid(const WinDev& w, int precision, bool vglobal, int method) {
    static const int on = [] { const char* e = getenv("QB_BP_PERSIST"); return e ? atoi(e) : 1; }();
    if (!on || !use_ms2(w, vglobal, method)) return 0;
    return 148 * (precision == 32 ? 4 : 2);
}

bool bp_supports(const WinDev& w, int method, bool vglobal) { (void)w; (void)vglobal; return method == 0 || method == 1; }

cudaError_t bp_configure(const WinDev& w, int precision, bool vglobal, int method) {
    if (w.cw > 16) return cudaErrorInvalidValue;
    if (!bp_supports(w, method, vglobal)) return cudaErrorInvalidValue;
    const size_t smem = bp_smem_bytes(w, precision, vglobal);
    if (smem > 227 * 1024) return cudaErrorInvalidValue;
    Variant& v = use_ms2(w, vglobal, method) ? ms2_variant(precision, w.unit_alpha != 0)
                 : (use_compact(w, vglobal) ? compact_variant(precision, method) : variant(precision, w.cw, vglobal, method));
    size_t& have = v.configured[device_slot()];
    if (smem <= have) return cudaSuccess;                   // the attribute only ever grows (decoders of different sizes coexist)
    cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess) have = smem;
    return e;
}

cudaError_t launch_bp(const WinDev& w, const BatchDev& b, const BpParams& p, int precision, bool vglobal, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    Variant& v = use_ms2(w, vglobal, p.method) ? ms2_variant(precision, w.unit_alpha != 0)
                 : (use_compact(w, vglobal) ? compact_variant(precision, p.method) : variant(precision, w.cw, vglobal, p.method));
    const size_t smem = bp_smem_bytes(w, precision, vglobal);
    v.fn<<<grid, v.threads, smem, st>>>(w, b, p);
    return cudaGetLastError();
}

}  // namespace qb
