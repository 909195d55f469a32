// quits_b200/csrc/osd.cu -- K4: OSD-0 for the windows BP left unconverged, one shot per 128-thread CTA (sm_100a).
//
// Replaces the OSD stage of ldpc.BpOsdDecoder.decode() (reference call site src/quits/decoder/sliding_window.py:171,182;
// osd_method 'osd_0', or 'osd_cs'/'osd_e' with osd_order = 0 which are the same thing).  Definition it is held to
// (the CPU oracle's): order the columns by ascending BP posterior, ties by column index; row-reduce in that
// column order picking as pivot row the first row at or below the current rank that has a 1 and swapping it into
// place; the solution is the reduced syndrome on the pivot columns, 0 elsewhere.
//
// How it is computed here.  Instead of reducing the m x n matrix, the kernel keeps only the accumulated row
// transformation T (m x m over GF(2), plus the syndrome as an extra column), one column per thread-slot IN
// REGISTERS as MW 32-bit words over the physical rows.  A candidate column of H is sparse (<= 6 rows), so its
// reduced form is the XOR of <= 6 columns of T; 128 candidates (the next 128 columns in sorted order) are carried
// in registers and receive the same rank-1 updates as T, so T only has to be spilled to shared memory when a new
// batch of candidates is fetched.  Row swaps are virtual: `seq` lists the free rows in the oracle's position order
// and a pivot only moves the head of that list into the vacated slot.  Work per pivot is one rank-1 update of
// (m + 1 + 128) register-resident bit-columns; no m x n traffic at all.
// The column order comes from a stable 4-pass LSD radix sort of the posteriors' order-preserving integer image
// (warp-private histograms + __match_any_sync ranking), which is exactly "ascending LLR, ties by index".
#include <cfloat>
#include <type_traits>

#include "qb_device.h"

namespace qb {

namespace {

constexpr int kOsdThreads = 128;
constexpr int kOsdWarps = 4;

struct OsdLayout {
    size_t keys, idxA, idxB, hist, rbuf, flags, seq, pivcol, pivrow, sprime, accs, car, total;
    int n_pad, TS;
};

__host__ __device__ inline size_t au(size_t x) { return (x + 15) / 16 * 16; }

__host__ __device__ inline OsdLayout osd_layout(const WinDev& w, int MW, int CPT, int ksize) {
    OsdLayout L;
    L.n_pad = (w.ncols + kOsdThreads - 1) / kOsdThreads * kOsdThreads;
    L.TS = CPT * kOsdThreads;
    size_t o = 0;
    // region 0: sort keys + first index buffer; reused afterwards as the spill area of T (MW * TS words)
    size_t sortA = au(static_cast<size_t>(L.n_pad) * ksize) + au(static_cast<size_t>(L.n_pad) * 2);
    size_t tdump = au(static_cast<size_t>(MW) * L.TS * 4);
    L.keys = o;
    L.idxA = o + au(static_cast<size_t>(L.n_pad) * ksize);
    o += sortA > tdump ? sortA : tdump;
    L.idxB = o; o += au(static_cast<size_t>(L.n_pad) * 2);
    L.hist = o; o += au(kOsdWarps * 256 * 4);
    L.rbuf = o; o += au(2 * static_cast<size_t>(MW) * 4);
    L.flags = o; o += au(2 * kOsdWarps * 4);
    L.seq = o; o += au(static_cast<size_t>(w.rows) * 2);
    L.pivcol = o; o += au(static_cast<size_t>(w.rows) * 2);
    L.pivrow = o; o += au(static_cast<size_t>(w.rows) * 2);
    L.sprime = o; o += au(static_cast<size_t>(MW) * 4);
    L.accs = o; o += au(static_cast<size_t>(w.KW) * 8);
    L.car = o; o += au(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4);
    L.total = o;
    return L;
}

// Word `wsel` of every register-resident column of this thread (T columns, then the candidate), and the pivot row
// leaves the free mask.  wsel is uniform over the CTA, so the switch is a non-divergent jump, not a select chain.
template <int MW, int CPT>
__device__ __forceinline__ void pivot_words(const uint32_t (&T)[CPT][MW], const uint32_t (&cand)[MW], uint32_t (&freem)[MW],
                                            const int wsel, const uint32_t bsel, uint32_t (&out)[CPT + 1]) {
#define QB_CASE(I)                                                   \
    case I:                                                          \
        if constexpr (I < MW) {                                      \
            _Pragma("unroll") for (int c = 0; c < CPT; ++c) out[c] = T[c][I]; \
            out[CPT] = cand[I];                                      \
            freem[I] &= ~bsel;                                       \
        }                                                            \
        break;
    switch (wsel) {
        QB_CASE(0) QB_CASE(1) QB_CASE(2) QB_CASE(3) QB_CASE(4) QB_CASE(5) QB_CASE(6) QB_CASE(7)
        QB_CASE(8) QB_CASE(9) QB_CASE(10) QB_CASE(11) QB_CASE(12) QB_CASE(13) QB_CASE(14) QB_CASE(15)
        QB_CASE(16) QB_CASE(17) QB_CASE(18) QB_CASE(19) QB_CASE(20) QB_CASE(21) QB_CASE(22) QB_CASE(23)
    default: break;
    }
#undef QB_CASE
}

// order-preserving unsigned image of a posterior (-0.0 is first folded into +0.0)
__device__ __forceinline__ uint32_t order_key(float f) {
    const uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t order_key(double f) {
    const uint64_t u = static_cast<uint64_t>(__double_as_longlong(f + 0.0));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

template <typename R, int MW, int CPT, bool EXACT>
__global__ void __launch_bounds__(kOsdThreads) osd_kernel(const WinDev w, const BatchDev b) {
    using KeyT = typename std::conditional<sizeof(R) == 4, uint32_t, uint64_t>::type;
    constexpr int kPasses = static_cast<int>(sizeof(KeyT));
    extern __shared__ __align__(16) unsigned char sm[];
    const OsdLayout L = osd_layout(w, MW, CPT, sizeof(KeyT));
    KeyT* keys = reinterpret_cast<KeyT*>(sm + L.keys);
    uint16_t* idxA = reinterpret_cast<uint16_t*>(sm + L.idxA);
    uint16_t* idxB = reinterpret_cast<uint16_t*>(sm + L.idxB);
    uint32_t* tdump = reinterpret_cast<uint32_t*>(sm + L.keys);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sm + L.hist);
    uint32_t* rbuf = reinterpret_cast<uint32_t*>(sm + L.rbuf);
    uint32_t* flags = reinterpret_cast<uint32_t*>(sm + L.flags);
    uint16_t* seq = reinterpret_cast<uint16_t*>(sm + L.seq);
    uint16_t* pivcol = reinterpret_cast<uint16_t*>(sm + L.pivcol);
    uint16_t* pivrow = reinterpret_cast<uint16_t*>(sm + L.pivrow);
    uint32_t* sprime = reinterpret_cast<uint32_t*>(sm + L.sprime);
    uint32_t* accs = reinterpret_cast<uint32_t*>(sm + L.accs);
    uint32_t* car = reinterpret_cast<uint32_t*>(sm + L.car);
    __shared__ int s_job;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = w.rows, n = w.ncols, TS = L.TS;
    const int carryW = (w.carry_rows + 31) / 32;
    const int count = *b.fail_count;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_job = atomicAdd(b.osd_next, 1);
        __syncthreads();
        const int job = s_job;
        if (job >= count) break;
        const int shot = b.fail_list[job];
        const R* llr = reinterpret_cast<const R*>(b.llr_buf) + static_cast<size_t>(shot) * b.llr_stride;
        const uint32_t* syn = b.syn_buf + static_cast<size_t>(shot) * b.syn_stride32;

        // ------------------------------------------------------------------ 1. sort columns by (LLR, index)
        for (int i = tid; i < n; i += kOsdThreads) keys[i] = order_key(llr[i]);
        const int quarter = ((n + kOsdWarps - 1) / kOsdWarps + 31) / 32 * 32;
        const int wbeg = warp * quarter, wend = min(n, wbeg + quarter);
#pragma unroll 1
        for (int pass = 0; pass < kPasses; ++pass) {
            const int shift = 8 * pass;
            const uint16_t* src = (pass & 1) ? idxA : idxB;       // pass 0 reads the identity
            uint16_t* dst = (pass & 1) ? idxB : idxA;
            for (int i = tid; i < kOsdWarps * 256; i += kOsdThreads) hist[i] = 0;
            __syncthreads();
            for (int i0 = wbeg; i0 < wend; i0 += 32) {
                const int i = i0 + lane;
                if (i < wend) {
                    const int id = pass == 0 ? i : src[i];
                    atomicAdd(&hist[warp * 256 + static_cast<uint32_t>((keys[id] >> shift) & 255u)], 1u);
                }
            }
            __syncthreads();
            {   // exclusive scan over (digit major, warp minor): thread t owns digits 2t, 2t+1
                uint32_t loc[2 * kOsdWarps];
                uint32_t sum = 0;
#pragma unroll
                for (int q = 0; q < 2 * kOsdWarps; ++q) {
                    const int d = 2 * tid + q / kOsdWarps, wq = q % kOsdWarps;
                    loc[q] = sum;
                    sum += hist[wq * 256 + d];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
                    if (lane >= o) inc += t;
                }
                __shared__ uint32_t wsum[kOsdWarps];
                if (lane == 31) wsum[warp] = inc;
                __syncthreads();
                uint32_t base = inc - sum;
                for (int q = 0; q < warp; ++q) base += wsum[q];
#pragma unroll
                for (int q = 0; q < 2 * kOsdWarps; ++q) {
                    const int d = 2 * tid + q / kOsdWarps, wq = q % kOsdWarps;
                    hist[wq * 256 + d] = base + loc[q];
                }
            }
            __syncthreads();
            for (int i0 = wbeg; i0 < wend; i0 += 32) {
                const int i = i0 + lane;
                const bool valid = i < wend;
                int id = 0;
                uint32_t dg = 256u + lane;                           // invalid lanes never match anybody
                if (valid) {
                    id = pass == 0 ? i : src[i];
                    dg = static_cast<uint32_t>((keys[id] >> shift) & 255u);
                }
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, dg);
                const uint32_t before = peers & ((1u << lane) - 1u);
                if (valid) {
                    const uint32_t pos = hist[warp * 256 + dg] + __popc(before);
                    dst[pos] = static_cast<uint16_t>(id);
                }
                __syncwarp();
                if (valid && (peers >> lane) <= 1u) hist[warp * 256 + dg] += __popc(peers);     // highest peer updates the base
                __syncwarp();
            }
            __syncthreads();
        }
        const uint16_t* order = idxB;        // the last (odd-numbered) pass wrote idxB

        // ------------------------------------------------------------------ 2. Gauss-Jordan on T (registers)
        uint32_t T[CPT][MW];
        uint32_t cand[MW];
        uint32_t freem[MW];
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int k = tid + c * kOsdThreads;
#pragma unroll
            for (int i = 0; i < MW; ++i) {
                uint32_t v = 0;
                if (k < m && (k >> 5) == i) v = 1u << (k & 31);
                if (k == m && i < w.rowsW32) v = syn[i];
                T[c][i] = v;
            }
        }
#pragma unroll
        for (int i = 0; i < MW; ++i) {
            const int left = m - 32 * i;
            freem[i] = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << left) - 1u : 0u);
            cand[i] = 0;
        }
        for (int i = tid; i < m; i += kOsdThreads) seq[i] = static_cast<uint16_t>(i);
        if (tid < 2 * w.KW) accs[tid] = 0;
        if (tid <= carryW) car[tid] = 0;
        int rank = 0, base = 0, par = 0, candcol = -1;
        int pend_p = -1, pend_rank = 0;          // deferred seq update (thread 0)
        bool refill = true;
        while (rank < m) {
            if (refill) {
                if (base >= n) break;
                __syncthreads();
#pragma unroll
                for (int c = 0; c < CPT; ++c)
#pragma unroll
                    for (int i = 0; i < MW; ++i) tdump[i * TS + tid + c * kOsdThreads] = T[c][i];
                __syncthreads();
                const int pos = base + tid;
#pragma unroll
                for (int i = 0; i < MW; ++i) cand[i] = 0;
                candcol = -1;
                if (pos < n) {
                    candcol = order[pos];
                    const int qb = __ldg(w.cptr + candcol), qe = __ldg(w.cptr + candcol + 1);
                    for (int q = qb; q < qe; ++q) {
                        const int row = __ldg(w.crow + q);
#pragma unroll
                        for (int i = 0; i < MW; ++i) cand[i] ^= tdump[i * TS + row];
                    }
                }
                base += kOsdThreads;
                refill = false;
            }
            // -- A: first candidate (in sorted order) that still has a 1 in a free row
            uint32_t any = 0;
#pragma unroll
            for (int i = 0; i < MW; ++i) any |= cand[i] & freem[i];
            const uint32_t ball = __ballot_sync(0xFFFFFFFFu, any != 0);
            if (lane == 0) flags[par * kOsdWarps + warp] = ball;
            __syncthreads();
            if (tid == 0 && pend_p >= 0) { seq[pend_p] = seq[pend_rank]; pend_p = -1; }
            int pt = -1;
#pragma unroll
            for (int q = kOsdWarps - 1; q >= 0; --q) {
                const uint32_t f = flags[par * kOsdWarps + q];
                if (f) pt = q * 32 + __ffs(f) - 1;
            }
            if (pt < 0) { refill = true; continue; }
            if (tid == pt) {
                pivcol[rank] = static_cast<uint16_t>(candcol);
                if (EXACT) {
#pragma unroll
                    for (int i = 0; i < MW; ++i) rbuf[par * MW + i] = cand[i];
                } else {
                    // full-row-rank window: the solution does not depend on which free row becomes the pivot row, so take
                    // the first one and publish the update vector with that bit already cleared
                    int pr = -1;
#pragma unroll
                    for (int i = 0; i < MW; ++i) {
                        const uint32_t x = cand[i] & freem[i];
                        if (pr < 0 && x) pr = 32 * i + __ffs(x) - 1;
                    }
#pragma unroll
                    for (int i = 0; i < MW; ++i) rbuf[par * MW + i] = cand[i] & ~((i == (pr >> 5)) ? (1u << (pr & 31)) : 0u);
                    pivrow[rank] = static_cast<uint16_t>(pr);
                }
            }
            __syncthreads();
            // -- C: pivot row, then the rank-1 update of every register-resident column
            const uint32_t* rb = rbuf + par * MW;
            int prow;
            if (EXACT) {
                // rank-deficient window (inconsistent syndromes are possible): follow the oracle's row order exactly,
                // i.e. the first free row in position order that has a 1
                int p = -1;
                for (int c0 = rank; c0 < m; c0 += 32) {
                    const int pos = c0 + lane;
                    uint32_t bit = 0;
                    if (pos < m) {
                        const int row = seq[pos];
                        bit = (rb[row >> 5] >> (row & 31)) & 1u;
                    }
                    const uint32_t bb = __ballot_sync(0xFFFFFFFFu, bit);
                    if (bb) { p = c0 + __ffs(bb) - 1; break; }
                }
                prow = seq[p];
                if (tid == 0) {
                    pivrow[rank] = static_cast<uint16_t>(prow);
                    pend_p = p;
                    pend_rank = rank;
                }
            } else {
                prow = pivrow[rank];
            }
            const int wsel = prow >> 5;
            const uint32_t bsel = 1u << (prow & 31);
            uint32_t r[MW];
#pragma unroll
            for (int i = 0; i < MW; ++i) r[i] = EXACT ? (rb[i] & ~((i == wsel) ? bsel : 0u)) : rb[i];
            uint32_t pw[CPT + 1];
            pivot_words<MW, CPT>(T, cand, freem, wsel, bsel, pw);
#pragma unroll
            for (int c = 0; c < CPT; ++c) {
                if (pw[c] & bsel) {
#pragma unroll
                    for (int i = 0; i < MW; ++i) T[c][i] ^= r[i];
                }
            }
            if (pw[CPT] & bsel) {
#pragma unroll
                for (int i = 0; i < MW; ++i) cand[i] ^= r[i];
            }
            ++rank;
            par ^= 1;
        }
        // ------------------------------------------------------------------ 3. solution on the pivots, commit
        __syncthreads();
        {
            const int ks = m % kOsdThreads, cs = m / kOsdThreads;       // owner of the syndrome column
            if (tid == ks) {
#pragma unroll
                for (int c = 0; c < CPT; ++c)
                    if (c == cs) {
#pragma unroll
                        for (int i = 0; i < MW; ++i) sprime[i] = T[c][i];
                    }
            }
        }
        __syncthreads();
        for (int rr = tid; rr < rank; rr += kOsdThreads) {
            const int row = pivrow[rr];
            if (!((sprime[row >> 5] >> (row & 31)) & 1u)) continue;
            const int j = pivcol[rr];
            if (b.ehat_out) atomicOr(&b.ehat_out[static_cast<size_t>(shot) * b.ehat_stride32 + (j >> 5)], 1u << (j & 31));
            if (j < w.ncommit) {
                for (int wd = 0; wd < w.KW; ++wd) {
                    const uint64_t lm = __ldg(w.lmask + static_cast<size_t>(j) * w.KW + wd);
                    if (static_cast<uint32_t>(lm)) atomicXor(&accs[2 * wd], static_cast<uint32_t>(lm));
                    if (static_cast<uint32_t>(lm >> 32)) atomicXor(&accs[2 * wd + 1], static_cast<uint32_t>(lm >> 32));
                }
                if (w.carry_rows) {
                    for (int q = __ldg(w.uptr + j); q < __ldg(w.uptr + j + 1); ++q) {
                        const uint32_t ur = __ldg(w.uidx + q);
                        atomicXor(&car[ur >> 5], 1u << (ur & 31));
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
        if (tid == 0) atomicAdd(&b.stats[2], 1ull);
    }
}

struct OsdShape { int MW, CPT; };

// instantiated (MW, CPT) shapes: MW 32-bit words cover the rows, CPT*128 thread-slots cover the rows+1 columns of T
inline bool osd_shape(const WinDev& w, OsdShape& s) {
    static const int shapes[][2] = {{4, 2}, {6, 2}, {8, 2}, {8, 3}, {12, 3}, {12, 4}, {17, 5}, {23, 6}};
    if (w.ncols > 65535 || w.rows > 65535) return false;
    for (auto& sh : shapes) {
        if (w.rows <= sh[0] * 32 && w.rows + 1 <= sh[1] * kOsdThreads) {
            s.MW = sh[0];
            s.CPT = sh[1];
            return true;
        }
    }
    return false;
}

template <typename R, bool EXACT, typename F>
inline cudaError_t osd_dispatch_r(const OsdShape& s, F&& f) {
    switch (s.MW * 10 + s.CPT) {
    case 42: return f(osd_kernel<R, 4, 2, EXACT>);
    case 62: return f(osd_kernel<R, 6, 2, EXACT>);
    case 82: return f(osd_kernel<R, 8, 2, EXACT>);
    case 83: return f(osd_kernel<R, 8, 3, EXACT>);
    case 123: return f(osd_kernel<R, 12, 3, EXACT>);
    case 124: return f(osd_kernel<R, 12, 4, EXACT>);
    case 175: return f(osd_kernel<R, 17, 5, EXACT>);
    case 236: return f(osd_kernel<R, 23, 6, EXACT>);
    }
    return cudaErrorInvalidValue;
}

template <typename F>
inline cudaError_t osd_dispatch(const WinDev& w, int precision, F&& f) {
    OsdShape s;
    if (!osd_shape(w, s)) return cudaErrorInvalidValue;
    if (w.full_row_rank) return precision == 32 ? osd_dispatch_r<float, false>(s, f) : osd_dispatch_r<double, false>(s, f);
    return precision == 32 ? osd_dispatch_r<float, true>(s, f) : osd_dispatch_r<double, true>(s, f);
}

}  // namespace

size_t osd_smem_bytes(const WinDev& w, int precision) {
    OsdShape s;
    if (!osd_shape(w, s)) return 0;
    return osd_layout(w, s.MW, s.CPT, precision == 32 ? 4 : 8).total;
}

bool osd_supported(const WinDev& w, int precision) {
    OsdShape s;
    return osd_shape(w, s) && osd_smem_bytes(w, precision) <= 220 * 1024;
}

cudaError_t osd_configure(const WinDev& w, int precision) {
    // several windows may share one instantiation: the attribute only ever grows
    static size_t configured[4][256] = {};
    const size_t smem = osd_smem_bytes(w, precision);
    OsdShape s;
    if (!osd_shape(w, s)) return cudaErrorInvalidValue;
    size_t& have = configured[(precision == 32 ? 0 : 1) + (w.full_row_rank ? 2 : 0)][(s.MW * 10 + s.CPT) & 255];
    if (smem <= have) return cudaSuccess;
    cudaError_t e = osd_dispatch(w, precision, [&](auto kern) {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    });
    if (e == cudaSuccess) have = smem;
    return e;
}

cudaError_t launch_osd(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const size_t smem = osd_smem_bytes(w, precision);
    return osd_dispatch(w, precision, [&](auto kern) {
        kern<<<grid, kOsdThreads, smem, st>>>(w, b);
        return cudaGetLastError();
    });
}

}  // namespace qb
