// quits_b200/csrc/osd.cu -- K4: OSD-0 for the windows BP left unconverged (sm_100a).
//
// Replaces the OSD stage of ldpc.BpOsdDecoder.decode() (reference call site src/quits/decoder/sliding_window.py:171,182;
// osd_method 'osd_0', or 'osd_cs'/'osd_e' with osd_order = 0 which are the same thing).  Definition it is held to
// (the CPU oracle's): order the columns by ascending BP posterior, ties by column index; row-reduce in that
// column order picking as pivot row the first row at or below the current rank that has a 1 and swapping it into
// place; the solution is the reduced syndrome on the pivot columns, 0 elsewhere.
//
// Two kernels per window, both persistent grids pulling failed shots from the list the BP kernel wrote:
//
//   osd_sort_kernel   one 128-thread CTA per shot: stable LSD radix sort (8-bit digits) of the posteriors' order-preserving
//                     integer image in shared memory (warp-private histograms + __match_any_sync ranking), which is exactly
//                     "ascending LLR, ties by index"; the column order goes to HBM as u16 (over the shot's posterior row, which is dead by then).
//   osd_elim_kernel   ONE WARP per shot, ~10 shots resident per SM, no block barrier anywhere.  Instead of reducing the
//                     m x n matrix the warp keeps the accumulated row transformation T (m x m over GF(2)) in shared
//                     memory, one 128-bit-packed column per pivot found so far (the columns of rows that are not yet pivot
//                     rows are unit vectors and stay implicit).  A candidate column of H is sparse (<= 6 rows), so its
//                     reduced form is the XOR of <= 6 columns of T; 32 candidates (the next 32 columns in sorted order) sit
//                     in registers, one per lane.  Per pivot: ballot for the first live candidate, publish its vector r,
//                     pick the pivot row, XOR r into every candidate and every stored column of T that has the pivot
//                     row's bit, and into the syndrome column.  Row swaps are virtual: `seq` lists the free rows in the
//                     oracle's position order and a pivot only moves the head of that list into the vacated slot.
//                     The elimination stops as soon as the reduced syndrome vanishes on all free rows: from then on no
//                     pivot can touch it, so the answer is final (typically after a few hundred of the ~3600 columns).
//                     Full-row-rank windows skip `seq`: with all rows eventually pivots and consistent syndromes the
//                     solution does not depend on which free row a pivot takes, so the first set bit is used.
#include <cfloat>
#include <type_traits>

#include "qb_device.h"

namespace qb {

namespace {

constexpr int kOsdMaxOrder = 32;       // osd_cs: pairs among the first <= 32 non-pivot columns; osd_e: order <= 12 (4095 patterns)
constexpr int kSortThreads = 128;
constexpr int kSortWarps = 4;
constexpr uint32_t kFull = 0xFFFFFFFFu;

__host__ __device__ inline size_t au(size_t x) { return (x + 15) / 16 * 16; }

// ====================================================================================================== sort
struct SortLayout {
    size_t keys, idxA, idxB, hist, total;      // offsets; `total` = shared-memory bytes
    size_t slab;                               // wide windows: keys / idxA / idxB live in a global slab of this size per CTA
    int n_pad;
};

constexpr size_t kSortSmemLimit = 200 * 1024;

__host__ __device__ inline SortLayout sort_layout(const WinDev& w, int ksize) {
    SortLayout L;
    L.n_pad = (w.ncols + kSortThreads - 1) / kSortThreads * kSortThreads;
    size_t o = 0;
    L.keys = o; o += au(static_cast<size_t>(L.n_pad) * ksize);
    L.idxA = o; o += au(static_cast<size_t>(L.n_pad) * 2);
    L.idxB = o; o += au(static_cast<size_t>(L.n_pad) * 2);
    L.slab = 0;
    if (o + au(kSortWarps * 256 * 4) > kSortSmemLimit) {      // too wide for shared memory: only the histograms stay there
        L.slab = o;
        o = 0;
    }
    L.hist = o; o += au(kSortWarps * 256 * 4);
    L.total = o;
    return L;
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

template <typename R>
__global__ void __launch_bounds__(kSortThreads) osd_sort_kernel(const WinDev w, const BatchDev b) {
    using KeyT = typename std::conditional<sizeof(R) == 4, uint32_t, uint64_t>::type;
    constexpr int kPasses = static_cast<int>(sizeof(KeyT));
    extern __shared__ __align__(16) unsigned char sm[];
    const SortLayout L = sort_layout(w, sizeof(KeyT));
    unsigned char* kbase = L.slab ? static_cast<unsigned char*>(b.sort_scratch) + static_cast<size_t>(blockIdx.x) * L.slab : sm;
    KeyT* keys = reinterpret_cast<KeyT*>(kbase + L.keys);
    uint16_t* idxA = reinterpret_cast<uint16_t*>(kbase + L.idxA);
    uint16_t* idxB = reinterpret_cast<uint16_t*>(kbase + L.idxB);
    uint32_t* hist = reinterpret_cast<uint32_t*>(sm + L.hist);
    __shared__ int s_job;
    __shared__ uint32_t wsum[kSortWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = w.ncols;
    const int count = *b.fail_count;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_job = atomicAdd(b.sort_next, 1);
        __syncthreads();
        const int job = s_job;
        if (job >= count) break;
        const int shot = b.fail_list[job];
        const R* llr = reinterpret_cast<const R*>(b.llr_buf) + static_cast<size_t>(shot) * b.llr_stride;
        // the column order overwrites the head of this shot's posterior row: the keys are in shared memory by then
        uint16_t* out = b.order_alt ? b.order_alt + static_cast<size_t>(shot) * b.llr_stride : reinterpret_cast<uint16_t*>(const_cast<R*>(llr));

        for (int i = tid; i < n; i += kSortThreads) keys[i] = order_key(llr[i]);
        __syncthreads();
        const int quarter = ((n + kSortWarps - 1) / kSortWarps + 31) / 32 * 32;
        const int wbeg = warp * quarter, wend = min(n, wbeg + quarter);
#pragma unroll 1
        for (int pass = 0; pass < kPasses; ++pass) {
            const int shift = 8 * pass;
            const uint16_t* src = (pass & 1) ? idxA : idxB;       // pass 0 reads the identity
            uint16_t* dst = pass == kPasses - 1 ? out : ((pass & 1) ? idxB : idxA);
            for (int i = tid; i < kSortWarps * 256; i += kSortThreads) hist[i] = 0;
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
                uint32_t loc[2 * kSortWarps];
                uint32_t sum = 0;
#pragma unroll
                for (int q = 0; q < 2 * kSortWarps; ++q) {
                    const int d = 2 * tid + q / kSortWarps, wq = q % kSortWarps;
                    loc[q] = sum;
                    sum += hist[wq * 256 + d];
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(kFull, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) wsum[warp] = inc;
                __syncthreads();
                uint32_t base = inc - sum;
                for (int q = 0; q < warp; ++q) base += wsum[q];
#pragma unroll
                for (int q = 0; q < 2 * kSortWarps; ++q) {
                    const int d = 2 * tid + q / kSortWarps, wq = q % kSortWarps;
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
                const uint32_t peers = __match_any_sync(kFull, dg);
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
    }
}

// ====================================================================================================== elimination
struct ElimLayout {
    size_t T, rvec, freem, svec, seq, pivcol, pivrow, slot, accs, car, selkey, selidx, bins, pivpos, ispiv, skeys, pinfo, pwt, ybuf, rw, scol, swt, flips, total;
    int TS;
};

__host__ __device__ inline ElimLayout elim_layout(const WinDev& w, int NQ, bool exact, int selcap, bool hi = false, bool tglobal = false) {
    ElimLayout L;
    L.TS = (w.rows + 31) / 32 * 32;
    size_t o = 0;
    L.T = o; o += tglobal ? 0 : static_cast<size_t>(NQ) * L.TS * 16;      // tall windows keep T in a global slab per warp
    L.rvec = o; o += static_cast<size_t>(NQ) * 16;
    L.freem = o; o += static_cast<size_t>(NQ) * 16;
    L.svec = o; o += static_cast<size_t>(NQ) * 16;
    L.seq = o; o += exact ? au(static_cast<size_t>(w.rows) * 2) : 0;
    L.pivcol = o; o += au(static_cast<size_t>(w.rows) * 2);
    L.pivrow = o; o += au(static_cast<size_t>(w.rows) * 2);
    L.slot = o; o += au(static_cast<size_t>(w.rows) * 2);
    L.accs = o; o += au(static_cast<size_t>(w.KW) * 8);
    L.car = o; o += au(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4);
    L.selkey = o; o += au(static_cast<size_t>(selcap) * 8);
    L.selidx = o; o += au(static_cast<size_t>(selcap) * 2);
    L.bins = o; o += selcap ? 128 : 0;
    // higher-order OSD (osd_e / osd_cs with order > 0)
    int P = 32;
    while (P < w.rows) P <<= 1;
    L.pivpos = o; o += hi ? au(static_cast<size_t>(w.rows) * 2) : 0;
    L.ispiv = o; o += hi ? au(static_cast<size_t>((w.ncols + 31) / 32) * 4) : 0;
    L.skeys = o; o += hi ? au(static_cast<size_t>(P) * 4) : 0;
    L.pinfo = o; o += hi ? au(static_cast<size_t>(w.rows) * 4) : 0;
    L.pwt = o; o += hi ? au(static_cast<size_t>(w.rows) * 8) : 0;
    L.ybuf = o; o += hi ? au(static_cast<size_t>(32) * (4 * NQ + 1) * 4) : 0;
    L.rw = o; o += hi ? static_cast<size_t>(kOsdMaxOrder) * NQ * 16 : 0;
    L.scol = o; o += hi ? au(static_cast<size_t>(kOsdMaxOrder) * 4 * 2) : 0;
    L.swt = o; o += hi ? au(static_cast<size_t>(kOsdMaxOrder) * 8) : 0;
    L.flips = o; o += hi ? au(static_cast<size_t>(kOsdMaxOrder + 2) * 4) : 0;
    L.total = o;
    return L;
}

__device__ __forceinline__ uint32_t comp(const uint4& v, int c) { return c == 0 ? v.x : (c == 1 ? v.y : (c == 2 ? v.z : v.w)); }
__device__ __forceinline__ void xor4(uint4& a, const uint4& b) { a.x ^= b.x; a.y ^= b.y; a.z ^= b.z; a.w ^= b.w; }
__device__ __forceinline__ uint32_t and_any(const uint4& a, const uint4& b) { return (a.x & b.x) | (a.y & b.y) | (a.z & b.z) | (a.w & b.w); }

// ====================================================================================================== higher-order OSD
// osd_cs / osd_e with order > 0 (ldpc osd.hpp, restated in oracle/cref.c osd_decode): after the COMPLETE elimination, flip sets
// of non-pivot columns (in LLR order) and keep the lightest solution, weight = sum over set bits, in column-index order, of
// log(1/p_j).  One candidate per lane: the lane forms the candidate's pivot bits y = T s xor T c (the same sparse gather the
// elimination uses), parks y in shared memory and walks the pivots in column order adding weights with __dadd_rn, the flipped
// columns merged in at their place -- the same additions in the same order as the CPU oracle, so ties between candidates
// (ubiquitous with 9 distinct priors) resolve identically: first candidate that is strictly lighter wins.
template <int NQ>
__device__ __forceinline__ void gather_reduced(const WinDev& w, const uint4* T4, const int TS, const uint16_t* slot_of_row, const int col,
                                               uint4 (&v)[NQ]) {
    const int qb = __ldg(w.cptr + col), qe = __ldg(w.cptr + col + 1);
    for (int q = qb; q < qe; ++q) {
        const int row = __ldg(w.crow + q);
        const uint32_t s = slot_of_row[row];
        if (s != 0xFFFFu) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) xor4(v[i], T4[i * TS + s]);
        }                                   // rows without a pivot carry no solution bit: ignored
    }
}

template <int NQ>
__device__ void osd_higher(const WinDev& w, const BatchDev& b, unsigned char* sm, const ElimLayout& L, uint4* Tbase, const int rank,
                           const uint16_t* order, const int n, const int lane) {
    constexpr int YS = 4 * NQ + 1;
    const uint4* T4 = Tbase;
    const int TS = L.TS;
    uint4* svec = reinterpret_cast<uint4*>(sm + L.svec);
    uint32_t* svec32 = reinterpret_cast<uint32_t*>(svec);
    const uint16_t* pivcol = reinterpret_cast<const uint16_t*>(sm + L.pivcol);
    const uint16_t* pivrow = reinterpret_cast<const uint16_t*>(sm + L.pivrow);
    const uint16_t* slot_of_row = reinterpret_cast<const uint16_t*>(sm + L.slot);
    const uint16_t* pivpos = reinterpret_cast<const uint16_t*>(sm + L.pivpos);
    uint32_t* ispiv = reinterpret_cast<uint32_t*>(sm + L.ispiv);
    uint32_t* skeys = reinterpret_cast<uint32_t*>(sm + L.skeys);
    uint32_t* pinfo = reinterpret_cast<uint32_t*>(sm + L.pinfo);
    double* pwt = reinterpret_cast<double*>(sm + L.pwt);
    uint32_t* ybuf = reinterpret_cast<uint32_t*>(sm + L.ybuf) + lane * YS;
    uint4* rw = reinterpret_cast<uint4*>(sm + L.rw);
    uint32_t* scol = reinterpret_cast<uint32_t*>(sm + L.scol);            // [kOsdMaxOrder] column, then [kOsdMaxOrder] pattern bit
    uint32_t* sbit = scol + kOsdMaxOrder;
    double* swt = reinterpret_cast<double*>(sm + L.swt);
    uint32_t* flips = reinterpret_cast<uint32_t*>(sm + L.flips);
    const double* wt = w.osd_wt;

    // ---- pivot positions as a bitmap over the sorted order; pivots sorted by column index
    const int npw = (n + 31) / 32;
    for (int i = lane; i < npw; i += 32) ispiv[i] = 0;
    int P = 32;
    while (P < rank) P <<= 1;
    for (int i = lane; i < P; i += 32) skeys[i] = i < rank ? (static_cast<uint32_t>(pivcol[i]) << 16) | static_cast<uint32_t>(i) : 0xFFFFFFFFu;
    __syncwarp();
    for (int s = lane; s < rank; s += 32) atomicOr(&ispiv[pivpos[s] >> 5], 1u << (pivpos[s] & 31));
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const uint32_t a = skeys[i], c = skeys[l];
                if ((a > c) == ((i & k) == 0)) { skeys[i] = c; skeys[l] = a; }
            }
            __syncwarp();
        }
    }
    for (int t = lane; t < rank; t += 32) {
        const uint32_t s = skeys[t] & 0xFFFFu, pc = skeys[t] >> 16;
        pinfo[t] = (static_cast<uint32_t>(pivrow[s]) << 16) | pc;
        pwt[t] = __ldg(wt + pc);
    }
    // ---- re-index the pivot rows by the rank of their pivot's column: bit t of a (permuted) vector <-> t-th pivot in column
    //      order, so that a candidate's weight is a walk over its SET bits in ascending order instead of over all pivots.
    //      Rows without a pivot carry no solution bit and are dropped.
    {
        uint16_t* rowperm = reinterpret_cast<uint16_t*>(sm + L.pivpos);           // pivpos is dead once ispiv is built
        __syncwarp();
        for (int t = lane; t < rank; t += 32) rowperm[t] = static_cast<uint16_t>(pinfo[t] >> 16);      // t -> row
        __syncwarp();
        uint16_t* t_of_row = reinterpret_cast<uint16_t*>(skeys);                  // scratch: row -> t (skeys is dead)
        for (int r = lane; r < w.rows; r += 32) t_of_row[r] = 0xFFFFu;
        __syncwarp();
        for (int t = lane; t < rank; t += 32) t_of_row[rowperm[t]] = static_cast<uint16_t>(t);
        __syncwarp();
        uint4* T4w = Tbase;
        auto permute = [&](uint4* vec, const int stride) {                        // in place, through this lane's ybuf row
            for (int i = 0; i < 4 * NQ; ++i) ybuf[i] = 0;
            for (int i = 0; i < NQ; ++i) {
                const uint4 q = vec[i * stride];
                const uint32_t wd[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t bits = wd[c];
                    while (bits) {
                        const int r = 32 * (4 * i + c) + __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t t = t_of_row[r];
                        if (t != 0xFFFFu) ybuf[t >> 5] |= 1u << (t & 31);
                    }
                }
            }
            for (int i = 0; i < NQ; ++i) vec[i * stride] = make_uint4(ybuf[4 * i], ybuf[4 * i + 1], ybuf[4 * i + 2], ybuf[4 * i + 3]);
        };
        for (int sl = lane; sl < rank; sl += 32) permute(T4w + sl, TS);
        if (lane == 0) permute(svec, 1);
        __syncwarp();
    }
    // ---- the first `ord` non-pivot columns: reduced vectors, and their columns sorted by index for the weight merge
    const int ord_req = min(b.osd_order, b.osd_method == 1 ? 12 : kOsdMaxOrder);
    int ord = 0;
    int mycol = -1;
    for (int c0 = 0; c0 < n && ord < ord_req; c0 += 32) {
        const int p = c0 + lane;
        const bool np = p < n && !((ispiv[p >> 5] >> (p & 31)) & 1u);
        const uint32_t bal = __ballot_sync(kFull, np);
        const int idx = ord + __popc(bal & ((1u << lane) - 1u));
        const int colp = np ? static_cast<int>(order[p]) : -1;
        // hand the column of the idx-th non-pivot position to lane idx
#pragma unroll 1
        for (uint32_t mm = bal; mm; mm &= mm - 1) {
            const int src = __ffs(mm) - 1;
            const int di = __shfl_sync(kFull, idx, src), dc = __shfl_sync(kFull, colp, src);
            if (di < ord_req && lane == di) mycol = dc;
        }
        ord = min(ord_req, ord + __popc(bal));
    }
    __syncwarp();
    {
        uint4 v[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) v[i] = make_uint4(0u, 0u, 0u, 0u);
        if (lane < ord) gather_reduced<NQ>(w, T4, TS, slot_of_row, mycol, v);
        if (lane < ord) {
#pragma unroll
            for (int i = 0; i < NQ; ++i) rw[lane * NQ + i] = v[i];
        }
        // rank of my column among the `ord` (distinct) columns
        int rk = 0;
        for (int o = 0; o < ord; ++o) rk += __shfl_sync(kFull, mycol, o) < mycol ? 1 : 0;
        if (lane < ord) { scol[rk] = static_cast<uint32_t>(mycol); sbit[rk] = static_cast<uint32_t>(lane); swt[rk] = __ldg(wt + mycol); }
    }
    __syncwarp();

    // weight of a candidate: its pivot bits v (permuted: bit t <-> t-th pivot in column order) plus its flipped columns -- one
    // column xc (xc < 0: none) or the members of `pat` among the first `ord` non-pivot columns -- added in ascending column
    // order.  tpos(c) = number of pivots whose column is < c (binary search over the sorted pivot columns).
    auto tpos = [&](const int c) {
        int lo = 0, hi = rank;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (static_cast<int>(pinfo[mid] & 0xFFFFu) < c) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    auto weigh = [&](const uint4 (&v)[NQ], const int xc, const uint32_t pat) -> double {
        double ws = 0.0;
        int ip = 0;                                       // next flipped column (pattern members in ascending column order)
        int nextcol = 0x7FFFFFFF, nextt = 0x7FFFFFFF;
        double nextw = 0.0;
        bool single = xc >= 0;
        auto advance = [&]() {
            nextt = 0x7FFFFFFF;
            if (single) { single = false; nextcol = xc; nextw = __ldg(wt + xc); nextt = tpos(xc); return; }
            while (ip < ord && !((pat >> sbit[ip]) & 1u)) ++ip;
            if (ip < ord) { nextcol = static_cast<int>(scol[ip]); nextw = swt[ip]; nextt = tpos(nextcol); ++ip; }
        };
        advance();
#pragma unroll
        for (int i = 0; i < NQ; ++i) {
            const uint32_t wd[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t bits = wd[c];
                while (bits) {
                    const int t = 32 * (4 * i + c) + __ffs(bits) - 1;
                    bits &= bits - 1;
                    while (nextt <= t) { ws = __dadd_rn(ws, nextw); advance(); }
                    ws = __dadd_rn(ws, pwt[t]);
                }
            }
        }
        while (nextt != 0x7FFFFFFF) { ws = __dadd_rn(ws, nextw); advance(); }
        return ws;
    };

    // ---- OSD-0 solution
    uint4 y0[NQ];
#pragma unroll
    for (int i = 0; i < NQ; ++i) y0[i] = svec[i];
    const double w0 = weigh(y0, -1, 0u);
    double bestw = w0;
    int bestid = 0x7FFFFFFF, bestkind = 0;
    uint32_t bestarg = 0;
    int nextid = 0;
    // ---- combination sweep: every single non-pivot column, in LLR order
    if (b.osd_method == 2) {
        for (int c0 = 0; c0 < n; c0 += 32) {
            const int p = c0 + lane;
            const bool np = p < n && !((ispiv[p >> 5] >> (p & 31)) & 1u);
            const uint32_t bal = __ballot_sync(kFull, np);
            if (!bal) continue;
            const int id = nextid + __popc(bal & ((1u << lane) - 1u));
            nextid += __popc(bal);
            int col = -1;
            uint4 v[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) v[i] = y0[i];
            if (np) {
                col = order[p];
                gather_reduced<NQ>(w, T4, TS, slot_of_row, col, v);
            }
            const double cw = np ? weigh(v, col, 0u) : 0.0;
            if (np && cw < bestw) { bestw = cw; bestid = id; bestkind = 1; bestarg = static_cast<uint32_t>(col); }
        }
    }
    // ---- patterns over the first `ord` non-pivot columns: pairs i < j (combination sweep) or every non-empty subset (exhaustive)
    {
        const int npat = b.osd_method == 2 ? ord * (ord - 1) / 2 : (1 << ord) - 1;
        for (int c0 = 0; c0 < npat; c0 += 32) {
            const int c = c0 + lane;
            const bool act = c < npat;
            uint32_t pat = 0;
            if (act) {
                if (b.osd_method == 2) {                 // c-th pair in the order i = 0.., j = i+1..
                    int i = 0, rem = c;
                    while (rem >= ord - 1 - i) { rem -= ord - 1 - i; ++i; }
                    pat = (1u << i) | (1u << (i + 1 + rem));
                } else {
                    pat = static_cast<uint32_t>(c) + 1u;
                }
            }
            uint4 v[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) v[i] = y0[i];
            for (int bb = 0; bb < ord; ++bb) {
                if ((pat >> bb) & 1u) {
#pragma unroll
                    for (int i = 0; i < NQ; ++i) xor4(v[i], rw[bb * NQ + i]);
                }
            }
            const double cw = act ? weigh(v, -1, pat) : 0.0;
            if (act && cw < bestw) { bestw = cw; bestid = nextid + c; bestkind = 2; bestarg = pat; }
        }
    }
    // ---- the earliest candidate of minimal weight, if strictly lighter than OSD-0
    double mw = bestw;
    int mid = bestid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ow = __shfl_xor_sync(kFull, mw, o);
        const int oid = __shfl_xor_sync(kFull, mid, o);
        if (ow < mw || (ow == mw && oid < mid)) { mw = ow; mid = oid; }
    }
    if (lane == 0) flips[0] = 0;
    __syncwarp();
    if (mw < w0 && bestid == mid && bestid != 0x7FFFFFFF) {        // exactly one lane owns the winning id
        uint4 v[NQ];
#pragma unroll
        for (int i = 0; i < NQ; ++i) v[i] = y0[i];
        int nf = 0;
        if (bestkind == 1) {
            gather_reduced<NQ>(w, T4, TS, slot_of_row, static_cast<int>(bestarg), v);
            flips[1 + nf++] = bestarg;
        } else {
            for (int bb = 0; bb < ord; ++bb) {
                if ((bestarg >> bb) & 1u) {
#pragma unroll
                    for (int i = 0; i < NQ; ++i) xor4(v[i], rw[bb * NQ + i]);
                }
            }
            for (int q = 0; q < ord; ++q)
                if ((bestarg >> sbit[q]) & 1u) flips[1 + nf++] = scol[q];
        }
#pragma unroll
        for (int i = 0; i < NQ; ++i) svec[i] = v[i];
        flips[0] = static_cast<uint32_t>(nf);
    }
    __syncwarp();
}

// One shot, one warp: eliminate over `order[0 .. n_avail)` (the first n_avail columns of the OSD order out of n).  Returns
// false -- and commits nothing -- when those columns ran out before the answer was final although more columns exist.
template <int NQ, bool EXACT, bool HI = false>
__device__ __forceinline__ bool elim_job(const WinDev& w, const BatchDev& b, unsigned char* sm, const ElimLayout& L, uint4* Tbase,
                                         const int shot, const uint16_t* order, const int n, const int n_total) {
    uint4* T4 = Tbase;
    const uint32_t* T32 = reinterpret_cast<const uint32_t*>(Tbase);
    uint4* rvec = reinterpret_cast<uint4*>(sm + L.rvec);
    uint4* freem = reinterpret_cast<uint4*>(sm + L.freem);
    uint4* svec = reinterpret_cast<uint4*>(sm + L.svec);
    uint32_t* rvec32 = reinterpret_cast<uint32_t*>(rvec);
    uint32_t* freem32 = reinterpret_cast<uint32_t*>(freem);
    uint32_t* svec32 = reinterpret_cast<uint32_t*>(svec);
    uint16_t* seq = reinterpret_cast<uint16_t*>(sm + L.seq);
    uint16_t* pivcol = reinterpret_cast<uint16_t*>(sm + L.pivcol);
    uint16_t* pivrow = reinterpret_cast<uint16_t*>(sm + L.pivrow);
    uint16_t* slot_of_row = reinterpret_cast<uint16_t*>(sm + L.slot);
    uint32_t* accs = reinterpret_cast<uint32_t*>(sm + L.accs);
    uint32_t* car = reinterpret_cast<uint32_t*>(sm + L.car);
    const int lane = threadIdx.x & 31;
    const int m = w.rows, TS = L.TS;
    const int carryW = (w.carry_rows + 31) / 32;
    const uint32_t* syn = b.syn_buf + static_cast<size_t>(shot) * b.syn_stride32;
    {
        __syncwarp();
        for (int i = lane; i < m; i += 32) {
            slot_of_row[i] = 0xFFFFu;
            if (EXACT) seq[i] = static_cast<uint16_t>(i);
        }
        for (int i = lane; i < 4 * NQ; i += 32) {
            const int left = m - 32 * i;
            freem32[i] = left >= 32 ? 0xFFFFFFFFu : (left > 0 ? (1u << left) - 1u : 0u);
            svec32[i] = i < w.rowsW32 ? syn[i] : 0u;
        }
        for (int i = lane; i < 2 * w.KW; i += 32) accs[i] = 0;
        for (int i = lane; i <= carryW; i += 32) car[i] = 0;
        __syncwarp();

        int rank = 0, base = 0;
        bool done = false;
        while (rank < m && base < n && !done) {
            // ---- next 32 columns in sorted order, reduced by the transformation so far: one per lane
            uint4 cand[NQ];
#pragma unroll
            for (int i = 0; i < NQ; ++i) cand[i] = make_uint4(0u, 0u, 0u, 0u);
            int candcol = -1;
            if (base + lane < n) {
                candcol = order[base + lane];
                const int qb = __ldg(w.cptr + candcol), qe = __ldg(w.cptr + candcol + 1);
                for (int q = qb; q < qe; ++q) {
                    const int row = __ldg(w.crow + q);
                    const uint32_t s = slot_of_row[row];
                    if (s != 0xFFFFu) {
#pragma unroll
                        for (int i = 0; i < NQ; ++i) xor4(cand[i], T4[i * TS + s]);
                    } else {                                     // the row is not a pivot row yet: its column of T is a unit vector
                        const uint32_t bit = 1u << (row & 31);
                        const int c = (row >> 5) & 3;
#pragma unroll
                        for (int i = 0; i < NQ; ++i) {
                            if (i == (row >> 7)) {
                                cand[i].x ^= c == 0 ? bit : 0u;
                                cand[i].y ^= c == 1 ? bit : 0u;
                                cand[i].z ^= c == 2 ? bit : 0u;
                                cand[i].w ^= c == 3 ? bit : 0u;
                            }
                        }
                    }
                }
            }
            base += 32;
            uint32_t live = 0;
#pragma unroll
            for (int i = 0; i < NQ; ++i) live |= and_any(cand[i], freem[i]);
            bool alive = live != 0;

            // ---- consume the panel: every live candidate becomes a pivot, in sorted order
            for (;;) {
                const uint32_t mask = __ballot_sync(kFull, alive);
                if (!mask) break;
                const int pl = __ffs(mask) - 1;
                if (lane == pl) {
#pragma unroll
                    for (int i = 0; i < NQ; ++i) {
                        rvec[i] = cand[i];
                        T4[i * TS + rank] = cand[i];             // column of T of the new pivot row: e_prow ^ r = the candidate itself
                        cand[i] = make_uint4(0u, 0u, 0u, 0u);
                    }
                    alive = false;
                }
                __syncwarp();
                int prow;
                if (!EXACT) {
                    prow = -1;
#pragma unroll
                    for (int w0 = 0; w0 < 4 * NQ; w0 += 32) {             // the first free row that carries the candidate
                        const uint32_t x = w0 + lane < 4 * NQ ? (rvec32[w0 + lane] & freem32[w0 + lane]) : 0u;
                        const uint32_t bal = __ballot_sync(kFull, x != 0u);
                        if (bal && prow < 0) {
                            const int wl = __ffs(bal) - 1;
                            prow = 32 * (w0 + wl) + __shfl_sync(kFull, __ffs(x) - 1, wl);
                        }
                    }
                } else {
                    // rank-deficient window (inconsistent syndromes are possible): follow the oracle's row order exactly,
                    // i.e. the first free row in position order that has a 1
                    int p = -1;
                    for (int c0 = rank; c0 < m; c0 += 32) {
                        const int pos = c0 + lane;
                        uint32_t bit = 0;
                        if (pos < m) {
                            const int row = seq[pos];
                            bit = (rvec32[row >> 5] >> (row & 31)) & 1u;
                        }
                        const uint32_t bb = __ballot_sync(kFull, bit);
                        if (bb) { p = c0 + __ffs(bb) - 1; break; }
                    }
                    prow = seq[p];
                    __syncwarp();
                    if (lane == 0) seq[p] = seq[rank];
                }
                const int wsel = prow >> 5;
                const uint32_t bsel = 1u << (prow & 31);
                const uint32_t sbit = svec32[wsel] & bsel;
                const int pcol = __shfl_sync(kFull, candcol, pl);
                __syncwarp();
                if (lane == 0) {
                    rvec32[wsel] &= ~bsel;                       // r = the candidate without the pivot bit
                    freem32[wsel] &= ~bsel;
                    pivcol[rank] = static_cast<uint16_t>(pcol);
                    pivrow[rank] = static_cast<uint16_t>(prow);
                    if (HI) reinterpret_cast<uint16_t*>(sm + L.pivpos)[rank] = static_cast<uint16_t>(base - 32 + pl);
                    slot_of_row[prow] = static_cast<uint16_t>(rank);
                }
                __syncwarp();
                if (sbit)
                    for (int i = lane; i < 4 * NQ; i += 32) svec32[i] ^= rvec32[i];
                uint4 r[NQ];
#pragma unroll
                for (int i = 0; i < NQ; ++i) r[i] = rvec[i];
                // own candidate
                {
                    uint4 v = cand[0];
#pragma unroll
                    for (int i = 1; i < NQ; ++i)
                        if (i == (wsel >> 2)) v = cand[i];
                    if (comp(v, wsel & 3) & bsel) {
                        uint32_t lv = 0;
#pragma unroll
                        for (int i = 0; i < NQ; ++i) {
                            xor4(cand[i], r[i]);
                            lv |= and_any(cand[i], freem[i]);
                        }
                        alive = lv != 0;
                    }
                }
                // stored columns of T (pivot slots found before this one)
                const int tword = ((wsel >> 2) * TS) * 4 + (wsel & 3);
                for (int s = lane; s < rank; s += 32) {
                    if (T32[tword + 4 * s] & bsel) {
#pragma unroll
                        for (int i = 0; i < NQ; ++i) {
                            uint4 t = T4[i * TS + s];
                            xor4(t, r[i]);
                            T4[i * TS + s] = t;
                        }
                    }
                }
                ++rank;
                __syncwarp();
                if (rank == m) break;
                // Early exit (exact): once the reduced syndrome is zero on every free row, no later pivot can change it
                // (an operation only acts on a vector that has the pivot row's bit, and pivot rows are taken from the
                // free rows), and every later pivot gets solution bit 0.
                if (!HI) {
                    uint32_t y = 0u;
                    for (int i = lane; i < 4 * NQ; i += 32) y |= svec32[i] & freem32[i];
                    if (!__any_sync(kFull, y != 0u)) { done = true; break; }
                }
            }
        }
        if (!done && rank < m && n < n_total) return false;      // ran out of sorted columns: the caller retries with the full order
        // ---- solution on the pivots (reduced syndrome; higher-order OSD may replace it and flip non-pivot columns), commit
        __syncwarp();
        int nflip = 0;
        const uint32_t* flips = reinterpret_cast<const uint32_t*>(sm + L.flips);
        if (HI) {
            osd_higher<NQ>(w, b, sm, L, Tbase, rank, order, n, lane);
            nflip = static_cast<int>(flips[0]);
        }
        auto commit_column = [&](const int j) {
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
        };
        if (HI) {                            // osd_higher left the solution indexed by pivot rank in column order
            const uint32_t* pinfo = reinterpret_cast<const uint32_t*>(sm + L.pinfo);
            for (int t = lane; t < rank; t += 32)
                if ((svec32[t >> 5] >> (t & 31)) & 1u) commit_column(static_cast<int>(pinfo[t] & 0xFFFFu));
        } else {
            for (int rr = lane; rr < rank; rr += 32) {
                const int row = pivrow[rr];
                if ((svec32[row >> 5] >> (row & 31)) & 1u) commit_column(pivcol[rr]);
            }
        }
        for (int f = lane; f < nflip; f += 32) commit_column(static_cast<int>(flips[1 + f]));
        __syncwarp();
        for (int i = lane; i < w.KW; i += 32) {
            const uint64_t v = (static_cast<uint64_t>(accs[2 * i + 1]) << 32) | accs[2 * i];
            b.acc[static_cast<size_t>(shot) * w.KW + i] ^= v;
        }
        for (int i = lane; i < carryW; i += 32) b.carry[static_cast<size_t>(shot) * b.carry_stride32 + i] = car[i];
        if (lane == 0) {
            atomicAdd(&b.stats[2], 1ull);
            atomicAdd(&b.stats[3], static_cast<unsigned long long>(min(base, n)));
            atomicAdd(&b.stats[4], static_cast<unsigned long long>(rank));
            atomicMax(&b.stats[5], static_cast<unsigned long long>(min(base, n)));
        }
    }
    return true;
}

// TG: the accumulated row transformation T (m x m bits) lives in a global slab per warp instead of shared memory -- windows of
// more than 768 checks (BASELINE config 5: 2250 checks => 633 KB), higher-order OSD only (OSD-0 at that size is osd_big_kernel).
template <int NQ, bool EXACT, bool HI, bool TG = false>
__global__ void __launch_bounds__(32) osd_elim_kernel(const WinDev w, const BatchDev b) {
    extern __shared__ __align__(16) unsigned char sm[];
    const ElimLayout L = elim_layout(w, NQ, EXACT, 0, HI, TG);
    uint4* const Tbase = TG ? reinterpret_cast<uint4*>(static_cast<unsigned char*>(b.lsd_scratch) + static_cast<size_t>(blockIdx.x) * b.lsd_slab)
                            : reinterpret_cast<uint4*>(sm + L.T);
    const int lane = threadIdx.x;
    const int count = *b.fail_count;
    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(b.osd_next, 1);
        job = __shfl_sync(kFull, job, 0);
        if (job >= count) break;
        const int shot = b.fail_list[job];
        const uint16_t* order = b.order_alt ? b.order_alt + static_cast<size_t>(shot) * b.llr_stride
                                            : reinterpret_cast<const uint16_t*>(static_cast<const unsigned char*>(b.llr_buf) +
                                                                                static_cast<size_t>(shot) * b.llr_stride * b.llr_esize);
        elim_job<NQ, EXACT, HI>(w, b, sm, L, Tbase, shot, order, w.ncols, w.ncols);
    }
}

// Fast path: the elimination almost always ends within the few dozen least reliable columns (see the early exit), so the
// BP kernel hands over just those (bp.cu select_for_osd: a 32-bin histogram of the posteriors, bins monotone in the LLR with
// width 1/10 of the window's smallest prior LLR, the smallest bin prefix holding >= kOsdSelTarget columns, at most kOsdSelCap)
// and the warp only has to order them: a bitonic sort on (key, index) in shared memory.  Shots whose elimination outruns the
// selection are appended to the overflow list and redone by osd_sort_kernel + osd_elim_kernel on the full posterior vector.
constexpr int kSelCap = kOsdSelCap;

template <typename R, int NQ, bool EXACT>
__global__ void __launch_bounds__(32) osd_fast_kernel(const WinDev w, const BatchDev b) {
    using KeyT = typename std::conditional<sizeof(R) == 4, uint32_t, uint64_t>::type;
    extern __shared__ __align__(16) unsigned char sm[];
    const ElimLayout L = elim_layout(w, NQ, EXACT, kSelCap);
    KeyT* selkey = reinterpret_cast<KeyT*>(sm + L.selkey);
    uint16_t* selidx = reinterpret_cast<uint16_t*>(sm + L.selidx);
    const int lane = threadIdx.x;
    const int n = w.ncols;
    const int count = *b.fail_count;
    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(b.fast_next, 1);
        job = __shfl_sync(kFull, job, 0);
        if (job >= count) break;
        const int shot = b.fail_list[job];
        // selection made by the BP kernel (bp.cu, select_for_osd): tier 1 at the front of the buffer, tier 2 at its back
        const int packed = b.sel_cnt[shot];
        const int S1 = packed & 0xFFFF, S2 = (packed >> 16) & 0xFFFF;
        const KeyT* gkey = reinterpret_cast<const KeyT*>(b.sel_key) + static_cast<size_t>(shot) * kOsdSelCap;
        const uint16_t* gidx = b.sel_idx + static_cast<size_t>(shot) * kOsdSelCap;
        bool finished = false;
        for (int attempt = 0; attempt < 2 && !finished; ++attempt) {
            // attempt 0 orders and eliminates over tier 1 alone; if that runs out, attempt 1 starts over on both tiers
            const int S = attempt == 0 ? S1 : S1 + S2;
            if (S == 0 || S > kOsdSelCap || (attempt == 1 && S2 == 0)) continue;
            for (int i = lane; i < S1; i += 32) { selkey[i] = gkey[i]; selidx[i] = gidx[i]; }
            if (attempt == 1)
                for (int i = lane; i < S2; i += 32) { selkey[S1 + i] = gkey[kOsdSelCap - 1 - i]; selidx[S1 + i] = gidx[kOsdSelCap - 1 - i]; }
            // ---- bitonic sort on (key, column index) in shared memory, padded to a power of two with +inf entries
            int P = 32;
            while (P < S) P <<= 1;
            for (int i = S + lane; i < P; i += 32) { selkey[i] = ~static_cast<KeyT>(0); selidx[i] = 0xFFFFu; }
            __syncwarp();
            for (int k = 2; k <= P; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = lane; t < (P >> 1); t += 32) {
                        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));       // index with bit j clear
                        const int l = i | j;
                        const KeyT ki = selkey[i], kl = selkey[l];
                        const uint16_t xi = selidx[i], xl = selidx[l];
                        const bool gt = ki > kl || (ki == kl && xi > xl);
                        if (gt == ((i & k) == 0)) { selkey[i] = kl; selkey[l] = ki; selidx[i] = xl; selidx[l] = xi; }
                    }
                    __syncwarp();
                }
            }
            finished = elim_job<NQ, EXACT>(w, b, sm, L, reinterpret_cast<uint4*>(sm + L.T), shot, selidx, S, n);
        }
        if (!finished && lane == 0) {
            b.ovf_list[atomicAdd(b.ovf_count, 1)] = shot;
            atomicAdd(&b.stats[6], 1ull);
        }
    }
}

inline int elim_nq(const WinDev& w) { return (w.rows + 127) / 128; }

template <bool HI, typename F>
inline cudaError_t elim_dispatch_h(const WinDev& w, F&& f) {
    const bool exact = !w.full_row_rank;
    switch (elim_nq(w)) {
    case 1: return exact ? f(osd_elim_kernel<1, true, HI>) : f(osd_elim_kernel<1, false, HI>);
    case 2: return exact ? f(osd_elim_kernel<2, true, HI>) : f(osd_elim_kernel<2, false, HI>);
    case 3: return exact ? f(osd_elim_kernel<3, true, HI>) : f(osd_elim_kernel<3, false, HI>);
    case 4: return exact ? f(osd_elim_kernel<4, true, HI>) : f(osd_elim_kernel<4, false, HI>);
    case 5: return exact ? f(osd_elim_kernel<5, true, HI>) : f(osd_elim_kernel<5, false, HI>);
    case 6: return exact ? f(osd_elim_kernel<6, true, HI>) : f(osd_elim_kernel<6, false, HI>);
    }
    return cudaErrorInvalidValue;
}

// tall windows (768 < checks <= 2304), higher-order OSD: T in the global slab, NQ in steps of three 128-row groups
template <typename F>
inline cudaError_t elim_dispatch_tall(const WinDev& w, F&& f) {
    const bool exact = !w.full_row_rank;
    const int nq = elim_nq(w);
    if (nq <= 9) return exact ? f(osd_elim_kernel<9, true, true, true>) : f(osd_elim_kernel<9, false, true, true>);
    if (nq <= 12) return exact ? f(osd_elim_kernel<12, true, true, true>) : f(osd_elim_kernel<12, false, true, true>);
    if (nq <= 15) return exact ? f(osd_elim_kernel<15, true, true, true>) : f(osd_elim_kernel<15, false, true, true>);
    if (nq <= 18) return exact ? f(osd_elim_kernel<18, true, true, true>) : f(osd_elim_kernel<18, false, true, true>);
    return cudaErrorInvalidValue;
}
inline int elim_tall_nq(const WinDev& w) { const int nq = elim_nq(w); return nq <= 9 ? 9 : (nq <= 12 ? 12 : (nq <= 15 ? 15 : 18)); }

template <typename F>
inline cudaError_t elim_dispatch(const WinDev& w, bool hi, F&& f) {
    if (hi && w.rows > 768) return elim_dispatch_tall(w, f);
    return hi ? elim_dispatch_h<true>(w, f) : elim_dispatch_h<false>(w, f);
}

template <typename R, typename F>
inline cudaError_t fast_dispatch_r(const WinDev& w, F&& f) {
    const bool exact = !w.full_row_rank;
    switch (elim_nq(w)) {
    case 1: return exact ? f(osd_fast_kernel<R, 1, true>) : f(osd_fast_kernel<R, 1, false>);
    case 2: return exact ? f(osd_fast_kernel<R, 2, true>) : f(osd_fast_kernel<R, 2, false>);
    case 3: return exact ? f(osd_fast_kernel<R, 3, true>) : f(osd_fast_kernel<R, 3, false>);
    case 4: return exact ? f(osd_fast_kernel<R, 4, true>) : f(osd_fast_kernel<R, 4, false>);
    case 5: return exact ? f(osd_fast_kernel<R, 5, true>) : f(osd_fast_kernel<R, 5, false>);
    case 6: return exact ? f(osd_fast_kernel<R, 6, true>) : f(osd_fast_kernel<R, 6, false>);
    }
    return cudaErrorInvalidValue;
}

template <typename F>
inline cudaError_t fast_dispatch(const WinDev& w, int precision, F&& f) {
    return precision == 32 ? fast_dispatch_r<float>(w, f) : fast_dispatch_r<double>(w, f);
}

}  // namespace

size_t osd_sort_smem_bytes(const WinDev& w, int precision) { return sort_layout(w, precision == 32 ? 4 : 8).total; }
size_t osd_sort_slab_bytes(const WinDev& w, int precision) { return sort_layout(w, precision == 32 ? 4 : 8).slab; }
size_t osd_elim_smem_bytes(const WinDev& w, bool hi) {
    if (hi && w.rows > 768) return elim_layout(w, elim_tall_nq(w), !w.full_row_rank, 0, true, true).total;
    return elim_layout(w, elim_nq(w), !w.full_row_rank, 0, hi).total;
}
// higher-order OSD on a window taller than the shared-memory elimination takes: T in a global slab of this many bytes per warp
bool osd_tall_hi_supported(const WinDev& w, int precision) {
    return w.rows > 768 && w.rows <= 2304 && w.ncols <= 65535 && osd_elim_smem_bytes(w, true) <= 200 * 1024;
}
size_t osd_tall_hi_slab_bytes(const WinDev& w) { return static_cast<size_t>(elim_tall_nq(w)) * ((w.rows + 31) / 32 * 32) * 16; }
cudaError_t osd_tall_hi_configure(const WinDev& w, int precision) {
    cudaError_t se = osd_sort_configure(w, precision);
    if (se != cudaSuccess) return se;
    const size_t es = osd_elim_smem_bytes(w, true);
    return elim_dispatch_tall(w, [&](auto kern) {
        return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(es));
    });
}
size_t osd_fast_smem_bytes(const WinDev& w) { return elim_layout(w, elim_nq(w), !w.full_row_rank, kSelCap).total; }

bool osd_supported(const WinDev& w, int precision) {
    return w.rows <= 768 && w.ncols <= 65535 && osd_sort_smem_bytes(w, precision) <= 220 * 1024 && osd_fast_smem_bytes(w) <= 220 * 1024 &&
           osd_elim_smem_bytes(w, true) <= 220 * 1024;
}

cudaError_t osd_sort_configure(const WinDev& w, int precision) {
    // several windows may share one instantiation: the attributes only ever grow
    static size_t sort_have_d[kMaxDevices][2] = {};
    const size_t ss = osd_sort_smem_bytes(w, precision);
    size_t& sh = sort_have_d[device_slot()][precision == 32 ? 0 : 1];
    if (ss > sh) {
        cudaError_t e = precision == 32
            ? cudaFuncSetAttribute(osd_sort_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ss))
            : cudaFuncSetAttribute(osd_sort_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(ss));
        if (e != cudaSuccess) return e;
        sh = ss;
    }
    return cudaSuccess;
}

cudaError_t osd_configure(const WinDev& w, int precision) {
    static size_t elim_have_d[kMaxDevices][2][2][8] = {}, fast_have_d[kMaxDevices][2][2][8] = {};
    const int dev = device_slot();
    auto& elim_have = elim_have_d[dev];
    auto& fast_have = fast_have_d[dev];
    const size_t fs = osd_fast_smem_bytes(w);
    cudaError_t se = osd_sort_configure(w, precision);
    if (se != cudaSuccess) return se;
    for (int hi = 0; hi < 2; ++hi) {
        const size_t es = osd_elim_smem_bytes(w, hi != 0);
        size_t& eh = elim_have[hi][w.full_row_rank ? 0 : 1][elim_nq(w) & 7];
        if (es > eh) {
            cudaError_t e = elim_dispatch(w, hi != 0, [&](auto kern) {
                return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(es));
            });
            if (e != cudaSuccess) return e;
            eh = es;
        }
    }
    size_t& fh = fast_have[precision == 32 ? 0 : 1][w.full_row_rank ? 0 : 1][elim_nq(w) & 7];
    if (fs > fh) {
        cudaError_t e = fast_dispatch(w, precision, [&](auto kern) {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fs));
        });
        if (e != cudaSuccess) return e;
        fh = fs;
    }
    return cudaSuccess;
}

cudaError_t launch_osd_fast(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const size_t smem = osd_fast_smem_bytes(w);
    return fast_dispatch(w, precision, [&](auto kern) {
        kern<<<grid, 32, smem, st>>>(w, b);
        return cudaGetLastError();
    });
}

cudaError_t launch_osd_sort(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const size_t smem = osd_sort_smem_bytes(w, precision);
    if (precision == 32) osd_sort_kernel<float><<<grid, kSortThreads, smem, st>>>(w, b);
    else osd_sort_kernel<double><<<grid, kSortThreads, smem, st>>>(w, b);
    return cudaGetLastError();
}

cudaError_t launch_osd_elim(const WinDev& w, const BatchDev& b, bool hi, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const size_t smem = osd_elim_smem_bytes(w, hi);
    return elim_dispatch(w, hi, [&](auto kern) {
        kern<<<grid, 32, smem, st>>>(w, b);
        return cudaGetLastError();
    });
}

}  // namespace qb
