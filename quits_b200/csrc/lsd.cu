// quits_b200/csrc/lsd.cu -- K4L: localized statistics decoding (order 0, and the per-cluster lsd_cs / lsd_e candidate sweep beyond it):
// the post-processing stage of ldpc's BpLsdDecoder
// (reference src/quits/decoder/bplsd.py:38-50,74-86 constructs it; sliding_window.py:171,182 calls decode()).  sm_100a.
//
// The algorithm is the one oracle/cref.c lsd_decode restates (Hillmann et al., "Localized statistics decoding"):
// one cluster per unsatisfied check; every invalid cluster, smallest first, grows by the least reliable bit (smallest BP
// posterior, ties by index) next to its boundary checks; clusters that meet on a check merge, the one with fewer bits into the
// other; a cluster is valid once its syndrome lies in the span of its columns (on-the-fly elimination: the survivor of a merge
// keeps its reduction, the columns of the absorbed cluster are reduced again behind it); the answer of a cluster is the
// solution supported on the first independent columns of its column list.
//
// Mapping: ONE WARP per failed shot (persistent grid pulling shot indices from the list the BP kernel wrote), because the
// algorithm is a serial walk over small data-dependent structures; the lanes share the work inside a step:
//   * GF(2) vectors over the window's checks are NW 32-bit words per lane (NW = 1 up to 1024 checks -- every BB / HGP window --,
//     2 or 3 up to 3072; word i of a vector sits in lane i % 32, slot i / 32): the reduced syndrome z, the pivot-row
//     mask P, the boundary mask B live in registers; clusters are disjoint in their checks, so ONE z / P / B serves all of them
//   * the row operations of all clusters sit in one array in creation order (pivot row in shared memory, 128-byte vector in
//     an L2-resident scratch slab); reducing a new column tests 32 operations per step (one per lane) and applies the hits in
//     order -- an operation of another cluster can never hit, its pivot row is not a check of this cluster.  The operations of
//     an absorbed cluster are marked dead; the array is compacted when it fills up
//   * cluster membership: u16 owner per check (shared memory) and per bit (scratch slab), singly linked lists for the checks
//     and the bits (= column order) of a cluster, so that a merge is a relabelling walk over the smaller side plus a splice
//   * the growth candidates of a cluster are found by the lanes striding over the CSR rows of its boundary checks
// Integer/bit work plus comparisons of posteriors the BP kernel left in HBM: results are bit-exact with the oracle.
#include <cstdint>

#include "qb_device.h"

namespace qb {
namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kNone = 0xFFFFu;
constexpr uint32_t kDead = 0xFFFFu;
constexpr uint32_t kDirty = 0xFFFEu;      // row cache: best candidate not computed / no longer free
constexpr uint32_t kActive = 1u, kValid = 2u;
constexpr int kLsdColBytes = 12;          // slab bytes per column: owner, successor + the four u16 arrays of the higher-order sweep

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) / 16 * 16; }

struct LsdLayout {
    size_t v, rbkey, rbcol, dlist, cown, cnext, chead, ctail, bhead, btail, bue, nbits, flag, inv, inv2, oppiv, opcol, ml, accs, car, total;
    int opcap;
};

// capacity of the operation array: the live operations have distinct pivot rows (<= rows), the dead ones go at a compaction
__host__ __device__ inline int lsd_opcap(const int rows) { return (rows + 31) / 32 * 32 + 8; }

__host__ __device__ inline LsdLayout lsd_layout(const WinDev& w) {
    LsdLayout L;
    const size_t m = static_cast<size_t>(w.rows);
    L.opcap = lsd_opcap(w.rows);
    size_t o = 0;
    L.v = o; o += 128 * 3;
    L.rbkey = o; o += al16(m * 8);
    L.rbcol = o; o += al16(m * 2);
    L.dlist = o; o += al16(m * 2);
    L.cown = o; o += al16(m * 2);
    L.cnext = o; o += al16(m * 2);
    L.chead = o; o += al16(m * 2);
    L.ctail = o; o += al16(m * 2);
    L.bhead = o; o += al16(m * 2);
    L.btail = o; o += al16(m * 2);
    L.bue = o; o += al16(m * 2);
    L.nbits = o; o += al16(m * 2);
    L.flag = o; o += al16(m);
    L.inv = o; o += al16(m * 2);
    L.inv2 = o; o += al16(m * 2);
    L.oppiv = o; o += al16(static_cast<size_t>(L.opcap) * 2);
    L.opcol = o; o += al16(static_cast<size_t>(L.opcap) * 2);
    L.ml = o; o += 64;
    L.accs = o; o += al16(static_cast<size_t>(2 * w.KW) * 4);
    L.car = o; o += al16(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4);
    L.total = o;
    return L;
}

__device__ __forceinline__ uint64_t lsd_key(float f) {
    const uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? static_cast<uint32_t>(~u) : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t lsd_key(double f) {
    const uint64_t u = static_cast<uint64_t>(__double_as_longlong(f + 0.0));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// word i of a vector over the checks: lane i % 32, slot i / 32
template <int NW>
__device__ __forceinline__ uint32_t vec_bit(const uint32_t (&v)[NW], const int r) {
    const int wi = r >> 5, slot = wi >> 5;
    uint32_t word = v[0];
#pragma unroll
    for (int s = 1; s < NW; ++s) word = slot == s ? v[s] : word;
    return (__shfl_sync(kFull, word, wi & 31) >> (r & 31)) & 1u;
}
template <int NW>
__device__ __forceinline__ void vec_set(uint32_t (&v)[NW], const int r, const int lane, const bool on) {
    const int wi = r >> 5, slot = wi >> 5;
    if (lane != (wi & 31)) return;
    const uint32_t bit = 1u << (r & 31);
#pragma unroll
    for (int s = 0; s < NW; ++s)
        if (slot == s) v[s] = on ? (v[s] | bit) : (v[s] & ~bit);
}

template <int NW>
__device__ __forceinline__ void vec_xor(uint32_t (&v)[NW], const int r, const int lane) {
    const int wi = r >> 5, slot = wi >> 5;
    if (lane != (wi & 31)) return;
    const uint32_t bit = 1u << (r & 31);
#pragma unroll
    for (int s = 0; s < NW; ++s)
        if (slot == s) v[s] ^= bit;
}

// HI: lsd_order > 0 -- after the growth every cluster runs the candidate sweep oracle/cref.c lsd_cluster_higher states (lsd_e /
// lsd_cs over the cluster's non-pivot columns, weights summed in column-list order with __dadd_rn: the same additions in the same
// order as the oracle, so ties resolve identically)
template <typename R, int NW, bool HI>
__global__ void __launch_bounds__(32) lsd_kernel(const WinDev w, const BatchDev b) {
    extern __shared__ __align__(16) unsigned char sm[];
    const LsdLayout L = lsd_layout(w);
    uint32_t* vsm = reinterpret_cast<uint32_t*>(sm + L.v);
    uint64_t* rbkey = reinterpret_cast<uint64_t*>(sm + L.rbkey);
    uint16_t* rbcol = reinterpret_cast<uint16_t*>(sm + L.rbcol);
    uint16_t* dlist = reinterpret_cast<uint16_t*>(sm + L.dlist);
    uint16_t* cown = reinterpret_cast<uint16_t*>(sm + L.cown);
    uint16_t* cnext = reinterpret_cast<uint16_t*>(sm + L.cnext);
    uint16_t* chead = reinterpret_cast<uint16_t*>(sm + L.chead);
    uint16_t* ctail = reinterpret_cast<uint16_t*>(sm + L.ctail);
    uint16_t* bhead = reinterpret_cast<uint16_t*>(sm + L.bhead);
    uint16_t* btail = reinterpret_cast<uint16_t*>(sm + L.btail);
    uint16_t* bue = reinterpret_cast<uint16_t*>(sm + L.bue);
    uint16_t* nbits = reinterpret_cast<uint16_t*>(sm + L.nbits);
    uint8_t* flag = reinterpret_cast<uint8_t*>(sm + L.flag);
    uint16_t* inv = reinterpret_cast<uint16_t*>(sm + L.inv);
    uint16_t* inv2 = reinterpret_cast<uint16_t*>(sm + L.inv2);
    uint16_t* oppiv = reinterpret_cast<uint16_t*>(sm + L.oppiv);
    uint16_t* opcol = reinterpret_cast<uint16_t*>(sm + L.opcol);
    uint16_t* ml = reinterpret_cast<uint16_t*>(sm + L.ml);
    uint32_t* accs = reinterpret_cast<uint32_t*>(sm + L.accs);
    uint32_t* car = reinterpret_cast<uint32_t*>(sm + L.car);
    const int lane = threadIdx.x;
    const int m = w.rows;
    const int carryW = (w.carry_rows + 31) / 32;
    const int opcap = L.opcap;
    // per-CTA scratch slab: bit owner, bit successor (column order of a cluster), operation vectors
    unsigned char* slab = static_cast<unsigned char*>(b.lsd_scratch) + static_cast<size_t>(blockIdx.x) * b.lsd_slab;
    // (the owner array sits at the same place for every window of the decoder: b.lsd_cols = widest window)
    uint16_t* bown = reinterpret_cast<uint16_t*>(slab);                                   // 0xFFFF between shots
    uint16_t* bnext = bown + b.lsd_cols;
    // higher order only: the cluster being swept as arrays -- its columns in list order, pivot row (or none) and non-pivot rank
    // of each position, position of a column
    uint16_t* clist = bnext + b.lsd_cols;
    uint16_t* cprow = clist + b.lsd_cols;
    uint16_t* crank = cprow + b.lsd_cols;
    uint16_t* bpos = crank + b.lsd_cols;
    uint32_t* opvec = reinterpret_cast<uint32_t*>(slab + al16(static_cast<size_t>(b.lsd_cols) * kLsdColBytes));      // [opcap][NW][32]
    const int count = *b.fail_count;
    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(b.osd_next, 1);
        job = __shfl_sync(kFull, job, 0);
        if (job >= count) break;
        const int shot = b.fail_list[job];
        const uint32_t* syn = b.syn_buf + static_cast<size_t>(shot) * b.syn_stride32;
        const R* llr = reinterpret_cast<const R*>(b.llr_buf) + static_cast<size_t>(shot) * b.llr_stride;
        __syncwarp();
        // ---- initial clusters: one per unsatisfied check, ascending
        uint32_t Sw[NW], zw[NW], Pw[NW], Bw[NW];          // raw syndrome, reduced syndrome, pivot rows, boundary checks
#pragma unroll
        for (int s = 0; s < NW; ++s) {
            const int wi = 32 * s + lane, left = m - 32 * wi;
            Sw[s] = (wi < w.rowsW32 ? syn[wi] : 0u) & (left >= 32 ? kFull : (left > 0 ? (1u << left) - 1u : 0u));
            zw[s] = Sw[s]; Pw[s] = 0u; Bw[s] = Sw[s];
        }
        for (int i = lane; i < m; i += 32) { cown[i] = static_cast<uint16_t>(kNone); rbcol[i] = static_cast<uint16_t>(kDirty); }
        for (int i = lane; i < 2 * w.KW; i += 32) accs[i] = 0;
        for (int i = lane; i <= carryW; i += 32) car[i] = 0;
        int nc = 0;
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NW; ++s) {
            const int mine = __popc(Sw[s]);
            int incl = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += t;
            }
            int id = nc + incl - mine;
            nc += __shfl_sync(kFull, incl, 31);
            for (uint32_t x = Sw[s]; x; x &= x - 1, ++id) {
                const int r = (32 * s + lane) * 32 + __ffs(x) - 1;
                cown[r] = static_cast<uint16_t>(id);
                cnext[r] = static_cast<uint16_t>(kNone);
                chead[id] = ctail[id] = static_cast<uint16_t>(r);
                bhead[id] = btail[id] = bue[id] = static_cast<uint16_t>(kNone);
                nbits[id] = 0;
                flag[id] = static_cast<uint8_t>(kActive);
                inv[id] = static_cast<uint16_t>(id);
            }
        }
        __syncwarp();
        int ninv = nc, nops = 0;
        unsigned long long grown = 0;
        while (ninv > 0) {
            for (int t = 0; t < ninv; ++t) {
                const int cid = inv[t];
                if (!(flag[cid] & kActive)) continue;
                // ---- growth candidates: the free bits next to the cluster's boundary checks.  (Every bit of a cluster has all its
                // checks in the cluster, so a bit next to one of this cluster's checks is either in this cluster or in none.)  The best
                // free bit of a row is cached; it stays the best until that very bit joins a cluster.
                uint64_t bkey = ~0ull;
                uint32_t bcol = kNone;
                int nd = 0;
                for (int i = 0; i < w.rowsW32; ++i) {
                    uint32_t bsrc = Bw[0];
#pragma unroll
                    for (int s = 1; s < NW; ++s) bsrc = (i >> 5) == s ? Bw[s] : bsrc;
                    const uint32_t bw = __shfl_sync(kFull, bsrc, i & 31);
                    const int r = 32 * i + lane;
                    const bool mine = ((bw >> lane) & 1u) && cown[r] == cid;
                    const uint32_t c = mine ? rbcol[r] : kNone;
                    const bool dirty = mine && c == kDirty;
                    if (mine && !dirty) {
                        const uint64_t k = rbkey[r];
                        if (k < bkey || (k == bkey && c < bcol)) { bkey = k; bcol = c; }
                    }
                    const uint32_t dm = __ballot_sync(kFull, dirty);
                    if (dirty) dlist[nd + __popc(dm & ((1u << lane) - 1u))] = static_cast<uint16_t>(r);
                    nd += __popc(dm);
                }
                __syncwarp();
                for (int d0 = 0; d0 < nd; d0 += 4) {                // rescan the rows without a cached candidate, four at a time
                    uint32_t rr[4], cc[4][2];
                    int e0[4], e1[4];
                    uint64_t kk[4];
                    uint32_t kc[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        rr[k] = d0 + k < nd ? dlist[d0 + k] : kNone;
                        e0[k] = rr[k] != kNone ? __ldg(w.rptr + rr[k]) : 0;
                        e1[k] = rr[k] != kNone ? __ldg(w.rptr + rr[k] + 1) : 0;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int q = 0; q < 2; ++q) cc[k][q] = e0[k] + lane + 32 * q < e1[k] ? __ldg(w.rcol + e0[k] + lane + 32 * q) : kNone;
                    uint32_t ow[4][2];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int q = 0; q < 2; ++q) ow[k][q] = cc[k][q] != kNone ? bown[cc[k][q]] : 0u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        kk[k] = ~0ull; kc[k] = kNone;
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            if (cc[k][q] != kNone && ow[k][q] == kNone) {
                                const uint64_t key = lsd_key(llr[cc[k][q]]);
                                if (key < kk[k] || (key == kk[k] && cc[k][q] < kc[k])) { kk[k] = key; kc[k] = cc[k][q]; }
                            }
                        }
                        for (int e = e0[k] + lane + 64; e < e1[k]; e += 32) {          // rows longer than 64 entries
                            const uint32_t c = __ldg(w.rcol + e);
                            if (bown[c] != kNone) continue;
                            const uint64_t key = lsd_key(llr[c]);
                            if (key < kk[k] || (key == kk[k] && c < kc[k])) { kk[k] = key; kc[k] = c; }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) {
                            const uint64_t ok = __shfl_xor_sync(kFull, kk[k], d);
                            const uint32_t oc = __shfl_xor_sync(kFull, kc[k], d);
                            if (ok < kk[k] || (ok == kk[k] && oc < kc[k])) { kk[k] = ok; kc[k] = oc; }
                        }
                        if (rr[k] == kNone) continue;
                        if (lane == 0) { rbkey[rr[k]] = kk[k]; rbcol[rr[k]] = static_cast<uint16_t>(kc[k]); }
                        if (kc[k] == kNone) {                        // no free bit left next to this check: it leaves the boundary
                            vec_set<NW>(Bw, static_cast<int>(rr[k]), lane, false);
                        } else if (kk[k] < bkey || (kk[k] == bkey && kc[k] < bcol)) { bkey = kk[k]; bcol = kc[k]; }
                    }
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) {
                    const uint64_t ok = __shfl_xor_sync(kFull, bkey, d);
                    const uint32_t oc = __shfl_xor_sync(kFull, bcol, d);
                    if (ok < bkey || (ok == bkey && oc < bcol)) { bkey = ok; bcol = oc; }
                }
                if (bcol == kNone) {                               // nothing left to add: stop growing (syndrome outside the image)
                    if (lane == 0) flag[cid] |= static_cast<uint8_t>(kValid);
                    __syncwarp();
                    continue;
                }
                ++grown;
                // ---- the bit joins the cluster; its checks join too, checks of other clusters are collisions
                const int best = static_cast<int>(bcol);
                const int q0 = __ldg(w.cptr + best), wt = __ldg(w.cptr + best + 1) - q0;
                int nm = 0;
                if (lane == 0) {
                    bown[best] = static_cast<uint16_t>(cid);
                    bnext[best] = static_cast<uint16_t>(kNone);
                    if (btail[cid] == kNone) bhead[cid] = static_cast<uint16_t>(best);
                    else bnext[btail[cid]] = static_cast<uint16_t>(best);
                    btail[cid] = static_cast<uint16_t>(best);
                    if (bue[cid] == kNone) bue[cid] = static_cast<uint16_t>(best);
                    nbits[cid] = static_cast<uint16_t>(nbits[cid] + 1);
                }
                __syncwarp();
                for (int q = 0; q < wt; ++q) {
                    const uint32_t r = __ldg(w.crow + q0 + q);
                    const uint32_t o = cown[r];
                    const bool stale = rbcol[r] == static_cast<uint32_t>(best);
                    __syncwarp();                                    // every lane has read the owner before lane 0 changes it
                    if (lane == 0 && stale) rbcol[r] = static_cast<uint16_t>(kDirty);
                    if (o == static_cast<uint32_t>(cid)) continue;
                    if (o == kNone) {
                        if (lane == 0) {
                            cown[r] = static_cast<uint16_t>(cid);
                            cnext[r] = static_cast<uint16_t>(kNone);
                            cnext[ctail[cid]] = static_cast<uint16_t>(r);
                            ctail[cid] = static_cast<uint16_t>(r);
                        }
                        vec_set<NW>(Bw, static_cast<int>(r), lane, true);
                        __syncwarp();
                        continue;
                    }
                    bool seen = false;
                    for (int k = 0; k < nm; ++k) seen |= ml[k] == o;
                    if (!seen) {
                        if (lane == 0) ml[nm] = static_cast<uint16_t>(o);
                        ++nm;
                        __syncwarp();
                    }
                }
                // ---- merges, in the order the collisions were met: fewer bits into more bits, ties into the growing side
                int big = cid;
                for (int k = 0; k < nm; ++k) {
                    const int o = ml[k];
                    int small;
                    if (nbits[big] < nbits[o]) { small = big; big = o; } else { small = o; }
                    // the absorbed cluster's reduction is discarded: its operations die, its checks go back to the raw syndrome
                    if (nbits[small]) {
                        for (int i = lane; i < nops; i += 32) {
                            const uint32_t p = oppiv[i];
                            if (p != kDead && cown[p] == small) oppiv[i] = static_cast<uint16_t>(kDead);
                        }
                    }
                    __syncwarp();
                    for (uint32_t r = chead[small]; r != kNone; r = cnext[r]) {
                        if (lane == 0) cown[r] = static_cast<uint16_t>(big);
                        if (lane == static_cast<int>((r >> 5) & 31u)) {
                            const uint32_t bit = 1u << (r & 31);
#pragma unroll
                            for (int s = 0; s < NW; ++s)
                                if (static_cast<int>(r >> 10) == s) { zw[s] = (zw[s] & ~bit) | (Sw[s] & bit); Pw[s] &= ~bit; }
                        }
                    }
                    for (uint32_t j = bhead[small]; j != kNone; j = bnext[j])
                        if (lane == 0) bown[j] = static_cast<uint16_t>(big);
                    if (lane == 0) {
                        cnext[ctail[big]] = chead[small];
                        ctail[big] = ctail[small];
                        if (bhead[small] != kNone) {
                            if (btail[big] == kNone) bhead[big] = bhead[small];
                            else bnext[btail[big]] = bhead[small];
                            btail[big] = btail[small];
                            if (bue[big] == kNone) bue[big] = bhead[small];
                        }
                        nbits[big] = static_cast<uint16_t>(nbits[big] + nbits[small]);
                        flag[small] = 0;
                    }
                    __syncwarp();
                }
                // ---- on-the-fly elimination of the survivor's new columns, in column-list order
                for (uint32_t j = bue[big]; j != kNone; j = bnext[j]) {
                    uint32_t vw[NW];
#pragma unroll
                    for (int s = 0; s < NW; ++s) vw[s] = 0u;
                    {
                        const int c0 = __ldg(w.cptr + j), c1 = __ldg(w.cptr + j + 1);
                        for (int q = c0; q < c1; ++q) vec_set<NW>(vw, static_cast<int>(__ldg(w.crow + q)), lane, true);
                    }
                    for (int base = 0; base < nops; base += 32) {
                        const uint32_t p = base + lane < nops ? oppiv[base + lane] : kDead;
                        uint32_t todo = kFull;
                        for (;;) {
                            __syncwarp();
#pragma unroll
                            for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = vw[s];
                            __syncwarp();
                            const bool hit = p != kDead && ((vsm[p >> 5] >> (p & 31)) & 1u);
                            const uint32_t mask = __ballot_sync(kFull, hit) & todo;
                            if (!mask) break;
                            const int i = __ffs(mask) - 1;
#pragma unroll
                            for (int s = 0; s < NW; ++s) vw[s] ^= opvec[(static_cast<size_t>(base + i) * NW + s) * 32 + lane];
                            todo = i == 31 ? 0u : (kFull << (i + 1));
                            if (!todo) break;
                        }
                    }
                    // pivot row: first check of the reduced column that is not a pivot row yet
                    int p = -1;
#pragma unroll
                    for (int s = 0; s < NW; ++s) {
                        const uint32_t freew = vw[s] & ~Pw[s];
                        const uint32_t fm = __ballot_sync(kFull, freew != 0u);
                        if (p < 0 && fm) {
                            const int pl = __ffs(fm) - 1;
                            p = (32 * s + pl) * 32 + __ffs(__shfl_sync(kFull, freew, pl)) - 1;
                        }
                    }
                    if (p < 0) continue;                             // dependent on the columns before it
                    if (nops == opcap) {                             // drop the dead operations, keeping the order
                        int wr = 0;
                        for (int i = 0; i < nops; ++i) {
                            const uint32_t pi = oppiv[i];
                            if (pi == kDead) continue;
                            if (wr != i) {
                                uint32_t x[NW];
#pragma unroll
                                for (int s = 0; s < NW; ++s) x[s] = opvec[(static_cast<size_t>(i) * NW + s) * 32 + lane];
                                const uint32_t ci = opcol[i];
                                __syncwarp();
#pragma unroll
                                for (int s = 0; s < NW; ++s) opvec[(static_cast<size_t>(wr) * NW + s) * 32 + lane] = x[s];
                                if (lane == 0) { oppiv[wr] = static_cast<uint16_t>(pi); opcol[wr] = static_cast<uint16_t>(ci); }
                                __syncwarp();
                            }
                            ++wr;
                        }
                        nops = wr;
                    }
                    vec_set<NW>(vw, p, lane, false);                 // the row operation leaves the pivot row alone
                    vec_set<NW>(Pw, p, lane, true);
#pragma unroll
                    for (int s = 0; s < NW; ++s) opvec[(static_cast<size_t>(nops) * NW + s) * 32 + lane] = vw[s];
                    if (lane == 0) { oppiv[nops] = static_cast<uint16_t>(p); opcol[nops] = static_cast<uint16_t>(j); }
                    ++nops;
                    if (vec_bit<NW>(zw, p)) {
#pragma unroll
                        for (int s = 0; s < NW; ++s) zw[s] ^= vw[s];
                    }
                    __syncwarp();
                }
                // ---- valid when the reduced syndrome vanishes on the cluster's non-pivot checks
                bool bad = false;
#pragma unroll
                for (int s = 0; s < NW; ++s)
                    for (uint32_t x = zw[s] & ~Pw[s]; x; x &= x - 1) bad |= cown[(32 * s + lane) * 32 + __ffs(x) - 1] == big;
                const bool invalid = __any_sync(kFull, bad);
                if (lane == 0) {
                    bue[big] = static_cast<uint16_t>(kNone);
                    flag[big] = static_cast<uint8_t>(kActive | (invalid ? 0u : kValid));
                }
                __syncwarp();
            }
            // ---- next round: the active invalid clusters, fewest bits first (stable in the cluster id)
            int k = 0;
            for (int base = 0; base < nc; base += 32) {
                const int id = base + lane;
                const bool in = id < nc && (flag[id] & (kActive | kValid)) == kActive;
                const uint32_t mask = __ballot_sync(kFull, in);
                if (in) inv2[k + __popc(mask & ((1u << lane) - 1u))] = static_cast<uint16_t>(id);
                k += __popc(mask);
            }
            __syncwarp();
            for (int a = lane; a < k; a += 32) {
                const uint32_t ida = inv2[a], na = nbits[ida];
                int rank = 0;
                for (int c = 0; c < k; ++c) {
                    const uint32_t idc = inv2[c], ncb = nbits[idc];
                    rank += (ncb < na || (ncb == na && idc < ida)) ? 1 : 0;
                }
                inv[rank] = static_cast<uint16_t>(ida);
            }
            ninv = k;
            __syncwarp();
        }
        if (HI) {
            // ---- higher order: candidate sweep inside every cluster; the winner's non-pivot columns are committed here, its pivot
            // part by folding the reduced candidate vector into z
            const double* const wt = w.osd_wt;
            const int ord = b.lsd_order;
            auto reduce_ops = [&](uint32_t (&vw)[NW]) {
                for (int base = 0; base < nops; base += 32) {
                    const uint32_t p = base + lane < nops ? oppiv[base + lane] : kDead;
                    uint32_t todo = kFull;
                    for (;;) {
                        __syncwarp();
#pragma unroll
                        for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = vw[s];
                        __syncwarp();
                        const bool hit = p != kDead && ((vsm[p >> 5] >> (p & 31)) & 1u);
                        const uint32_t mask = __ballot_sync(kFull, hit) & todo;
                        if (!mask) break;
                        const int i = __ffs(mask) - 1;
#pragma unroll
                        for (int s = 0; s < NW; ++s) vw[s] ^= opvec[(static_cast<size_t>(base + i) * NW + s) * 32 + lane];
                        todo = i == 31 ? 0u : (kFull << (i + 1));
                        if (!todo) break;
                    }
                }
            };
            auto commit_col = [&](const int j) {                 // one lane
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
            for (int id = 0; id < nc; ++id) {
                if (!(flag[id] & kActive) || nbits[id] == 0) continue;
                const int nb = nbits[id];
                __syncwarp();
                if (lane == 0) {
                    int k = 0;
                    for (uint32_t j = bhead[id]; j != kNone; j = bnext[j], ++k) {
                        clist[k] = static_cast<uint16_t>(j); bpos[j] = static_cast<uint16_t>(k); cprow[k] = static_cast<uint16_t>(kNone);
                    }
                }
                __syncwarp();
                for (int i = lane; i < nops; i += 32) {
                    const uint32_t p = oppiv[i];
                    if (p != kDead && cown[p] == id) cprow[bpos[opcol[i]]] = static_cast<uint16_t>(p);
                }
                __syncwarp();
                int nnp = 0;                                     // non-pivot ranks; the first 32 non-pivot positions go to ml[]
                for (int base = 0; base < nb; base += 32) {
                    const int k = base + lane;
                    const bool np = k < nb && cprow[k] == kNone;
                    const uint32_t mk = __ballot_sync(kFull, np);
                    const int rk = nnp + __popc(mk & ((1u << lane) - 1u));
                    if (k < nb) crank[k] = static_cast<uint16_t>(np ? rk : 0xFFFF);
                    if (np && rk < 32) ml[rk] = static_cast<uint16_t>(k);
                    nnp += __popc(mk);
                }
                __syncwarp();
                if (nnp == 0) continue;
                const int wsub = ord < nnp ? ord : nnp;         // candidates beyond singles range over the first wsub non-pivots (<= 32)
                // candidate = one position (or -1) plus a pattern over the first 32 non-pivot ranks
                auto cand_vec = [&](const int single, const uint32_t pat, uint32_t (&vw)[NW]) {
#pragma unroll
                    for (int s = 0; s < NW; ++s) vw[s] = 0u;
                    auto add_col = [&](const int k) {
                        const int j = clist[k];
                        const int c0 = __ldg(w.cptr + j), c1 = __ldg(w.cptr + j + 1);
                        for (int q = c0; q < c1; ++q) vec_xor<NW>(vw, static_cast<int>(__ldg(w.crow + q)), lane);
                    };
                    if (single >= 0) add_col(single);
                    for (uint32_t x = pat; x; x &= x - 1) add_col(ml[__ffs(x) - 1]);
                    reduce_ops(vw);
                };
                auto weigh = [&](const uint32_t (&yw)[NW], const int single, const uint32_t pat) -> double {
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = yw[s];
                    __syncwarp();
                    double ws = 0.0;
                    for (int base = 0; base < nb; base += 32) {
                        const int k = base + lane;
                        bool on = false;
                        double wk = 0.0;
                        if (k < nb) {
                            const uint32_t pr = cprow[k];
                            if (pr != kNone) on = (vsm[pr >> 5] >> (pr & 31)) & 1u;
                            else { const uint32_t rk = crank[k]; on = k == single || (rk < 32u && ((pat >> rk) & 1u)); }
                            if (on) wk = __ldg(wt + clist[k]);
                        }
                        uint32_t mask = __ballot_sync(kFull, on);
                        while (mask) {
                            const int src = __ffs(mask) - 1;
                            mask &= mask - 1;
                            ws = __dadd_rn(ws, __shfl_sync(kFull, wk, src));
                        }
                    }
                    return ws;
                };
                double best = weigh(zw, -1, 0u);
                int best_single = -1;
                uint32_t best_pat = 0u;
                bool found = false;
                auto try_cand = [&](const int single, const uint32_t pat) {
                    uint32_t vw[NW], yw[NW];
                    cand_vec(single, pat, vw);
#pragma unroll
                    for (int s = 0; s < NW; ++s) yw[s] = zw[s] ^ vw[s];
                    const double cw = weigh(yw, single, pat);
                    if (cw < best) { best = cw; best_single = single; best_pat = pat; found = true; }
                };
                if (b.lsd_method == 2) {                         // lsd_cs
                    for (int k = 0; k < nb; ++k)
                        if (cprow[k] == kNone) try_cand(k, 0u);
                    for (int i = 0; i < wsub; ++i)
                        for (int j = i + 1; j < wsub; ++j) try_cand(-1, (1u << i) | (1u << j));
                } else {                                         // lsd_e
                    for (uint32_t pat = 1; pat < (1u << wsub); ++pat) try_cand(-1, pat);
                }
                if (found) {
                    uint32_t vw[NW];
                    cand_vec(best_single, best_pat, vw);
#pragma unroll
                    for (int s = 0; s < NW; ++s) zw[s] ^= vw[s];
                    if (lane == 0) {
                        if (best_single >= 0) commit_col(clist[best_single]);
                        for (uint32_t x = best_pat; x; x &= x - 1) commit_col(clist[ml[__ffs(x) - 1]]);
                    }
                }
                __syncwarp();
            }
        }
        // ---- solution: pivot column i is set iff the reduced syndrome has its pivot row; commit (acc ^= L e, carry = U e)
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = zw[s];
        __syncwarp();
        int alive = 0;
        for (int i = lane; i < nops; i += 32) {
            const uint32_t p = oppiv[i];
            if (p == kDead) continue;
            ++alive;
            if (!((vsm[p >> 5] >> (p & 31)) & 1u)) continue;
            const int j = opcol[i];
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
        // the bit owners go back to "none" for the next shot of this slab
        for (int id = 0; id < nc; ++id) {
            if (!(flag[id] & kActive)) continue;
            for (uint32_t j = bhead[id]; j != kNone; j = bnext[j])
                if (lane == 0) bown[j] = static_cast<uint16_t>(kNone);
        }
        __syncwarp();
        for (int i = lane; i < w.KW; i += 32) {
            const uint64_t v = (static_cast<uint64_t>(accs[2 * i + 1]) << 32) | accs[2 * i];
            b.acc[static_cast<size_t>(shot) * w.KW + i] ^= v;
        }
        for (int i = lane; i < carryW; i += 32) b.carry[static_cast<size_t>(shot) * b.carry_stride32 + i] = car[i];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) alive += __shfl_xor_sync(kFull, alive, d);
        if (lane == 0) {
            atomicAdd(&b.stats[2], 1ull);
            atomicAdd(&b.stats[3], grown);
            atomicAdd(&b.stats[4], static_cast<unsigned long long>(alive));
            atomicMax(&b.stats[5], grown);
        }
    }
}

// =====================================================================================================================
// OSD-0 for windows taller than the shared-memory elimination of osd.cu takes (768 < checks <= 3072; BASELINE config 5 has
// 2250): the same lane-word vectors and row-operation array as the LSD kernel above, ONE warp per failed shot.  The columns come
// in the order osd_sort_kernel wrote (ascending posterior, ties by index); each is reduced by the operations so far, and if it is
// independent its pivot row is the first row IN THE ORACLE'S CURRENT ROW ORDER that carries it -- the oracle swaps the pivot row
// up to position `rank`, which matters only for an inconsistent syndrome on a rank-deficient window, and is followed here
// through the position list `seq`.  Exact early exit as in osd.cu: once the reduced syndrome vanishes on the free rows no later
// pivot changes the answer.  An inconsistent syndrome runs until the rank of the window is reached.
// =====================================================================================================================
struct OsdBigLayout {
    size_t v, seq, oppiv, opcol, accs, car, total;
    int opcap;
};

__host__ __device__ inline OsdBigLayout osd_big_layout(const WinDev& w) {
    OsdBigLayout L;
    const size_t m = static_cast<size_t>(w.rows);
    L.opcap = (w.rows + 31) / 32 * 32;
    size_t o = 0;
    L.v = o; o += 128 * 3;
    L.seq = o; o += al16(m * 2 + 64);
    L.oppiv = o; o += al16(static_cast<size_t>(L.opcap) * 2);
    L.opcol = o; o += al16(static_cast<size_t>(L.opcap) * 2);
    L.accs = o; o += al16(static_cast<size_t>(2 * w.KW) * 4);
    L.car = o; o += al16(static_cast<size_t>((w.carry_rows + 31) / 32 + 1) * 4);
    L.total = o;
    return L;
}

template <int NW>
__global__ void __launch_bounds__(32) osd_big_kernel(const WinDev w, const BatchDev b) {
    extern __shared__ __align__(16) unsigned char sm[];
    const OsdBigLayout L = osd_big_layout(w);
    uint32_t* vsm = reinterpret_cast<uint32_t*>(sm + L.v);
    uint16_t* seq = reinterpret_cast<uint16_t*>(sm + L.seq);
    uint16_t* oppiv = reinterpret_cast<uint16_t*>(sm + L.oppiv);
    uint16_t* opcol = reinterpret_cast<uint16_t*>(sm + L.opcol);
    uint32_t* accs = reinterpret_cast<uint32_t*>(sm + L.accs);
    uint32_t* car = reinterpret_cast<uint32_t*>(sm + L.car);
    const int lane = threadIdx.x;
    const int m = w.rows, n = w.ncols;
    const int carryW = (w.carry_rows + 31) / 32;
    uint32_t* opvec = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(b.lsd_scratch) + static_cast<size_t>(blockIdx.x) * b.lsd_slab);
    const int count = *b.fail_count;
    for (;;) {
        int job = 0;
        if (lane == 0) job = atomicAdd(b.osd_next, 1);
        job = __shfl_sync(kFull, job, 0);
        if (job >= count) break;
        const int shot = b.fail_list[job];
        const uint32_t* syn = b.syn_buf + static_cast<size_t>(shot) * b.syn_stride32;
        const uint16_t* order = b.order_alt ? b.order_alt + static_cast<size_t>(shot) * b.llr_stride
                                            : reinterpret_cast<const uint16_t*>(static_cast<const unsigned char*>(b.llr_buf) +
                                                                                static_cast<size_t>(shot) * b.llr_stride * b.llr_esize);
        __syncwarp();
        uint32_t zw[NW], Pw[NW];
#pragma unroll
        for (int s = 0; s < NW; ++s) {
            const int wi = 32 * s + lane, left = m - 32 * wi;
            zw[s] = (wi < w.rowsW32 ? syn[wi] : 0u) & (left >= 32 ? kFull : (left > 0 ? (1u << left) - 1u : 0u));
            Pw[s] = 0u;
        }
        for (int i = lane; i < m; i += 32) seq[i] = static_cast<uint16_t>(i);
        for (int i = lane; i < 2 * w.KW; i += 32) accs[i] = 0;
        for (int i = lane; i <= carryW; i += 32) car[i] = 0;
        __syncwarp();
        int rank = 0, examined = 0;
        auto satisfied = [&]() {
            bool open = false;
#pragma unroll
            for (int s = 0; s < NW; ++s) open |= (zw[s] & ~Pw[s]) != 0u;
            return !__any_sync(kFull, open);
        };
        bool done = satisfied();
        for (int k = 0; k < n && !done && rank < w.rank; ++k) {
            const int j = order[k];
            ++examined;
            uint32_t vw[NW];
#pragma unroll
            for (int s = 0; s < NW; ++s) vw[s] = 0u;
            {
                const int c0 = __ldg(w.cptr + j), c1 = __ldg(w.cptr + j + 1);
                for (int q = c0; q < c1; ++q) vec_set<NW>(vw, static_cast<int>(__ldg(w.crow + q)), lane, true);
            }
            for (int base = 0; base < rank; base += 32) {
                const uint32_t p = base + lane < rank ? oppiv[base + lane] : kDead;
                uint32_t todo = kFull;
                for (;;) {
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = vw[s];
                    __syncwarp();
                    const bool hit = p != kDead && ((vsm[p >> 5] >> (p & 31)) & 1u);
                    const uint32_t mask = __ballot_sync(kFull, hit) & todo;
                    if (!mask) break;
                    const int i = __ffs(mask) - 1;
#pragma unroll
                    for (int s = 0; s < NW; ++s) vw[s] ^= opvec[(static_cast<size_t>(base + i) * NW + s) * 32 + lane];
                    todo = i == 31 ? 0u : (kFull << (i + 1));
                    if (!todo) break;
                }
            }
            // pivot row: the first position >= rank of the oracle's row order whose row carries the reduced column
            __syncwarp();
#pragma unroll
            for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = vw[s];
            __syncwarp();
            int pos = -1;
            for (int base = rank; base < m && pos < 0; base += 32) {
                const int i = base + lane;
                const uint32_t r = i < m ? seq[i] : 0u;
                const bool has = i < m && ((vsm[r >> 5] >> (r & 31)) & 1u);
                const uint32_t hm = __ballot_sync(kFull, has);
                if (hm) pos = base + __ffs(hm) - 1;
            }
            if (pos < 0) continue;                               // dependent on the columns before it
            const int p = seq[pos];
            __syncwarp();
            if (lane == 0) {
                seq[pos] = seq[rank];
                seq[rank] = static_cast<uint16_t>(p);
                oppiv[rank] = static_cast<uint16_t>(p);
                opcol[rank] = static_cast<uint16_t>(j);
            }
            vec_set<NW>(vw, p, lane, false);                     // the row operation leaves the pivot row alone
            vec_set<NW>(Pw, p, lane, true);
#pragma unroll
            for (int s = 0; s < NW; ++s) opvec[(static_cast<size_t>(rank) * NW + s) * 32 + lane] = vw[s];
            if (vec_bit<NW>(zw, p)) {
#pragma unroll
                for (int s = 0; s < NW; ++s) zw[s] ^= vw[s];
            }
            ++rank;
            __syncwarp();
            done = satisfied();
        }
        // ---- solution on the pivots; commit (acc ^= L e, carry = U e)
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NW; ++s) vsm[32 * s + lane] = zw[s];
        __syncwarp();
        for (int i = lane; i < rank; i += 32) {
            const uint32_t p = oppiv[i];
            if (!((vsm[p >> 5] >> (p & 31)) & 1u)) continue;
            const int j = opcol[i];
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
        __syncwarp();
        for (int i = lane; i < w.KW; i += 32) {
            const uint64_t v = (static_cast<uint64_t>(accs[2 * i + 1]) << 32) | accs[2 * i];
            b.acc[static_cast<size_t>(shot) * w.KW + i] ^= v;
        }
        for (int i = lane; i < carryW; i += 32) b.carry[static_cast<size_t>(shot) * b.carry_stride32 + i] = car[i];
        if (lane == 0) {
            atomicAdd(&b.stats[2], 1ull);
            atomicAdd(&b.stats[3], static_cast<unsigned long long>(examined));
            atomicAdd(&b.stats[4], static_cast<unsigned long long>(rank));
            atomicMax(&b.stats[5], static_cast<unsigned long long>(examined));
        }
    }
}

inline int lsd_nw(const int rows) { return (rows + 1023) / 1024; }

template <typename F>
inline cudaError_t lsd_dispatch(const WinDev& w, int precision, bool hi, F&& f) {
    const int nw = lsd_nw(w.rows);
    if (hi) {
        if (precision == 32) return nw <= 1 ? f(lsd_kernel<float, 1, true>) : (nw == 2 ? f(lsd_kernel<float, 2, true>) : f(lsd_kernel<float, 3, true>));
        return nw <= 1 ? f(lsd_kernel<double, 1, true>) : (nw == 2 ? f(lsd_kernel<double, 2, true>) : f(lsd_kernel<double, 3, true>));
    }
    if (precision == 32) return nw <= 1 ? f(lsd_kernel<float, 1, false>) : (nw == 2 ? f(lsd_kernel<float, 2, false>) : f(lsd_kernel<float, 3, false>));
    return nw <= 1 ? f(lsd_kernel<double, 1, false>) : (nw == 2 ? f(lsd_kernel<double, 2, false>) : f(lsd_kernel<double, 3, false>));
}

}  // namespace

size_t lsd_smem_bytes(const WinDev& w) { return lsd_layout(w).total; }
size_t lsd_slab_bytes(int cols_cap, int max_rows) {
    return al16(static_cast<size_t>(cols_cap) * kLsdColBytes) + static_cast<size_t>(lsd_opcap(max_rows)) * 128 * lsd_nw(max_rows);
}
bool lsd_supported(const WinDev& w) { return w.rows <= 3072 && w.ncols < 65535 && lsd_smem_bytes(w) <= 200 * 1024; }

cudaError_t lsd_configure(const WinDev& w, int precision, bool hi) {
    static size_t have_d[kMaxDevices][2][2][4] = {};
    size_t& have = have_d[device_slot()][hi ? 1 : 0][precision == 32 ? 0 : 1][lsd_nw(w.rows) & 3];
    const size_t s = lsd_smem_bytes(w);
    if (s > have) {
        cudaError_t e = lsd_dispatch(w, precision, hi, [&](auto kern) {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(s));
        });
        if (e != cudaSuccess) return e;
        have = s;
    }
    return cudaSuccess;
}

cudaError_t launch_lsd(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const size_t smem = lsd_smem_bytes(w);
    return lsd_dispatch(w, precision, b.lsd_order > 0, [&](auto kern) {
        kern<<<grid, 32, smem, st>>>(w, b);
        return cudaGetLastError();
    });
}

size_t osd_big_smem_bytes(const WinDev& w) { return osd_big_layout(w).total; }
size_t osd_big_slab_bytes(int max_rows) { return static_cast<size_t>((max_rows + 31) / 32 * 32) * 128 * lsd_nw(max_rows); }
bool osd_big_supported(const WinDev& w) { return w.rows <= 3072 && w.ncols < 65535 && osd_big_smem_bytes(w) <= 200 * 1024; }

template <typename F>
static cudaError_t osd_big_dispatch(const WinDev& w, F&& f) {
    const int nw = lsd_nw(w.rows);
    return nw <= 1 ? f(osd_big_kernel<1>) : (nw == 2 ? f(osd_big_kernel<2>) : f(osd_big_kernel<3>));
}

cudaError_t osd_big_configure(const WinDev& w) {
    static size_t have_d[kMaxDevices][4] = {};
    size_t& have = have_d[device_slot()][lsd_nw(w.rows) & 3];
    const size_t s = osd_big_smem_bytes(w);
    if (s > have) {
        cudaError_t e = osd_big_dispatch(w, [&](auto kern) {
            return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(s));
        });
        if (e != cudaSuccess) return e;
        have = s;
    }
    return cudaSuccess;
}

cudaError_t launch_osd_big(const WinDev& w, const BatchDev& b, int grid, cudaStream_t st) {
    if (b.n_shots == 0) return cudaSuccess;
    const size_t smem = osd_big_smem_bytes(w);
    return osd_big_dispatch(w, [&](auto kern) {
        kern<<<grid, 32, smem, st>>>(w, b);
        return cudaGetLastError();
    });
}

}  // namespace qb
