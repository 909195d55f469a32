// quits_b200/csrc/layout.cpp -- shared-memory layout of one window's BP messages, chosen once per decoder so that the
// bit sweep of bp_kernel_compact (bp.cu) is free of bank conflicts.
//
// The bit sweep gathers, per edge, the message V[row*RS + slot] and the row summary rsum[row], 32 columns per warp.
// Two things are free to choose without touching the arithmetic (the check sweep takes min / parity over a row, which do
// not depend on the order of the row's slots; the bit sweep sums a column's edges in ascending row order whatever the
// column's position in the record array):
//   * which columns share a warp sub-group (records are only required to be sorted by weight), and
//   * which slot of its row an edge occupies.
// The hardware serves a shared-memory request in one pass per sub-group when the sub-group's addresses fall into distinct
// bank classes (or coincide):
//                            message V                       row summary rsum
//   fp64   8-byte accesses:  half-warp of 16, (addr) mod 16  16-byte accesses: quarter-warp of 8, row mod 8
//   fp32   4-byte accesses:  warp of 32,      (addr) mod 32   8-byte accesses: half-warp of 16,  row mod 16
// (bp_kernel_ms2 keeps min1 / min2 in two arrays of single values: its row-summary gather has the message geometry -- `rb`
// overrides the row-summary class count: 16 in fp64, 32 in fp32)
// Step 1 groups columns of equal weight so that, at every edge position, the rows of a sub-group are distinct modulo the
// class count (greedy seed-and-extend, then plateau-walking swaps).  Step 2 assigns slots so that the message addresses
// of a sub-group are distinct modulo the class count (first-fit, then plateau-walking slot swaps inside a row).
// Both are local searches with a fixed work budget; any residual conflict costs a replay, never correctness.
#include <algorithm>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <numeric>

#include "qb_host.h"

namespace qb {

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull) {}
    uint32_t next() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return static_cast<uint32_t>(s >> 32);
    }
    uint32_t below(uint32_t n) { return static_cast<uint32_t>((static_cast<uint64_t>(next()) * n) >> 32); }
};

struct Cols {
    const Window& w;
    int n;
    explicit Cols(const Window& win) : w(win), n(win.ncols) {}
    int wt(int j) const { return static_cast<int>(w.cptr[j + 1] - w.cptr[j]); }
    int row(int j, int q) const { return w.crow[w.cptr[j] + q]; }
};

// excess of one (sub-group, edge position): sum over classes of (distinct rows - 1)
int group_cost(const Cols& c, const int* members, int nm, int RB, int dummy_row) {
    int cost = 0;
    int rows[32];
    for (int q = 0; q < 6; ++q) {
        int k = 0;
        bool any = false, pad = false;
        for (int i = 0; i < nm; ++i) {
            if (c.wt(members[i]) > q) { rows[k++] = c.row(members[i], q); any = true; }
            else pad = true;
        }
        if (!any) break;
        if (pad) rows[k++] = dummy_row;
        std::sort(rows, rows + k);
        k = static_cast<int>(std::unique(rows, rows + k) - rows);
        uint8_t cnt[32] = {0};
        for (int i = 0; i < k; ++i) {
            uint8_t& x = cnt[rows[i] % RB];
            if (x) ++cost;
            ++x;
        }
    }
    return cost;
}

}  // namespace

static void search_bp_layout(const Window& hw, int rs, int precision, int rb, BpLayout& out);

// The search is deterministic, so its result is cached per process by window structure: decoders are rebuilt for every
// call of the drop-in functions, and consecutive sliding windows usually have identical structure.
void optimize_bp_layout(const Window& hw, int rs, int precision, BpLayout& out, int rb) {
    struct Entry { std::vector<int64_t> cptr; std::vector<int32_t> crow; int rs, precision, rb; BpLayout lay; };
    static std::mutex mu;
    static std::vector<std::unique_ptr<Entry>> cache;
    {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& e : cache)
            if (e->rs == rs && e->precision == precision && e->rb == rb && e->crow == hw.crow && e->cptr == hw.cptr) { out = e->lay; return; }
    }
    search_bp_layout(hw, rs, precision, rb, out);
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() >= 64) cache.erase(cache.begin());
    std::unique_ptr<Entry> e(new Entry{hw.cptr, hw.crow, rs, precision, rb, out});
    cache.push_back(std::move(e));
}

static void search_bp_layout(const Window& hw, int rs, int precision, int rb, BpLayout& out) {
    const Cols c(hw);
    const int n = c.n, rows = hw.rows;
    const int RB = rb > 0 ? rb : (precision == 32 ? 16 : 8);        // row-summary classes = sub-group size
    const int VB = precision == 32 ? 32 : 16;       // message classes = sub-group size
    const char* env = std::getenv("QB_LAYOUT_OPT");
    const bool enabled = !(env && env[0] == '0');
    Rng rng(0xB200 + static_cast<uint64_t>(n) * 31 + rows);

    // ------------------------------------------------------------------ step 1: record order
    std::vector<int>& order = out.order;
    order.resize(static_cast<size_t>(n));
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c.wt(a) > c.wt(b); });
    if (enabled && n > RB) {
        // greedy seed-and-extend inside each weight class (class boundaries are kept: the kernel wants records sorted by weight)
        std::vector<int> grouped;
        grouped.reserve(order.size());
        size_t lo = 0;
        while (lo < order.size()) {
            size_t hi = lo;
            const int wc = c.wt(order[lo]);
            while (hi < order.size() && c.wt(order[hi]) == wc) ++hi;
            std::vector<int> pool(order.begin() + lo, order.begin() + hi);
            std::vector<char> used(pool.size(), 0);
            size_t first_free = 0, left = pool.size();
            while (left) {
                while (used[first_free]) ++first_free;
                int classrow[6][32];
                for (int q = 0; q < 6; ++q) std::fill(classrow[q], classrow[q] + 32, -1);
                auto add = [&](int j) {
                    for (int q = 0; q < wc; ++q) {
                        int& slot = classrow[q][c.row(j, q) % RB];
                        if (slot < 0) slot = c.row(j, q);
                    }
                };
                used[first_free] = 1; --left;
                grouped.push_back(pool[first_free]);
                add(pool[first_free]);
                // fill up to the next sub-group boundary of the global record array
                while (left && grouped.size() % RB != 0) {
                    int best = -1, bc = 1 << 30, scanned = 0;
                    for (size_t i = first_free; i < pool.size() && scanned < 768; ++i) {
                        if (used[i]) continue;
                        ++scanned;
                        int cost = 0;
                        for (int q = 0; q < wc; ++q) {
                            const int r = c.row(pool[i], q), have = classrow[q][r % RB];
                            if (have >= 0 && have != r) ++cost;
                        }
                        if (cost < bc) { bc = cost; best = static_cast<int>(i); if (!cost) break; }
                    }
                    used[best] = 1; --left;
                    grouped.push_back(pool[best]);
                    add(pool[best]);
                }
            }
            lo = hi;
        }
        order.swap(grouped);
        // plateau-walking swaps between sub-groups (same weight only)
        const int ng = (n + RB - 1) / RB;
        std::vector<int> gcost(static_cast<size_t>(ng));
        auto members = [&](int g, int& nm) { nm = std::min(RB, n - g * RB); return order.data() + static_cast<size_t>(g) * RB; };
        long total = 0;
        for (int g = 0; g < ng; ++g) { int nm; const int* m = members(g, nm); gcost[g] = group_cost(c, m, nm, RB, rows); total += gcost[g]; }
        std::vector<std::pair<int, int>> wrange(7, {0, 0});            // record range of every weight
        for (int r = 0; r < n; ++r) {
            auto& pr = wrange[c.wt(order[r])];
            if (pr.second == 0) pr.first = r;
            pr.second = r + 1;
        }
        long budget = 1500000;
        for (int sweep = 0; sweep < 600 && total > 0 && budget > 0; ++sweep) {
            long before = total;
            for (int g = 0; g < ng && budget > 0; ++g) {
                if (!gcost[g]) continue;
                int nm; members(g, nm);
                const int ra = g * RB + static_cast<int>(rng.below(static_cast<uint32_t>(nm)));
                const auto pr = wrange[c.wt(order[ra])];
                for (int t = 0; t < 12 && budget > 0; ++t, --budget) {
                    const int rb = pr.first + static_cast<int>(rng.below(static_cast<uint32_t>(pr.second - pr.first)));
                    const int g2 = rb / RB;
                    if (g2 == g) continue;
                    std::swap(order[ra], order[rb]);
                    int n1, n2;
                    const int* m1 = members(g, n1);
                    const int c1 = group_cost(c, m1, n1, RB, rows);
                    const int* m2 = members(g2, n2);
                    const int c2 = group_cost(c, m2, n2, RB, rows);
                    if (c1 + c2 <= gcost[g] + gcost[g2]) {
                        total += c1 + c2 - gcost[g] - gcost[g2];
                        gcost[g] = c1; gcost[g2] = c2;
                        if (!c1) break;
                    } else {
                        std::swap(order[ra], order[rb]);
                    }
                }
            }
            (void)before;
        }
        out.rsum_excess = total;
    }

    // ------------------------------------------------------------------ step 2: slot of every edge
    const size_t nnz = hw.crow.size();
    std::vector<int>& slot = out.slot;                 // indexed like hw.crow
    slot.assign(nnz, 0);
    std::vector<int> rowlen(static_cast<size_t>(rows), 0);
    for (size_t e = 0; e < nnz; ++e) rowlen[hw.crow[e]]++;
    if (!enabled) {
        std::vector<int> fill(static_cast<size_t>(rows), 0);
        for (int j = 0; j < n; ++j)
            for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) slot[e] = fill[hw.crow[e]]++;
        return;
    }
    // message sub-groups: (record / VB, q)
    const int nvg = (n + VB - 1) / VB;
    std::vector<int> egroup(nnz);                      // edge -> sub-group id (nvg * 6 ids)
    std::vector<std::vector<uint8_t>> cnt(static_cast<size_t>(nvg) * 6, std::vector<uint8_t>(static_cast<size_t>(VB), 0));
    const int dummy_class = static_cast<int>((static_cast<long long>(rows) * rs) % VB);
    for (int r = 0; r < n; ++r) {
        const int j = order[r], g = r / VB;
        for (int q = 0; q < c.wt(j); ++q) egroup[hw.cptr[j] + q] = g * 6 + q;
    }
    for (int g = 0; g < nvg; ++g) {                    // dummy edges of lighter columns occupy the dummy slot's class
        int wmax = 0, wmin = 6;
        const int nm = std::min(VB, n - g * VB);
        for (int i = 0; i < nm; ++i) { const int wv = c.wt(order[g * VB + i]); wmax = std::max(wmax, wv); wmin = std::min(wmin, wv); }
        const int lead = (g * VB) / 32 * 32;           // the warp executes the weight of its first record
        wmax = std::max(wmax, c.wt(order[lead]));
        if (nm < VB) wmin = 0;
        for (int q = wmin; q < wmax; ++q) cnt[static_cast<size_t>(g) * 6 + q][dummy_class] = 1;
    }
    std::vector<std::vector<int>> rowedges(static_cast<size_t>(rows));
    for (int j = 0; j < n; ++j)
        for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) rowedges[hw.crow[e]].push_back(static_cast<int>(e));
    auto cls = [&](int e) { return static_cast<int>((static_cast<long long>(hw.crow[e]) * rs + slot[e]) % VB); };
    // first fit, in record order: a free slot of the row whose class is still unused in the edge's sub-group
    {
        std::vector<std::vector<char>> taken(static_cast<size_t>(rows));
        for (int i = 0; i < rows; ++i) taken[i].assign(static_cast<size_t>(rowlen[i]), 0);
        for (int r = 0; r < n; ++r) {
            const int j = order[r];
            for (int64_t e = hw.cptr[j]; e < hw.cptr[j + 1]; ++e) {
                const int i = hw.crow[e];
                int pick = -1, fallback = -1;
                for (int s = 0; s < rowlen[i]; ++s) {
                    if (taken[i][s]) continue;
                    if (fallback < 0) fallback = s;
                    if (!cnt[egroup[e]][(static_cast<long long>(i) * rs + s) % VB]) { pick = s; break; }
                }
                if (pick < 0) pick = fallback;
                taken[i][pick] = 1;
                slot[e] = pick;
                cnt[egroup[e]][cls(static_cast<int>(e))]++;
            }
        }
    }
    long excess = 0;
    for (auto& v : cnt) for (uint8_t x : v) if (x > 1) excess += x - 1;
    std::vector<int> bad;
    for (int sweep = 0; sweep < 400 && excess > 0; ++sweep) {
        bad.clear();
        for (size_t e = 0; e < nnz; ++e) if (cnt[egroup[e]][cls(static_cast<int>(e))] > 1) bad.push_back(static_cast<int>(e));
        for (size_t i = bad.size(); i > 1; --i) std::swap(bad[i - 1], bad[rng.below(static_cast<uint32_t>(i))]);
        for (int e : bad) {
            const int g = egroup[e], r = cls(e);
            if (cnt[g][r] <= 1) continue;
            int best_d = 1, nbest = 0, pick = -1;
            for (int e2 : rowedges[hw.crow[e]]) {
                if (e2 == e) continue;
                const int g2 = egroup[e2], r2 = cls(e2);
                if (r2 == r || g2 == g) continue;
                const int d = -1 + (cnt[g][r2] >= 1 ? 1 : 0) - (cnt[g2][r2] > 1 ? 1 : 0) + (cnt[g2][r] >= 1 ? 1 : 0);
                if (d > 0) continue;
                if (d < best_d) { best_d = d; nbest = 1; pick = e2; }
                else if (d == best_d && rng.below(static_cast<uint32_t>(++nbest)) == 0) pick = e2;      // reservoir choice among equals
            }
            if (pick < 0) continue;
            const int g2 = egroup[pick], r2 = cls(pick);
            cnt[g][r]--; cnt[g][r2]++; cnt[g2][r2]--; cnt[g2][r]++;
            std::swap(slot[e], slot[pick]);
            excess += best_d;
        }
    }
    out.v_excess = excess;
}

// Predicted shared-memory passes of one bit sweep relative to the conflict-free minimum, for the message gather
// (v_ratio; the scatter is the same) and the row-summary gather (r_ratio).  Same model ncu confirms on the device
// ("L1 Wavefronts Shared" / "Ideal" per source line).
void layout_wavefronts(const Window& hw, int rs, int precision, const BpLayout& lay, double* v_ratio, double* r_ratio, int rb) {
    const Cols c(hw);
    const int n = c.n, rows = hw.rows;
    const int RB = rb > 0 ? rb : (precision == 32 ? 16 : 8), VB = precision == 32 ? 32 : 16;
    const int npad = (n + 31) / 32 * 32;
    const long long dummy = static_cast<long long>(rows) * rs;
    long vw = 0, vi = 0, rw = 0, ri = 0;
    for (int w0 = 0; w0 < npad; w0 += 32) {
        const int wmax = w0 < n ? c.wt(lay.order[w0]) : 0;
        for (int q = 0; q < wmax; ++q) {
            long long addr[32];
            int row[32];
            for (int l = 0; l < 32; ++l) {
                const int r = w0 + l;
                if (r < n && c.wt(lay.order[r]) > q) {
                    const int j = lay.order[r];
                    row[l] = c.row(j, q);
                    addr[l] = static_cast<long long>(row[l]) * rs + lay.slot[hw.cptr[j] + q];
                } else { row[l] = rows; addr[l] = dummy; }
            }
            auto passes = [](long long* v, int k, int classes) {
                std::sort(v, v + k);
                k = static_cast<int>(std::unique(v, v + k) - v);
                int cnt[32] = {0}, mx = 0;
                for (int i = 0; i < k; ++i) mx = std::max(mx, ++cnt[v[i] % classes]);
                return mx;
            };
            for (int g = 0; g < 32; g += VB) { long long t[32]; std::copy(addr + g, addr + g + VB, t); vw += passes(t, VB, VB); ++vi; }
            for (int g = 0; g < 32; g += RB) { long long t[32]; for (int i = 0; i < RB; ++i) t[i] = row[g + i]; rw += passes(t, RB, RB); ++ri; }
        }
    }
    if (v_ratio) *v_ratio = vi ? static_cast<double>(vw) / vi : 1.0;
    if (r_ratio) *r_ratio = ri ? static_cast<double>(rw) / ri : 1.0;
}

}  // namespace qb
