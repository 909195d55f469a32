// quits_b200/csrc/qb_host.cpp -- host-side set-up: Stim-text front end, device tape, detector error model,
// QUITS check-matrix conversion and sliding-window plan.  See qb_host.h for the reference citations.
#include "qb_host.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace qb {

// =================================================================================================
// 1. Stim-text front end.  Grammar = what reference src/quits/circuit.py:58-279 can emit (and what
//    stim prints back for such circuits: fused target lists, "REPEAT n {" blocks, rec[-k] targets).
// =================================================================================================
namespace {

struct Line {
    int32_t kind;                       // OpKind, or -1 REPEAT-open, -2 REPEAT-close
    double arg;
    long long count;                    // REPEAT
    std::vector<int32_t> targets;       // qubits, or lookbacks k>0 for DET/OBS
    int lineno;
};

[[noreturn]] void fail_line(int lineno, const std::string& msg) {
    throw value_error("stim text line " + std::to_string(lineno) + ": " + msg);
}

const char* skip_ws(const char* p, const char* e) {
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\r')) ++p;
    return p;
}

int32_t kind_of(const std::string& name) {
    static const std::pair<const char*, int32_t> tbl[] = {
        {"R", OP_R}, {"RZ", OP_R}, {"RX", OP_RX}, {"H", OP_H}, {"CX", OP_CX}, {"CNOT", OP_CX}, {"ZCX", OP_CX},
        {"M", OP_M}, {"MZ", OP_M}, {"MX", OP_MX}, {"MR", OP_MR}, {"MRZ", OP_MR},
        {"X_ERROR", OP_XERR}, {"Z_ERROR", OP_ZERR}, {"DEPOLARIZE1", OP_DEP1}, {"DEPOLARIZE2", OP_DEP2},
        {"DETECTOR", OP_DET}, {"OBSERVABLE_INCLUDE", OP_OBS}};
    for (auto& kv : tbl)
        if (name == kv.first) return kv.second;
    return -100;
}

void tokenize(const char* text, size_t len, std::vector<Line>& lines) {
    const char* p = text;
    const char* end = text + len;
    int lineno = 0;
    while (p < end) {
        const char* nl = static_cast<const char*>(memchr(p, '\n', static_cast<size_t>(end - p)));
        const char* e = nl ? nl : end;
        ++lineno;
        const char* hash = static_cast<const char*>(memchr(p, '#', static_cast<size_t>(e - p)));
        const char* le = hash ? hash : e;
        const char* q = skip_ws(p, le);
        while (le > q && (le[-1] == ' ' || le[-1] == '\t' || le[-1] == '\r')) --le;
        p = nl ? nl + 1 : end;
        if (q == le) continue;
        if (*q == '}') {
            if (skip_ws(q + 1, le) != le) fail_line(lineno, "unexpected text after '}'");
            lines.push_back({-2, 0.0, 0, {}, lineno});
            continue;
        }
        const char* n0 = q;
        while (q < le && ((*q >= 'A' && *q <= 'Z') || (*q >= '0' && *q <= '9') || *q == '_')) ++q;
        std::string name(n0, q);
        if (name.empty()) fail_line(lineno, "unsupported instruction '" + std::string(n0, le) + "'");
        if (name == "REPEAT") {
            q = skip_ws(q, le);
            char* after = nullptr;
            long long cnt = strtoll(q, &after, 10);
            if (after == q || cnt < 0) fail_line(lineno, "bad REPEAT header");
            q = skip_ws(after, le);
            if (q >= le || *q != '{' || skip_ws(q + 1, le) != le) fail_line(lineno, "bad REPEAT header");
            lines.push_back({-1, 0.0, cnt, {}, lineno});
            continue;
        }
        std::vector<double> args;
        if (q < le && *q == '(') {
            ++q;
            while (true) {
                q = skip_ws(q, le);
                if (q < le && *q == ')') { ++q; break; }
                char* after = nullptr;
                double v = strtod(q, &after);
                if (after == q) fail_line(lineno, "bad argument list");
                args.push_back(v);
                q = skip_ws(after, le);
                if (q < le && *q == ',') { ++q; continue; }
                if (q < le && *q == ')') { ++q; break; }
                fail_line(lineno, "bad argument list");
            }
        }
        if (name == "TICK" || name == "QUBIT_COORDS" || name == "SHIFT_COORDS") continue;     // annotations without effect
        if (name == "PAULI_CHANNEL_1" || name == "PAULI_CHANNEL_2")
            throw unsupported_error(name + ": circuit-level decoding is only defined for scalar error rates "
                                           "(the reference calls detector_error_model without approximate_disjoint_errors)");
        int32_t kind = kind_of(name);
        if (kind == -100) fail_line(lineno, "unsupported instruction '" + name + "'");
        Line ln{kind, 0.0, 0, {}, lineno};
        if (kind >= OP_XERR && kind <= OP_DEP2) {
            if (args.size() != 1) fail_line(lineno, name + " needs exactly one probability");
            ln.arg = args[0];
        } else if (kind == OP_OBS) {
            if (args.size() != 1 || args[0] < 0 || args[0] != std::floor(args[0])) fail_line(lineno, "bad observable index");
            ln.arg = args[0];
        } else if (kind != OP_DET && !args.empty()) {
            fail_line(lineno, name + " takes no arguments here");
        }
        while (true) {
            q = skip_ws(q, le);
            if (q >= le) break;
            if (kind == OP_DET || kind == OP_OBS) {
                if (le - q < 6 || strncmp(q, "rec[-", 5) != 0) fail_line(lineno, "bad record target");
                q += 5;
                char* after = nullptr;
                long k = strtol(q, &after, 10);
                if (after == q || k <= 0 || after >= le || *after != ']') fail_line(lineno, "bad record target");
                ln.targets.push_back(static_cast<int32_t>(k));
                q = after + 1;
            } else {
                char* after = nullptr;
                long v = strtol(q, &after, 10);
                if (after == q || v < 0 || (after < le && *after != ' ' && *after != '\t')) fail_line(lineno, "bad qubit target");
                if (v >= (1 << 24)) fail_line(lineno, "qubit index too large");
                ln.targets.push_back(static_cast<int32_t>(v));
                q = after;
            }
        }
        if ((kind == OP_CX || kind == OP_DEP2) && (ln.targets.size() & 1)) fail_line(lineno, name + " needs an even number of targets");
        lines.push_back(std::move(ln));
    }
}

struct Flattener {
    const std::vector<Line>& lines;
    FlatCircuit& fc;
    int64_t meas = 0;
    explicit Flattener(const std::vector<Line>& l, FlatCircuit& f) : lines(l), fc(f) {}

    // walks lines[i..] until the matching close (or the end at depth 0); returns the index after the block
    size_t walk(size_t i, int depth, bool emit) {
        while (i < lines.size()) {
            const Line& ln = lines[i];
            if (ln.kind == -2) {
                if (depth == 0) fail_line(ln.lineno, "unmatched '}'");
                return i + 1;
            }
            if (ln.kind == -1) {
                size_t after = i + 1;
                if (ln.count == 0 || !emit) {
                    after = walk(i + 1, depth + 1, false);
                } else {
                    for (long long r = 0; r < ln.count; ++r) after = walk(i + 1, depth + 1, true);
                }
                i = after;
                continue;
            }
            if (emit) emit_line(ln);
            ++i;
        }
        if (depth != 0) throw value_error("stim text: unterminated REPEAT block");
        return i;
    }

    void emit_line(const Line& ln) {
        FlatOp op{ln.kind, ln.arg, {}};
        if (ln.kind == OP_DET || ln.kind == OP_OBS) {
            for (int32_t k : ln.targets) {
                if (k > meas) fail_line(ln.lineno, "rec[-" + std::to_string(k) + "] looks back past the start of the record");
                op.targets.push_back(static_cast<int32_t>(meas - k));
                fc.max_lookback = std::max(fc.max_lookback, k);
            }
            if (ln.kind == OP_DET) {
                op.arg = static_cast<double>(fc.n_det++);
            } else {
                fc.n_obs = std::max(fc.n_obs, static_cast<int>(ln.arg) + 1);
            }
        } else {
            op.targets = ln.targets;
            for (int32_t t : ln.targets) fc.n_qubits = std::max(fc.n_qubits, t + 1);
            if (ln.kind == OP_M || ln.kind == OP_MX || ln.kind == OP_MR) {
                meas += static_cast<int64_t>(ln.targets.size());
                if (meas > (1ll << 30)) throw value_error("stim text: too many measurements");
            }
            if (ln.kind >= OP_XERR && ln.kind <= OP_DEP2) {
                if (!(ln.arg >= 0.0 && ln.arg <= 0.5)) fail_line(ln.lineno, "noise probability must be in [0, 0.5]");
                int64_t ns = ln.kind == OP_DEP2 ? static_cast<int64_t>(ln.targets.size() / 2) : static_cast<int64_t>(ln.targets.size());
                fc.n_sites = ((fc.n_sites + 3) & ~int64_t(3)) + ns;
            }
        }
        fc.ops.push_back(std::move(op));
        if (fc.ops.size() > (size_t(1) << 26)) throw value_error("stim text: circuit too large after REPEAT unrolling");
    }
};

}  // namespace

void parse_flatten(const char* text, size_t len, FlatCircuit& out) {
    std::vector<Line> lines;
    tokenize(text, len, lines);
    out = FlatCircuit();
    Flattener f(lines, out);
    f.walk(0, 0, true);
    out.n_meas = static_cast<int>(f.meas);
    if (out.n_sites >= 0xFFFFFFFFll) throw value_error("stim text: too many noise sites");
}

// =================================================================================================
// 2. Noise thresholds + device tape
//    Sampling scheme (shared definition with the parity oracle): for noise site s and 64-shot word w
//      level 1: word has >= 1 fault iff philox(seed; s>>2, w, 0)[s&3] < T1,  T1 = floor(2^32 (1-(1-p)^64))
//      level 2: number of faults n = 1 + #{k>=1 : u >= C_k},  C_k = floor(2^64 P(N<=k | N>=1)), N~Bin(64,p)
//      level 3: per fault a Pauli code and a distinct bit position.
// =================================================================================================
void noise_tables(double p, uint32_t* t1, uint64_t* c) {
    for (int k = 0; k < 64; ++k) c[k] = 0;
    *t1 = 0;
    if (!(p > 0.0)) return;
    const double log_keep = log1p(-p);
    const double p_any = -expm1(64.0 * log_keep);
    const double scaled = p_any * 4294967296.0;
    *t1 = scaled >= 4294967295.0 ? 0xFFFFFFFFu : static_cast<uint32_t>(scaled);
    double pmf = exp(64.0 * log_keep);          // P(N = 0)
    const double odds = p / (1.0 - p);
    double cdf = 0.0;                           // P(1 <= N <= k)
    for (int k = 1; k < 64; ++k) {
        pmf = pmf * static_cast<double>(64 - k + 1) / static_cast<double>(k) * odds;
        cdf += pmf;
        const double ratio = cdf / p_any;
        c[k] = ratio >= 1.0 ? UINT64_MAX : static_cast<uint64_t>(ratio * 18446744073709551616.0);
        if (k > 1 && c[k] < c[k - 1]) c[k] = c[k - 1];
    }
}

void build_tape(const FlatCircuit& fc, Tape& tape) {
    tape = Tape();
    int ring = 1;
    while (ring < fc.max_lookback) ring <<= 1;
    tape.ring = ring;
    std::vector<int> stamp(static_cast<size_t>(std::max(fc.n_qubits, 1)), -1);
    int cur_stamp = 0;
    int64_t site = 0, meas = 0;
    std::unordered_map<uint64_t, int> tab_of;          // bit pattern of p -> table index
    for (size_t i = 0; i < fc.ops.size(); ++i) {
        const FlatOp& op = fc.ops[i];
        const int32_t k = op.kind;
        const int nt = static_cast<int>(op.targets.size());
        if (k == OP_DET) {
            const uint32_t det_id = static_cast<uint32_t>(op.arg);
            bool merged = false;
            if (!tape.ops.empty()) {
                TapeOp& last = tape.ops.back();
                if (last.kind == OP_DET && last.flat + last.n == static_cast<int32_t>(i) && last.aux + static_cast<uint32_t>(last.n) == det_id) {
                    last.n += 1;
                    merged = true;
                }
            }
            if (!merged) {
                TapeOp t{};
                t.kind = OP_DET; t.n = 1; t.aux = det_id; t.flat = static_cast<int32_t>(i);
                t.t0 = static_cast<uint32_t>(tape.detptr.size());          // detptr[t0 .. t0+n] bound this block's detectors
                tape.detptr.push_back(static_cast<uint32_t>(tape.detidx.size()));
                tape.ops.push_back(t);
            }
            for (int32_t mi : op.targets) tape.detidx.push_back(static_cast<uint32_t>(mi));
            tape.detptr.push_back(static_cast<uint32_t>(tape.detidx.size()));
            continue;
        }
        if (k == OP_OBS) {
            TapeOp t{};
            t.kind = OP_OBS; t.n = nt; t.t0 = static_cast<uint32_t>(tape.targets.size()); t.aux = static_cast<uint32_t>(op.arg);
            t.flat = static_cast<int32_t>(i);
            for (int32_t mi : op.targets) tape.targets.push_back(static_cast<uint32_t>(mi));
            tape.ops.push_back(t);
            continue;
        }
        if (k >= OP_XERR && k <= OP_DEP2) {
            const int ns = k == OP_DEP2 ? nt / 2 : nt;
            site = (site + 3) & ~int64_t(3);
            TapeOp t{};
            t.kind = k; t.n = ns; t.t0 = static_cast<uint32_t>(tape.targets.size()); t.aux = static_cast<uint32_t>(site);
            t.flat = static_cast<int32_t>(i);
            uint64_t bits;
            memcpy(&bits, &op.arg, 8);
            auto it = tab_of.find(bits);
            if (it == tab_of.end()) {
                it = tab_of.emplace(bits, static_cast<int>(tape.tab_p.size())).first;
                tape.tab_p.push_back(op.arg);
                tape.ctab.resize(tape.ctab.size() + 64);
                uint32_t t1;
                noise_tables(op.arg, &t1, &tape.ctab[tape.ctab.size() - 64]);
            }
            t.tab = it->second;
            uint32_t t1;
            uint64_t dummy[64];
            noise_tables(op.arg, &t1, dummy);
            t.thr = t1;
            // foff = 1: no qubit occurs twice among the targets (what the reference's emitters produce), so the lanes of the frame
            // kernel own their qubits' frame words and update them without atomics
            ++cur_stamp;
            bool distinct = true;
            for (int32_t q : op.targets) { if (stamp[q] == cur_stamp) distinct = false; stamp[q] = cur_stamp; }
            t.foff = distinct ? 1 : 0;
            for (int32_t q : op.targets) tape.targets.push_back(static_cast<uint32_t>(q));
            site += ns;
            tape.ops.push_back(t);      // kept even when thr == 0 so that explicit-fault injection can address it
            continue;
        }
        // gates: split wherever a qubit would be touched twice inside one slice
        const int step = k == OP_CX ? 2 : 1;
        int j = 0;
        while (j < nt) {
            ++cur_stamp;
            if (k == OP_CX && (tape.targets.size() & 1)) tape.targets.push_back(0);     // CX pairs are read as aligned uint2
            TapeOp t{};
            t.kind = k; t.t0 = static_cast<uint32_t>(tape.targets.size()); t.flat = static_cast<int32_t>(i); t.foff = j / step;
            if (k == OP_M || k == OP_MX || k == OP_MR) t.aux = static_cast<uint32_t>(meas + j);
            int cnt = 0;
            while (j < nt) {
                bool clash = stamp[op.targets[j]] == cur_stamp;
                if (step == 2) clash = clash || stamp[op.targets[j + 1]] == cur_stamp || op.targets[j] == op.targets[j + 1];
                if (clash) {
                    if (cnt == 0) throw value_error("stim text: CX with identical control and target");
                    break;
                }
                for (int s = 0; s < step; ++s) {
                    stamp[op.targets[j + s]] = cur_stamp;
                    tape.targets.push_back(static_cast<uint32_t>(op.targets[j + s]));
                }
                j += step;
                ++cnt;
            }
            t.n = cnt;
            if ((k == OP_M || k == OP_MX || k == OP_MR) && cnt > tape.ring) {
                // a single measure slice wider than the ring would overwrite itself: grow the ring
                while (tape.ring < cnt) tape.ring <<= 1;
            }
            tape.ops.push_back(t);
        }
        if (k == OP_M || k == OP_MX || k == OP_MR) meas += nt;
    }
    if (tape.detptr.empty()) tape.detptr.push_back(0);
    if (tape.targets.empty()) tape.targets.push_back(0);
    if (tape.detidx.empty()) tape.detidx.push_back(0);
    if (tape.ctab.empty()) tape.ctab.resize(64);
}

// =================================================================================================
// 3. Circuit -> detector error model: backward sensitivity sweep (per qubit, the set of detectors and
//    observables flipped by an X resp. Z error at this point), independent-component decomposition of
//    the depolarising channels, XOR-combination of equal symptoms, Stim's output order.
// =================================================================================================
namespace {

using Sym = std::vector<int32_t>;       // sorted ids: detector d -> d, observable o -> n_det + o

void sym_xor(const Sym& a, const Sym& b, Sym& out) {
    out.clear();
    size_t i = 0, j = 0;
    while (i < a.size() && j < b.size()) {
        if (a[i] < b[j]) out.push_back(a[i++]);
        else if (b[j] < a[i]) out.push_back(b[j++]);
        else { ++i; ++j; }
    }
    while (i < a.size()) out.push_back(a[i++]);
    while (j < b.size()) out.push_back(b[j++]);
}

void sym_xor_in(Sym& a, const Sym& b, Sym& tmp) {
    if (b.empty()) return;
    sym_xor(a, b, tmp);
    a.swap(tmp);
}

struct SymHash {
    size_t operator()(const Sym& s) const {
        uint64_t h = 1469598103934665603ull;
        for (int32_t v : s) { h ^= static_cast<uint32_t>(v); h *= 1099511628211ull; }
        return static_cast<size_t>(h ^ (h >> 29));
    }
};

}  // namespace

void analyze(const FlatCircuit& fc, Dem& dem) {
    const int D = fc.n_det;
    std::vector<Sym> sens(static_cast<size_t>(fc.n_meas));
    for (const FlatOp& op : fc.ops) {
        int32_t id;
        if (op.kind == OP_DET) id = static_cast<int32_t>(op.arg);
        else if (op.kind == OP_OBS) id = D + static_cast<int32_t>(op.arg);
        else continue;
        for (int32_t m : op.targets) sens[m].push_back(id);
    }
    for (Sym& s : sens) {           // toggle semantics: an id listed twice cancels
        std::sort(s.begin(), s.end());
        Sym out;
        for (size_t i = 0; i < s.size();) {
            size_t j = i;
            while (j < s.size() && s[j] == s[i]) ++j;
            if ((j - i) & 1) out.push_back(s[i]);
            i = j;
        }
        s.swap(out);
    }
    std::vector<int64_t> mbase(fc.ops.size(), 0);
    {
        int64_t cnt = 0;
        for (size_t i = 0; i < fc.ops.size(); ++i) {
            int32_t k = fc.ops[i].kind;
            if (k == OP_M || k == OP_MX || k == OP_MR) { mbase[i] = cnt; cnt += static_cast<int64_t>(fc.ops[i].targets.size()); }
        }
    }
    std::vector<Sym> xs(static_cast<size_t>(fc.n_qubits)), zs(static_cast<size_t>(fc.n_qubits));
    std::unordered_map<Sym, int, SymHash> index;
    std::vector<Sym> syms;
    std::vector<double> prob;
    std::vector<int32_t> rop, rtg, rcd;
    Sym tmp, acc, acc2;

    auto add = [&](const Sym& sym, double q, int32_t opi, int32_t tgt, int32_t code) {
        if (sym.empty() || q == 0.0) return;
        auto it = index.find(sym);
        if (it != index.end()) {
            double p0 = prob[it->second];
            prob[it->second] = p0 * (1.0 - q) + q * (1.0 - p0);
        } else {
            index.emplace(sym, static_cast<int>(syms.size()));
            syms.push_back(sym);
            prob.push_back(q);
            rop.push_back(opi); rtg.push_back(tgt); rcd.push_back(code);
        }
    };

    for (int64_t i = static_cast<int64_t>(fc.ops.size()) - 1; i >= 0; --i) {
        const FlatOp& op = fc.ops[i];
        const std::vector<int32_t>& t = op.targets;
        const int nt = static_cast<int>(t.size());
        switch (op.kind) {
        case OP_DET: case OP_OBS: break;
        case OP_CX:
            for (int j = nt - 2; j >= 0; j -= 2) {
                sym_xor_in(xs[t[j]], xs[t[j + 1]], tmp);
                sym_xor_in(zs[t[j + 1]], zs[t[j]], tmp);
            }
            break;
        case OP_H:
            for (int32_t q : t) xs[q].swap(zs[q]);
            break;
        case OP_R:
            for (int32_t q : t) {
                if (!zs[q].empty()) throw value_error("non-deterministic detector/observable: sensitive to Z on qubit " + std::to_string(q) + " right after R");
                xs[q].clear(); zs[q].clear();
            }
            break;
        case OP_RX:
            for (int32_t q : t) {
                if (!xs[q].empty()) throw value_error("non-deterministic detector/observable: sensitive to X on qubit " + std::to_string(q) + " right after RX");
                xs[q].clear(); zs[q].clear();
            }
            break;
        case OP_M:
            for (int j = nt - 1; j >= 0; --j) sym_xor_in(xs[t[j]], sens[mbase[i] + j], tmp);
            break;
        case OP_MX:
            for (int j = nt - 1; j >= 0; --j) sym_xor_in(zs[t[j]], sens[mbase[i] + j], tmp);
            break;
        case OP_MR:
            for (int j = nt - 1; j >= 0; --j) {
                int32_t q = t[j];
                if (!zs[q].empty()) throw value_error("non-deterministic detector/observable: sensitive to Z on qubit " + std::to_string(q) + " right after MR");
                xs[q] = sens[mbase[i] + j];
                zs[q].clear();
            }
            break;
        case OP_XERR:
            for (int j = nt - 1; j >= 0; --j) add(xs[t[j]], op.arg, static_cast<int32_t>(i), j, 1);
            break;
        case OP_ZERR:
            for (int j = nt - 1; j >= 0; --j) add(zs[t[j]], op.arg, static_cast<int32_t>(i), j, 2);
            break;
        case OP_DEP1: {
            const double q1 = 0.5 - 0.5 * std::sqrt(1.0 - 4.0 * op.arg / 3.0);
            for (int j = nt - 1; j >= 0; --j) {
                int32_t a = t[j];
                sym_xor(xs[a], zs[a], acc);
                add(xs[a], q1, static_cast<int32_t>(i), j, 1);
                add(zs[a], q1, static_cast<int32_t>(i), j, 2);
                add(acc, q1, static_cast<int32_t>(i), j, 3);
            }
        } break;
        case OP_DEP2: {
            const double q2 = 0.5 - 0.5 * std::pow(1.0 - 16.0 * op.arg / 15.0, 0.125);
            for (int j = nt / 2 - 1; j >= 0; --j) {
                int32_t a = t[2 * j], b = t[2 * j + 1];
                for (int c = 1; c < 16; ++c) {
                    acc.clear();
                    if (c & 1) sym_xor_in(acc, xs[a], tmp);
                    if (c & 2) sym_xor_in(acc, zs[a], tmp);
                    if (c & 4) sym_xor_in(acc, xs[b], tmp);
                    if (c & 8) sym_xor_in(acc, zs[b], tmp);
                    add(acc, q2, static_cast<int32_t>(i), j, c);
                }
            }
        } break;
        default: throw unsupported_error("unknown op kind in analyzer");
        }
    }
    // Stim's order: ascending lexicographic on (sorted detector ids, then observable ids); ids of observables
    // are n_det + o, i.e. larger than every detector id, so plain lexicographic order on the symptom vector.
    std::vector<int> order(syms.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = static_cast<int>(i);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return syms[a] < syms[b]; });
    dem = Dem();
    dem.n_det = D;
    dem.n_obs = fc.n_obs;
    for (int i : order) {
        std::vector<int32_t> d, o;
        for (int32_t v : syms[i]) (v < D ? d : o).push_back(v < D ? v : v - D);
        dem.dets.push_back(std::move(d));
        dem.obs.push_back(std::move(o));
        dem.probs.push_back(prob[i]);
        dem.rep_op.push_back(rop[i]); dem.rep_tgt.push_back(rtg[i]); dem.rep_code.push_back(rcd[i]);
    }
}

// =================================================================================================
// 4. DEM -> (H, L, priors)   [reference decoder/base.py:74-127]
//    key = detector set only; a repeated key XOR-combines the probability and keeps the observable set
//    of the first sighting (base.py:93-99); columns are numbered by first sighting.
// =================================================================================================
void dem_to_matrix(const Dem& dem, CheckMatrix& cm) {
    cm = CheckMatrix();
    cm.n_det = dem.n_det;
    cm.n_obs = dem.n_obs;
    std::unordered_map<Sym, int, SymHash> index;
    for (size_t e = 0; e < dem.probs.size(); ++e) {
        Sym key = dem.dets[e];
        std::sort(key.begin(), key.end());
        key.erase(std::unique(key.begin(), key.end()), key.end());           // frozenset semantics
        if (key.empty()) cm.n_detless++;
        const double prob = dem.probs[e];
        auto it = index.find(key);
        if (it == index.end()) {
            index.emplace(key, static_cast<int>(cm.priors.size()));
            cm.col_dets.push_back(key);
            Sym o = dem.obs[e];
            std::sort(o.begin(), o.end());
            o.erase(std::unique(o.begin(), o.end()), o.end());
            cm.col_obs.push_back(std::move(o));
            cm.priors.push_back(prob);
        } else {
            double& pr = cm.priors[it->second];
            pr = pr * (1 - prob) + prob * (1 - pr);
        }
    }
}

// =================================================================================================
// 5. Window plan   [reference decoder/sliding_window.py:130-141 and decoder/base.py:149-188]
// =================================================================================================
void plan_windows(const CheckMatrix& cm, int m, int W, int F, int n_cor_override, WindowPlan& plan) {
    if (m <= 0) throw value_error("hz must have at least one row");
    if (F == 0) throw value_error("Input parameter F cannot be zero.");
    if (F < 0 || W <= 0) throw value_error("W must be positive and F must be positive");
    plan = WindowPlan();
    plan.m = m; plan.K = cm.n_obs; plan.D = cm.n_det; plan.W = W; plan.F = F;
    const int D = cm.n_det;
    const int num_rounds = D / m - 2;
    plan.num_rounds = num_rounds;
    int n_cor;
    if (2 + num_rounds - W >= 0) {
        n_cor = (2 + num_rounds - W) / F;
        if ((2 + num_rounds - W) % F != 0) n_cor += 1;
    } else {
        n_cor = 0;
        plan.whole_history = true;
    }
    if (n_cor_override >= 0) n_cor = n_cor_override;
    plan.n_cor = n_cor;
    const int C = static_cast<int>(cm.priors.size());
    int col_min = 0;
    auto clampD = [&](long long r) { return static_cast<int>(std::min<long long>(std::max<long long>(r, 0), D)); };
    auto fill = [&](Window& w) {
        // CSC of H restricted to the window rows, L and U over the committed prefix
        w.cptr.assign(1, 0);
        w.lptr.assign(1, 0);
        w.uptr.assign(1, 0);
        for (int c = 0; c < w.ncols; ++c) {
            const Sym& d = cm.col_dets[w.col0 + c];
            for (int32_t r : d)
                if (r >= w.row0 && r < w.row0 + w.rows) w.crow.push_back(r - w.row0);
            w.cptr.push_back(static_cast<int64_t>(w.crow.size()));
            w.priors.push_back(cm.priors[w.col0 + c]);
            if (c < w.ncommit) {
                for (int32_t o : cm.col_obs[w.col0 + c]) w.lidx.push_back(o);
                w.lptr.push_back(static_cast<int64_t>(w.lidx.size()));
                for (int32_t r : d)
                    if (r >= w.urow0 && r < w.urow0 + w.urows) w.uidx.push_back(r - w.urow0);
                w.uptr.push_back(static_cast<int64_t>(w.uidx.size()));
            }
        }
    };
    for (int k = 0; k < n_cor; ++k) {
        Window w;
        w.row0 = clampD(static_cast<long long>(k) * F * m);
        const int row1 = clampD((static_cast<long long>(k) * F + W) * m);
        w.rows = row1 - w.row0;
        w.col0 = col_min;
        if (col_min >= C)
            throw value_error("There is no noise in one of the decoding window. This means there are redundant detectors that do not check for any error.");
        const int rowF = clampD(static_cast<long long>(w.row0) + static_cast<long long>(F) * m);
        int col_max = -1, cor_max = -1;
        for (int c = col_min; c < C; ++c) {
            bool touch = false, touchF = false;
            for (int32_t r : cm.col_dets[c]) {
                if (r >= w.row0 && r < row1) {
                    touch = true;
                    if (r < rowF) touchF = true;
                }
            }
            if (touch) col_max = c - col_min;
            if (touchF) cor_max = c - col_min;
        }
        if (col_max < 0 || cor_max < 0)
            throw value_error("zero-size array to reduction operation maximum which has no identity (a decoding window is not touched by any fault)");
        w.ncols = col_max + 1;
        w.ncommit = cor_max + 1;
        w.urow0 = clampD(static_cast<long long>(k + 1) * F * m);
        w.urows = clampD((static_cast<long long>(k + 1) * F + 1) * m) - w.urow0;
        fill(w);
        col_min += cor_max + 1;
        plan.windows.push_back(std::move(w));
    }
    Window last;
    last.row0 = clampD(static_cast<long long>(F) * n_cor * m);
    last.rows = D - last.row0;
    last.col0 = col_min;
    last.ncols = std::max(0, C - col_min);
    last.ncommit = last.ncols;
    last.urow0 = 0;
    last.urows = 0;
    fill(last);
    plan.windows.push_back(std::move(last));
}

}  // namespace qb
