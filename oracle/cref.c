/*
 * oracle/cref.c -- CPU restatement of the QUITS Monte-Carlo hot path.  TEST INFRASTRUCTURE ONLY
 * (see oracle/__init__.py: only tests/, __graft_entry__.smoke() and bench.py's CPU arms may use it;
 * parity unpinned at the stim/ldpc boundary).
 *
 *   qo_sample / qo_inject   Pauli-frame propagation of the memory circuit, i.e. what the reference gets from
 *                           stim's compile_detector_sampler().sample()   (reference src/quits/simulation.py:22-27)
 *   qo_bp_*                 ldpc.BpOsdDecoder.decode(): flooding / serial BP (min-sum, product-sum) and OSD-0
 *                           (called at reference src/quits/decoder/sliding_window.py:171,182)
 *   qo_sw_decode            the per-shot window loop of reference src/quits/decoder/sliding_window.py:162-186
 *
 * Frame rules (standard Pauli-frame semantics): R/RX clear x,z; H swaps; CX c t: x_t ^= x_c, z_c ^= z_t;
 * M: rec <- x; MX: rec <- z; MR: rec <- x then clear; X_ERROR: x ^= B(p); Z_ERROR: z ^= B(p);
 * DEPOLARIZE1: with prob p one of X,Y,Z; DEPOLARIZE2: with prob p one of the 15 non-identity pairs.
 *
 * Randomness is counter based so that CPU and GPU agree bit for bit (stim's own RNG stream is an
 * implementation detail and is not reproduced).  For noise site s and 64-shot word w:
 *   level 1  r = philox(key=seed, ctr=(s>>2, w_lo, w_hi, 0));  the word has >= 1 fault iff r[s&3] < T1(p),
 *            T1 = floor(2^32 (1-(1-p)^64))
 *   level 2  r = philox(ctr=(s, w_lo, w_hi, 1));  u = r0<<32|r1;  n = 1 + #{k>=1 : u >= C_k(p)},
 *            C_k = floor(2^64 P(N<=k | N>=1)),  N ~ Binomial(64,p)
 *   level 3  fault j<n: r = philox(ctr=(s, w_lo, w_hi, 2+j)); Pauli code = 1 + mulhi(r0, 3 or 15) (fixed for X/Z_ERROR);
 *            position = first of the fifteen 6-bit fields of r1,r2,r3 not yet used (then linear probing)
 * which is exactly "each of the 64 shots independently with probability p" up to the 2^-32 / 2^-64 threshold rounding.
 * Sites are numbered in flattened circuit order; every noise instruction starts at a multiple of 4.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { OP_R = 0, OP_RX, OP_H, OP_CX, OP_M, OP_MX, OP_MR, OP_XERR, OP_ZERR, OP_DEP1, OP_DEP2, OP_DET, OP_OBS };

typedef struct { uint32_t v[4]; } ph4;

static inline ph4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3)
{
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    ph4 o = {{c0, c1, c2, c3}};
    return o;
}

/* known-answer hook for tests (Random123 KAT vectors) */
void qo_philox(uint32_t k0, uint32_t k1, const uint32_t* ctr, uint32_t* out)
{
    ph4 r = philox4x32_10(k0, k1, ctr[0], ctr[1], ctr[2], ctr[3]);
    memcpy(out, r.v, 16);
}

/* thresholds for probability p:  out32[0] = T1,  out64[1..63] = C_k  (out64[0] unused = 0) */
void qo_noise_tables(double p, uint32_t* t1, uint64_t* c)
{
    memset(c, 0, 64 * sizeof(uint64_t));
    if (!(p > 0.0)) { *t1 = 0; return; }
    double lq = log1p(-p);
    double pany = -expm1(64.0 * lq);
    double s = pany * 4294967296.0;
    *t1 = s >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)s;
    double pmf = exp(64.0 * lq);
    double odds = p / (1.0 - p);
    double acc = 0.0;
    for (int k = 1; k < 64; ++k) {
        pmf = pmf * (double)(64 - k + 1) / (double)k * odds;
        acc += pmf;
        double ratio = acc / pany;
        c[k] = ratio >= 1.0 ? UINT64_MAX : (uint64_t)(ratio * 18446744073709551616.0);
        if (k > 1 && c[k] < c[k - 1]) c[k] = c[k - 1];
    }
}

typedef struct {
    int n_ops;
    const int32_t* kind;
    const double* arg;
    const int64_t* tstart;      /* n_ops + 1 */
    const int32_t* targets;
    int n_qubits, n_meas, n_det, n_obs;
} flatc;

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

/* apply the faults of site s in word w (level 2+3); returns via callback-free inline writes */
static inline void site_faults(uint32_t k0, uint32_t k1, uint32_t s, uint64_t w, const uint64_t* ctab, int npauli,
                               int fixed_code, uint64_t* m /* m[0..3]: x_a, z_a, x_b, z_b masks */)
{
    uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
    ph4 r = philox4x32_10(k0, k1, s, wl, wh, 1u);
    uint64_t u = ((uint64_t)r.v[0] << 32) | r.v[1];
    int n = 1;
    while (n < 64 && u >= ctab[n]) ++n;
    uint64_t used = 0;
    m[0] = m[1] = m[2] = m[3] = 0;
    for (int j = 0; j < n; ++j) {
        ph4 q = philox4x32_10(k0, k1, s, wl, wh, 2u + (uint32_t)j);
        int code = fixed_code ? fixed_code : 1 + (int)mulhi32(q.v[0], (uint32_t)npauli);
        int pos = -1, last = 0;
        for (int i = 0; i < 15; ++i) {
            int cand = (int)((q.v[1 + i / 5] >> (6 * (i % 5))) & 63u);
            last = cand;
            if (!((used >> cand) & 1u)) { pos = cand; break; }
        }
        if (pos < 0) { pos = last; while ((used >> pos) & 1u) pos = (pos + 1) & 63; }
        used |= 1ull << pos;
        for (int b = 0; b < 4; ++b) if ((code >> b) & 1) m[b] |= 1ull << pos;
    }
}

/*
 * Sample words [word0, word0+nwords) of 64 shots each.  det: [n_det][nwords], obs: [n_obs][nwords] (bit-sliced).
 * inj_*: optional explicit fault list replacing the noise (noise instructions are then skipped):
 *        fault f flips Pauli code inj_code[f] on target/pair inj_tgt[f] of flat op inj_op[f] in shot inj_shot[f]
 *        (shot index local to this call).  Faults must be sorted by inj_op.
 */
int qo_run(int n_ops, const int32_t* kind, const double* arg, const int64_t* tstart, const int32_t* targets,
           int n_qubits, int n_meas, int n_det, int n_obs,
           uint64_t seed, uint64_t word0, uint64_t nwords, uint64_t* det, uint64_t* obs,
           int64_t n_inj, const int32_t* inj_op, const int32_t* inj_tgt, const int32_t* inj_code, const int64_t* inj_shot,
           int nthreads)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    /* per-op noise tables and site bases */
    uint32_t* t1 = (uint32_t*)calloc((size_t)n_ops, sizeof(uint32_t));
    uint64_t* ctab = (uint64_t*)calloc((size_t)n_ops * 64, sizeof(uint64_t));
    uint32_t* sbase = (uint32_t*)calloc((size_t)n_ops, sizeof(uint32_t));
    int64_t* mbase = (int64_t*)calloc((size_t)n_ops, sizeof(int64_t));
    uint64_t site = 0; int64_t mc = 0;
    for (int i = 0; i < n_ops; ++i) {
        int k = kind[i]; int64_t nt = tstart[i + 1] - tstart[i];
        if (k >= OP_XERR && k <= OP_DEP2) {
            if (arg[i] < 0.0 || arg[i] > 0.5) { free(t1); free(ctab); free(sbase); free(mbase); return -2; }
            site = (site + 3) & ~3ull;
            sbase[i] = (uint32_t)site;
            site += (uint64_t)(k == OP_DEP2 ? nt / 2 : nt);
            qo_noise_tables(arg[i], &t1[i], &ctab[(size_t)i * 64]);
        }
        if (k == OP_M || k == OP_MX || k == OP_MR) { mbase[i] = mc; mc += nt; }
    }
    if (site >= 0xFFFFFFFFull) { free(t1); free(ctab); free(sbase); free(mbase); return -3; }
    memset(det, 0, (size_t)n_det * nwords * 8);
    memset(obs, 0, (size_t)n_obs * nwords * 8);
    int err = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        uint64_t* x = (uint64_t*)malloc((size_t)n_qubits * 8);
        uint64_t* z = (uint64_t*)malloc((size_t)n_qubits * 8);
        uint64_t* rec = (uint64_t*)malloc((size_t)(n_meas > 0 ? n_meas : 1) * 8);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t wi = 0; wi < (int64_t)nwords; ++wi) {
            uint64_t w = word0 + (uint64_t)wi;
            uint32_t wl = (uint32_t)w, wh = (uint32_t)(w >> 32);
            memset(x, 0, (size_t)n_qubits * 8);
            memset(z, 0, (size_t)n_qubits * 8);
            int64_t fi = 0;   /* cursor into the injection list */
            for (int i = 0; i < n_ops; ++i) {
                const int32_t* t = targets + tstart[i];
                int64_t nt = tstart[i + 1] - tstart[i];
                switch (kind[i]) {
                case OP_R: case OP_RX: for (int64_t j = 0; j < nt; ++j) { x[t[j]] = 0; z[t[j]] = 0; } break;
                case OP_H: for (int64_t j = 0; j < nt; ++j) { uint64_t a = x[t[j]]; x[t[j]] = z[t[j]]; z[t[j]] = a; } break;
                case OP_CX: for (int64_t j = 0; j + 1 < nt; j += 2) { x[t[j + 1]] ^= x[t[j]]; z[t[j]] ^= z[t[j + 1]]; } break;
                case OP_M: for (int64_t j = 0; j < nt; ++j) rec[mbase[i] + j] = x[t[j]]; break;
                case OP_MX: for (int64_t j = 0; j < nt; ++j) rec[mbase[i] + j] = z[t[j]]; break;
                case OP_MR: for (int64_t j = 0; j < nt; ++j) { rec[mbase[i] + j] = x[t[j]]; x[t[j]] = 0; z[t[j]] = 0; } break;
                case OP_DET: { uint64_t a = 0; for (int64_t j = 0; j < nt; ++j) a ^= rec[t[j]];
                               det[(size_t)(int64_t)arg[i] * nwords + wi] ^= a; } break;
                case OP_OBS: { uint64_t a = 0; for (int64_t j = 0; j < nt; ++j) a ^= rec[t[j]];
                               obs[(size_t)(int64_t)arg[i] * nwords + wi] ^= a; } break;
                default: {
                    int k = kind[i];
                    int two = (k == OP_DEP2);
                    if (n_inj >= 0) {
                        while (fi < n_inj && inj_op[fi] < i) ++fi;
                        for (; fi < n_inj && inj_op[fi] == i; ++fi) {
                            int64_t sh = inj_shot[fi];
                            if ((uint64_t)(sh >> 6) != (uint64_t)wi) continue;
                            uint64_t bit = 1ull << (sh & 63);
                            int code = inj_code[fi]; int j = inj_tgt[fi];
                            int a = two ? t[2 * j] : t[j];
                            if (code & 1) x[a] ^= bit;
                            if (code & 2) z[a] ^= bit;
                            if (two) { int b = t[2 * j + 1]; if (code & 4) x[b] ^= bit; if (code & 8) z[b] ^= bit; }
                        }
                        /* rewind is unnecessary: ops are visited in increasing order and the list is sorted */
                        break;
                    }
                    int64_t ns = two ? nt / 2 : nt;
                    uint32_t thr = t1[i];
                    if (thr == 0) break;
                    int npauli = k == OP_DEP1 ? 3 : (k == OP_DEP2 ? 15 : 0);
                    int fixed = k == OP_XERR ? 1 : (k == OP_ZERR ? 2 : 0);
                    for (int64_t g = 0; g < ns; g += 4) {
                        uint32_t s0 = sbase[i] + (uint32_t)g;
                        ph4 r = philox4x32_10(k0, k1, s0 >> 2, wl, wh, 0u);
                        for (int l = 0; l < 4 && g + l < ns; ++l) {
                            if (!(r.v[l] < thr || thr == 0xFFFFFFFFu)) continue;
                            uint64_t m[4];
                            site_faults(k0, k1, s0 + (uint32_t)l, w, &ctab[(size_t)i * 64], npauli, fixed, m);
                            int64_t j = g + l;
                            int a = two ? t[2 * j] : t[j];
                            x[a] ^= m[0]; z[a] ^= m[1];
                            if (two) { int b = t[2 * j + 1]; x[b] ^= m[2]; z[b] ^= m[3]; }
                        }
                    }
                } break;
                }
            }
        }
        free(x); free(z); free(rec);
    }
    free(t1); free(ctab); free(sbase); free(mbase);
    return err;
}

/* ============================================================================================
 * BP (+OSD-0).  Restated from the published ldpc v2 algorithm (bp.hpp: bp_decode_parallel /
 * bp_decode_serial; osd.hpp: OSD-0 = LLR-ordered row reduction + solve).  ldpc is not vendored.
 *   - messages start at llr0_j = log((1-p_j)/p_j)
 *   - flooding check update: forward/backward running min (min-sum) or tanh product (product-sum)
 *     over the row in ascending column order; sign from syndrome + #{v<=0}
 *   - bit update: forward prefix sums over the column in ascending row order give v[e] and the
 *     posterior; hard decision e_j = [LLR_j <= 0]; stop as soon as H e == s; otherwise the
 *     backward suffix pass completes v[e]
 *   - higher-order OSD (osd.hpp, restated from the published description; ldpc is not vendored): complete the row
 *     reduction, then try flipping sets of NON-pivot columns (in the LLR order): osd_cs(w) = every single non-pivot
 *     column, then every pair among the first w; osd_e(w) = every non-empty subset of the first w (pattern p = 1..2^w-1,
 *     bit b <-> b-th non-pivot column).  A candidate's pivot bits are the reduced syndrome XOR the reduced flipped
 *     columns; its weight is sum_{j: x_j = 1} log(1/p_j) over the column index (priors, not posteriors); the first
 *     candidate strictly lighter than everything before it wins, starting from the OSD-0 solution.
 *   - OSD-0: columns by ascending posterior LLR (stable: ties by column index -- ldpc's own tie order
 *     comes from std::sort and is unspecified), row-reduce choosing as pivot row the first row at or
 *     below the current rank that has a 1 (rows swapped into place), solve on the pivots, rest 0.
 * Two precisions: f64 (what ldpc computes in) and f32 (what the CUDA kernels compute in; the GPU is
 * held bit-exact to this one, and within 1e-4 of the f64 posteriors).
 * ============================================================================================ */
typedef struct {
    int m, n, nnz;
    int* rowptr; int* colidx;         /* CSR, ascending column inside a row; edge id = CSR position */
    int* colptr; int* coledge; int* colrow;   /* CSC view: edge ids / rows of column j in ascending row order */
    double* prior;
    double* llr0;                     /* log((1-p_j)/p_j) */
    int max_iter; int method;         /* 0 = min-sum, 1 = product-sum */
    int schedule;                     /* 0 = parallel (flooding), 1 = serial */
    double alpha;                     /* ms_scaling_factor; 0 => 1 - 2^-it */
    int precision;                    /* 64 or 32 */
    int osd;                          /* 1: OSD when BP fails; 0: return BP output */
    int osd_method;                   /* 0 osd_0 | 1 osd_e (exhaustive) | 2 osd_cs (combination sweep) | 3 lsd_0 | 4 lsd_e | 5 lsd_cs */
    int osd_order;
} qo_bp;

qo_bp* qo_bp_create(int m, int n, const int64_t* indptr /*csc n+1*/, const int32_t* indices /*rows, ascending*/,
                    const double* priors, int max_iter, int method, int schedule, double alpha, int precision, int osd)
{
    qo_bp* d = (qo_bp*)calloc(1, sizeof(qo_bp));
    int nnz = (int)indptr[n];
    d->m = m; d->n = n; d->nnz = nnz;
    d->rowptr = (int*)calloc((size_t)m + 1, sizeof(int));
    d->colidx = (int*)malloc((size_t)nnz * sizeof(int) + 4);
    d->colptr = (int*)malloc(((size_t)n + 1) * sizeof(int));
    d->coledge = (int*)malloc((size_t)nnz * sizeof(int) + 4);
    d->colrow = (int*)malloc((size_t)nnz * sizeof(int) + 4);
    d->prior = (double*)malloc((size_t)n * sizeof(double) + 8);
    d->llr0 = (double*)malloc((size_t)n * sizeof(double) + 8);
    for (int j = 0; j < n; ++j) { d->prior[j] = priors[j]; d->llr0[j] = log((1.0 - priors[j]) / priors[j]); d->colptr[j] = (int)indptr[j]; }
    d->colptr[n] = nnz;
    for (int e = 0; e < nnz; ++e) d->rowptr[indices[e] + 1]++;
    for (int i = 0; i < m; ++i) d->rowptr[i + 1] += d->rowptr[i];
    int* fill = (int*)malloc((size_t)m * sizeof(int) + 4);
    for (int i = 0; i < m; ++i) fill[i] = d->rowptr[i];
    for (int j = 0; j < n; ++j)                       /* columns ascending => ascending column inside each row */
        for (int64_t q = indptr[j]; q < indptr[j + 1]; ++q) {
            int r = indices[q]; int e = fill[r]++;
            d->colidx[e] = j; d->coledge[q] = e; d->colrow[q] = r;
        }
    free(fill);
    d->max_iter = max_iter; d->method = method; d->schedule = schedule; d->alpha = alpha;
    d->precision = precision; d->osd = osd;
    d->osd_method = 0; d->osd_order = 0;
    return d;
}

/* osd_method: 0 osd_0 | 1 osd_e | 2 osd_cs (order 0 is OSD-0 whatever the method) | 3 lsd_0 (LSD post-processing, order 0) |
 * 4 lsd_e | 5 lsd_cs (LSD with the per-cluster candidate sweep below; order 0 is LSD-0) */
void qo_bp_set_osd(qo_bp* d, int osd_method, int osd_order)
{
    d->osd_method = osd_method;
    d->osd_order = osd_order < 0 ? 0 : ((osd_method == 1 || osd_method == 4) && osd_order > 20 ? 20 : (osd_order > 64 ? 64 : osd_order));
}

void qo_bp_free(qo_bp* d)
{
    if (!d) return;
    free(d->rowptr); free(d->colidx); free(d->colptr); free(d->coledge); free(d->colrow); free(d->prior); free(d->llr0); free(d);
}

#define REAL double
#define RMAX DBL_MAX
#define RTANH tanh
#define RLOG log
#define RFABS fabs
#define BPFN(x) x##_f64
#include "bp_impl.inc"
#undef REAL
#undef RMAX
#undef RTANH
#undef RLOG
#undef RFABS
#undef BPFN
#define REAL float
#define RMAX FLT_MAX
#define RTANH tanhf
#define RLOG logf
#define RFABS fabsf
#define BPFN(x) x##_f32
#include "bp_impl.inc"

/* OSD on posterior LLRs (as doubles; f32 posteriors are widened exactly, so the order is unchanged). */
static double sol_weight(const qo_bp* d, const uint8_t* x)
{
    double w = 0.0;
    for (int j = 0; j < d->n; ++j) if (x[j]) w += log(1.0 / d->prior[j]);
    return w;
}

static void osd_decode(const qo_bp* d, const uint8_t* syn, const double* llr, uint8_t* ehat)
{
    int m = d->m, n = d->n;
    const int higher = d->osd_method != 0 && d->osd_order > 0;
    int* order = (int*)malloc((size_t)n * sizeof(int));
    /* stable merge sort on (llr, index) */
    int* tmp = (int*)malloc((size_t)n * sizeof(int));
    for (int j = 0; j < n; ++j) order[j] = j;
    for (int width = 1; width < n; width *= 2) {
        for (int lo = 0; lo < n; lo += 2 * width) {
            int mid = lo + width < n ? lo + width : n, hi = lo + 2 * width < n ? lo + 2 * width : n;
            int a = lo, b = mid, o = lo;
            while (a < mid && b < hi) tmp[o++] = (llr[order[b]] < llr[order[a]]) ? order[b++] : order[a++];
            while (a < mid) tmp[o++] = order[a++];
            while (b < hi) tmp[o++] = order[b++];
        }
        int* sw = order; order = tmp; tmp = sw;
    }
    int nw = (n + 1 + 63) / 64;                       /* +1: augmented syndrome column at bit n */
    uint64_t* A = (uint64_t*)calloc((size_t)m * nw, 8);
    for (int k = 0; k < n; ++k) {                      /* permuted column k = original column order[k] */
        int j = order[k];
        for (int q = d->colptr[j]; q < d->colptr[j + 1]; ++q)
            A[(size_t)d->colrow[q] * nw + (k >> 6)] |= 1ull << (k & 63);
    }
    for (int i = 0; i < m; ++i) if (syn[i] & 1) A[(size_t)i * nw + (n >> 6)] |= 1ull << (n & 63);
    int* pivcol = (int*)malloc((size_t)m * sizeof(int));
    uint8_t* ispiv = (uint8_t*)calloc((size_t)n + 1, 1);
    int rank = 0;
    uint64_t* swp = (uint64_t*)malloc((size_t)nw * 8);
    for (int k = 0; k < n && rank < m; ++k) {
        int w = k >> 6; uint64_t bit = 1ull << (k & 63);
        int p = -1;
        for (int i = rank; i < m; ++i) if (A[(size_t)i * nw + w] & bit) { p = i; break; }
        if (p < 0) continue;
        if (p != rank) {
            memcpy(swp, &A[(size_t)p * nw], (size_t)nw * 8);
            memcpy(&A[(size_t)p * nw], &A[(size_t)rank * nw], (size_t)nw * 8);
            memcpy(&A[(size_t)rank * nw], swp, (size_t)nw * 8);
        }
        const uint64_t* pr = &A[(size_t)rank * nw];
        for (int i = 0; i < m; ++i) {
            if (i == rank || !(A[(size_t)i * nw + w] & bit)) continue;
            uint64_t* ri = &A[(size_t)i * nw];
            for (int q = w; q < nw; ++q) ri[q] ^= pr[q];
        }
        ispiv[k] = 1;
        pivcol[rank++] = k;
    }
    memset(ehat, 0, (size_t)n);
    for (int r = 0; r < rank; ++r)
        if (A[(size_t)r * nw + (n >> 6)] & (1ull << (n & 63))) ehat[order[pivcol[r]]] = 1;
    if (higher) {
        /* non-pivot columns in LLR order */
        int nnp = 0;
        int* np = (int*)malloc((size_t)n * sizeof(int) + 4);
        for (int k = 0; k < n; ++k) if (!ispiv[k]) np[nnp++] = k;
        int w = d->osd_order < nnp ? d->osd_order : nnp;
        uint8_t* x = (uint8_t*)malloc((size_t)n + 8);
        double best = sol_weight(d, ehat);
        /* candidate = set of at most `w` non-pivot positions (CS singles range over all of them) */
        int flip[64];
        #define QO_TRY(NF)                                                                                   \
            do {                                                                                             \
                memset(x, 0, (size_t)n);                                                                     \
                for (int r = 0; r < rank; ++r) {                                                             \
                    uint64_t bitv = (A[(size_t)r * nw + (n >> 6)] >> (n & 63)) & 1ull;                       \
                    for (int f = 0; f < (NF); ++f) bitv ^= (A[(size_t)r * nw + (flip[f] >> 6)] >> (flip[f] & 63)) & 1ull; \
                    if (bitv) x[order[pivcol[r]]] = 1;                                                       \
                }                                                                                            \
                for (int f = 0; f < (NF); ++f) x[order[flip[f]]] = 1;                                        \
                double cw = sol_weight(d, x);                                                                \
                if (cw < best) { best = cw; memcpy(ehat, x, (size_t)n); }                                    \
            } while (0)
        if (d->osd_method == 2) {                     /* combination sweep */
            for (int i = 0; i < nnp; ++i) { flip[0] = np[i]; QO_TRY(1); }
            for (int i = 0; i < w; ++i)
                for (int j = i + 1; j < w; ++j) { flip[0] = np[i]; flip[1] = np[j]; QO_TRY(2); }
        } else {                                      /* exhaustive over the first w (w <= 20 enforced by the caller) */
            for (uint32_t pat = 1; pat < (1u << w); ++pat) {
                int nf = 0;
                for (int b = 0; b < w; ++b) if ((pat >> b) & 1u) flip[nf++] = np[b];
                QO_TRY(nf);
            }
        }
        #undef QO_TRY
        free(np); free(x);
    }
    free(order); free(tmp); free(A); free(pivcol); free(swp); free(ispiv);
}

/* ============================================================================================
 * LSD-0 (ldpc v2 BpLsdDecoder, lsd_order = 0, bits_per_step = 1, on-the-fly elimination).  Restated from the published
 * algorithm (Hillmann, Berent, Quintavalle, Eisert, Wille, Roffe: "Localized statistics decoding", 2024, and ldpc's
 * lsd.hpp as far as remembered; ldpc is not vendored -- parity unpinned, like the rest of this section):
 *   - one cluster per unsatisfied check (ascending check index = cluster id): checks {i}, boundary {i}, no bits
 *   - while some active cluster is invalid: every cluster that was active and invalid at the start of the round, in
 *     ascending (number of bits, id) order, grows by ONE bit if it is still active: among the bits adjacent to its boundary checks
 *     and not in the cluster, the one with the smallest BP posterior LLR (ties: smallest index; ldpc's tie order comes from
 *     std::sort over a hash-set walk and is unspecified).  Boundary checks without such a bit leave the boundary.
 *     The bit's column joins the cluster with its checks; a check that belongs to another cluster makes the two collide:
 *     they merge (in the order the collisions were met), the one with fewer bits into the other (ties: into the growing side),
 *     its bits appended to the survivor's column list in their own order, its elimination discarded.
 *   - the survivor's new columns are row-reduced "on the fly" after the columns it had already reduced; the cluster is valid
 *     when its syndrome lies in the span of its columns
 *   - solution of a cluster: the unique one supported on its pivot columns = the first linearly independent columns of its
 *     column list (so the pivot-row rule is immaterial); bits outside every cluster are 0
 * A cluster that is invalid and has no bit left to add (syndrome outside the image of H) stops growing and contributes the
 * pivot part of its reduced syndrome.
 * ============================================================================================ */
typedef struct {
    int active, valid, stuck;
    int nbits, capbits; int* cols;
    int nelim;
    int nchecks, capchecks; int* checks;
    int nops, capops; int* oprow; int* opcol; uint64_t* opvec;
} lsd_cl;

static void lsd_push(int** a, int* n, int* cap, int v)
{
    if (*n == *cap) { *cap = *cap ? 2 * *cap : 8; *a = (int*)realloc(*a, (size_t)*cap * sizeof(int)); }
    (*a)[(*n)++] = v;
}

/* diagnostics of the last lsd_decode of this thread: row operations created (survivors re-reduce absorbed columns, so this can
 * exceed the number of checks), bits added, merges -- lets a test prove that a case drives the GPU kernel's operation array
 * through its compaction */
static _Thread_local int64_t lsd_diag[3];
void qo_lsd_diag(int64_t* out) { out[0] = lsd_diag[0]; out[1] = lsd_diag[1]; out[2] = lsd_diag[2]; }

/* --------------------------------------------------------------------------------------------
 * LSD beyond order 0 (lsd_method = lsd_e / lsd_cs, lsd_order = w > 0).  PARITY UNPINNED: the published description ("the
 * higher-order reprocessing of OSD applied inside every cluster") fixes the idea, not ldpc's tie-breaking details, and ldpc is
 * not available here.  This restatement applies the candidate sweep of osd_decode above to each final cluster on its own:
 *   - columns of the cluster in column-list order b_0 .. b_{nb-1}; pivots = the first independent ones (as in LSD-0); the other
 *     columns, in list order, are the non-pivots np_0 .. np_{t-1}
 *   - candidates, in this order -- lsd_cs: {np_i} for every i, then {np_i, np_j} for i < j < min(w, t); lsd_e: every non-empty
 *     subset of the first min(w, t) non-pivots, as the bit patterns 1, 2, 3, ...
 *   - a candidate sets its non-pivot columns to 1 and solves the pivots for the cluster's syndrome
 *   - weight of a solution = sum over its set columns, in list order, of log(1 / p_j) (sequential double additions from 0);
 *     the first candidate strictly lighter than everything before it (LSD-0's solution included) wins
 * ldpc additionally lets clusters with fewer than w non-pivots grow by further bits first; that step is not restated.
 * -------------------------------------------------------------------------------------------- */
static void lsd_reduce(const lsd_cl* B, uint64_t* v, int nw)
{
    for (int o = 0; o < B->nops; ++o) {
        int p = B->oprow[o];
        if ((v[p >> 6] >> (p & 63)) & 1ull)
            for (int q = 0; q < nw; ++q) v[q] ^= B->opvec[(size_t)o * nw + q];
    }
}

static double lsd_cl_weight(const qo_bp* d, const lsd_cl* B, const int* prow, const uint64_t* y, const uint8_t* inF)
{
    double w = 0.0;
    for (int k = 0; k < B->nbits; ++k) {
        int on = prow[k] >= 0 ? (int)((y[prow[k] >> 6] >> (prow[k] & 63)) & 1ull) : inF[k];
        if (on) w += log(1.0 / d->prior[B->cols[k]]);
    }
    return w;
}

/* z: the cluster's reduced syndrome.  Writes the cluster's part of ehat. */
static void lsd_cluster_higher(const qo_bp* d, const lsd_cl* B, const uint64_t* z, int nw, int* posof, uint8_t* ehat)
{
    const int nb = B->nbits;
    int* prow = (int*)malloc((size_t)nb * sizeof(int) + 4);
    int* np = (int*)malloc((size_t)nb * sizeof(int) + 4);
    uint8_t* inF = (uint8_t*)calloc((size_t)nb + 1, 1);
    uint64_t* v = (uint64_t*)malloc((size_t)nw * 8);
    uint64_t* ybest = (uint64_t*)malloc((size_t)nw * 8);
    int bestF[64], nbestF = 0, flip[64];
    for (int k = 0; k < nb; ++k) { prow[k] = -1; posof[B->cols[k]] = k; }
    for (int o = 0; o < B->nops; ++o) prow[posof[B->opcol[o]]] = B->oprow[o];
    int nnp = 0;
    for (int k = 0; k < nb; ++k) if (prow[k] < 0) np[nnp++] = k;
    memcpy(ybest, z, (size_t)nw * 8);
    double best = lsd_cl_weight(d, B, prow, z, inF);
    const int w = d->osd_order < nnp ? d->osd_order : nnp;
    #define QO_LSD_TRY(NF)                                                                            \
        do {                                                                                          \
            memset(v, 0, (size_t)nw * 8);                                                             \
            for (int f_ = 0; f_ < (NF); ++f_) {                                                       \
                int j_ = B->cols[flip[f_]];                                                           \
                inF[flip[f_]] = 1;                                                                    \
                for (int q_ = d->colptr[j_]; q_ < d->colptr[j_ + 1]; ++q_) v[d->colrow[q_] >> 6] ^= 1ull << (d->colrow[q_] & 63); \
            }                                                                                         \
            lsd_reduce(B, v, nw);                                                                     \
            for (int q_ = 0; q_ < nw; ++q_) v[q_] ^= z[q_];                                           \
            double cw_ = lsd_cl_weight(d, B, prow, v, inF);                                           \
            if (cw_ < best) { best = cw_; memcpy(ybest, v, (size_t)nw * 8); nbestF = (NF); memcpy(bestF, flip, sizeof(int) * (size_t)(NF)); } \
            for (int f_ = 0; f_ < (NF); ++f_) inF[flip[f_]] = 0;                                      \
        } while (0)
    if (d->osd_method == 5) {
        for (int i = 0; i < nnp; ++i) { flip[0] = np[i]; QO_LSD_TRY(1); }
        for (int i = 0; i < w; ++i)
            for (int j = i + 1; j < w; ++j) { flip[0] = np[i]; flip[1] = np[j]; QO_LSD_TRY(2); }
    } else {
        for (uint32_t pat = 1; pat < (1u << w); ++pat) {
            int nf = 0;
            for (int b = 0; b < w; ++b) if ((pat >> b) & 1u) flip[nf++] = np[b];
            QO_LSD_TRY(nf);
        }
    }
    #undef QO_LSD_TRY
    for (int k = 0; k < nb; ++k)
        if (prow[k] >= 0 && ((ybest[prow[k] >> 6] >> (prow[k] & 63)) & 1ull)) ehat[B->cols[k]] = 1;
    for (int f = 0; f < nbestF; ++f) ehat[B->cols[bestF[f]]] = 1;
    free(prow); free(np); free(inF); free(v); free(ybest);
}

static void lsd_decode(const qo_bp* d, const uint8_t* syn, const double* llr, uint8_t* ehat)
{
    lsd_diag[0] = lsd_diag[1] = lsd_diag[2] = 0;
    const int m = d->m, n = d->n, nw = (m + 63) / 64;
    memset(ehat, 0, (size_t)n);
    int nc = 0;
    for (int i = 0; i < m; ++i) nc += syn[i] & 1;
    if (nc == 0) return;
    lsd_cl* cl = (lsd_cl*)calloc((size_t)nc, sizeof(lsd_cl));
    int* bit_owner = (int*)malloc((size_t)n * sizeof(int));
    int* check_owner = (int*)malloc((size_t)m * sizeof(int));
    uint8_t* isb = (uint8_t*)calloc((size_t)m, 1);
    uint8_t* ispiv = (uint8_t*)calloc((size_t)m, 1);
    uint64_t* v = (uint64_t*)malloc((size_t)nw * 8);
    uint64_t* z = (uint64_t*)malloc((size_t)nw * 8);
    int* inv = (int*)malloc((size_t)nc * sizeof(int));
    int* mlist = (int*)malloc((size_t)nc * sizeof(int));
    for (int j = 0; j < n; ++j) bit_owner[j] = -1;
    for (int i = 0; i < m; ++i) check_owner[i] = -1;
    for (int i = 0, c = 0; i < m; ++i)
        if (syn[i] & 1) {
            cl[c].active = 1;
            lsd_push(&cl[c].checks, &cl[c].nchecks, &cl[c].capchecks, i);
            check_owner[i] = c; isb[i] = 1;
            ++c;
        }
    int ninv = nc;
    for (int c = 0; c < nc; ++c) inv[c] = c;
    #define LSD_REDUCED_SYNDROME(C)                                                                   \
        do {                                                                                          \
            memset(z, 0, (size_t)nw * 8);                                                             \
            for (int q_ = 0; q_ < (C)->nchecks; ++q_) {                                               \
                int r_ = (C)->checks[q_];                                                             \
                if (syn[r_] & 1) z[r_ >> 6] |= 1ull << (r_ & 63);                                     \
            }                                                                                         \
            for (int o_ = 0; o_ < (C)->nops; ++o_) {                                                  \
                int p_ = (C)->oprow[o_];                                                              \
                if ((z[p_ >> 6] >> (p_ & 63)) & 1ull)                                                 \
                    for (int q_ = 0; q_ < nw; ++q_) z[q_] ^= (C)->opvec[(size_t)o_ * nw + q_];        \
            }                                                                                         \
        } while (0)
    while (ninv > 0) {
        for (int t = 0; t < ninv; ++t) {
            lsd_cl* c = &cl[inv[t]];
            const int cid = inv[t];
            if (!c->active) continue;
            /* growth candidates */
            int best = -1;
            for (int q = 0; q < c->nchecks; ++q) {
                int r = c->checks[q];
                if (!isb[r]) continue;
                int any = 0;
                for (int e = d->rowptr[r]; e < d->rowptr[r + 1]; ++e) {
                    int j = d->colidx[e];
                    if (bit_owner[j] == cid) continue;
                    any = 1;
                    if (best < 0 || llr[j] < llr[best] || (llr[j] == llr[best] && j < best)) best = j;
                }
                if (!any) isb[r] = 0;
            }
            if (best < 0) { if (!c->valid) { c->valid = 1; c->stuck = 1; } continue; }
            /* the bit joins the cluster; collisions are noted */
            int nm = 0;
            lsd_diag[1]++;
            bit_owner[best] = cid;
            lsd_push(&c->cols, &c->nbits, &c->capbits, best);
            for (int q = d->colptr[best]; q < d->colptr[best + 1]; ++q) {
                int r = d->colrow[q], o = check_owner[r];
                if (o == cid) continue;
                if (o < 0) { check_owner[r] = cid; isb[r] = 1; lsd_push(&c->checks, &c->nchecks, &c->capchecks, r); continue; }
                int seen = 0;
                for (int k = 0; k < nm; ++k) seen |= mlist[k] == o;
                if (!seen) mlist[nm++] = o;
            }
            int big = cid;
            for (int k = 0; k < nm; ++k) {
                int a = big, b2 = mlist[k];
                int small;
                if (cl[a].nbits < cl[b2].nbits) { small = a; big = b2; } else { small = b2; big = a; }
                lsd_cl* S = &cl[small]; lsd_cl* B = &cl[big];
                for (int q = 0; q < S->nbits; ++q) { bit_owner[S->cols[q]] = big; lsd_push(&B->cols, &B->nbits, &B->capbits, S->cols[q]); }
                for (int q = 0; q < S->nchecks; ++q) { check_owner[S->checks[q]] = big; lsd_push(&B->checks, &B->nchecks, &B->capchecks, S->checks[q]); }
                for (int o = 0; o < S->nops; ++o) ispiv[S->oprow[o]] = 0;
                S->nops = 0; S->active = 0;
                lsd_diag[2]++;
            }
            /* on-the-fly elimination of the survivor's new columns */
            lsd_cl* B = &cl[big];
            for (int k = B->nelim; k < B->nbits; ++k) {
                int j = B->cols[k];
                memset(v, 0, (size_t)nw * 8);
                for (int q = d->colptr[j]; q < d->colptr[j + 1]; ++q) v[d->colrow[q] >> 6] |= 1ull << (d->colrow[q] & 63);
                for (int o = 0; o < B->nops; ++o) {
                    int p = B->oprow[o];
                    if ((v[p >> 6] >> (p & 63)) & 1ull)
                        for (int q = 0; q < nw; ++q) v[q] ^= B->opvec[(size_t)o * nw + q];
                }
                int p = -1;
                for (int r = 0; r < m && p < 0; ++r) if (((v[r >> 6] >> (r & 63)) & 1ull) && !ispiv[r]) p = r;
                if (p < 0) continue;
                if (B->nops == B->capops) {
                    B->capops = B->capops ? 2 * B->capops : 8;
                    B->oprow = (int*)realloc(B->oprow, (size_t)B->capops * sizeof(int));
                    B->opcol = (int*)realloc(B->opcol, (size_t)B->capops * sizeof(int));
                    B->opvec = (uint64_t*)realloc(B->opvec, (size_t)B->capops * nw * 8);
                }
                v[p >> 6] &= ~(1ull << (p & 63));                 /* the row operation leaves the pivot row itself alone */
                memcpy(&B->opvec[(size_t)B->nops * nw], v, (size_t)nw * 8);
                B->oprow[B->nops] = p; B->opcol[B->nops] = j; B->nops++;
                lsd_diag[0]++;
                ispiv[p] = 1;
            }
            B->nelim = B->nbits;
            LSD_REDUCED_SYNDROME(B);
            int ok = 1;
            for (int q = 0; q < B->nchecks && ok; ++q) {
                int r = B->checks[q];
                if (((z[r >> 6] >> (r & 63)) & 1ull) && !ispiv[r]) ok = 0;
            }
            B->valid = ok; B->stuck = 0;
        }
        /* next round: active and invalid clusters, smallest first (stable) */
        ninv = 0;
        for (int c = 0; c < nc; ++c) if (cl[c].active && !cl[c].valid) inv[ninv++] = c;
        for (int a = 1; a < ninv; ++a) {
            int x = inv[a], b2 = a - 1;
            while (b2 >= 0 && cl[inv[b2]].nbits > cl[x].nbits) { inv[b2 + 1] = inv[b2]; --b2; }
            inv[b2 + 1] = x;
        }
    }
    for (int c = 0; c < nc; ++c) {
        lsd_cl* B = &cl[c];
        if (B->active) {
            LSD_REDUCED_SYNDROME(B);
            if (d->osd_method >= 4 && d->osd_order > 0 && B->nbits > 0) {
                lsd_cluster_higher(d, B, z, nw, bit_owner /* scratch from here on */, ehat);
            } else
            for (int o = 0; o < B->nops; ++o) {
                int p = B->oprow[o];
                if ((z[p >> 6] >> (p & 63)) & 1ull) ehat[B->opcol[o]] = 1;
            }
        }
        free(B->cols); free(B->checks); free(B->oprow); free(B->opcol); free(B->opvec);
    }
    #undef LSD_REDUCED_SYNDROME
    free(cl); free(bit_owner); free(check_owner); free(isb); free(ispiv); free(v); free(z); free(inv); free(mlist);
}

/* decode one syndrome.  Returns 1 if BP converged.  llr_out (n doubles) = BP posteriors; iters_out = iterations run;
 * used_osd_out = 1 if the returned ehat came from OSD. */
int qo_bp_decode(const qo_bp* d, const uint8_t* syn, uint8_t* ehat, double* llr_out, int* iters_out, int* used_osd_out)
{
    int conv;
    double* llr = llr_out ? llr_out : (double*)malloc((size_t)d->n * sizeof(double));
    if (d->precision == 32) conv = bp_run_f32(d, syn, ehat, llr, iters_out);
    else conv = bp_run_f64(d, syn, ehat, llr, iters_out);
    int used = 0;
    if (!conv && d->osd) { if (d->osd_method >= 3) lsd_decode(d, syn, llr, ehat); else osd_decode(d, syn, llr, ehat); used = 1; }
    if (used_osd_out) *used_osd_out = used;
    if (!llr_out) free(llr);
    return conv;
}

/* ============================================================================================
 * Sliding-window loop: restates reference src/quits/decoder/sliding_window.py:162-186.
 *   per shot: acc = 0, carry = 0; for window k: s = det[F k m : (F k + W) m]; s[:m] ^= carry   (:168-169)
 *             e = decode_k(s)                                                                (:171)
 *             acc ^= L_k e[:ncommit_k] ; carry = U_k e[:ncommit_k]                          (:172-175)
 *             last window: s = det[F n_win m :]; s[:m] ^= carry; acc ^= L_last e             (:179-184)
 * L_k / U_k are given in CSC over the committed columns.
 * ============================================================================================ */
int qo_sw_decode(int n_windows, qo_bp* const* dec, const int32_t* row0 /*first detector of window*/,
                 const int32_t* ncommit,
                 const int64_t* const* Lptr, const int32_t* const* Lidx,      /* per window CSC of L (K x ncommit) */
                 const int64_t* const* Uptr, const int32_t* const* Uidx,      /* per window CSC of U (m x ncommit); NULL for last */
                 int m, int K, int D,
                 const uint8_t* det /*[N][D]*/, int64_t N, uint8_t* pred /*[N][K]*/,
                 int64_t* stats /* [n_windows][3]: converged windows, BP iterations, OSD calls */, int nthreads)
{
    int64_t* st = (int64_t*)calloc((size_t)n_windows * 3, sizeof(int64_t));
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel
#endif
    {
        int maxn = 0, maxm = 0;
        for (int k = 0; k < n_windows; ++k) { if (dec[k]->n > maxn) maxn = dec[k]->n; if (dec[k]->m > maxm) maxm = dec[k]->m; }
        uint8_t* e = (uint8_t*)malloc((size_t)maxn + 8);
        uint8_t* s = (uint8_t*)malloc((size_t)maxm + 8);
        uint8_t* carry = (uint8_t*)malloc((size_t)m + 8);
        double* llr = (double*)malloc((size_t)maxn * sizeof(double) + 8);
        int64_t* lst = (int64_t*)calloc((size_t)n_windows * 3, sizeof(int64_t));
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 8)
#endif
        for (int64_t i = 0; i < N; ++i) {
            uint8_t* acc = pred + (size_t)i * K;
            memset(acc, 0, (size_t)K);
            memset(carry, 0, (size_t)m);
            for (int k = 0; k < n_windows; ++k) {
                const qo_bp* d = dec[k];
                for (int r = 0; r < d->m; ++r) s[r] = det[(size_t)i * D + row0[k] + r] & 1;
                for (int r = 0; r < m && r < d->m; ++r) s[r] ^= carry[r];
                int it = 0, used = 0;
                int conv = qo_bp_decode(d, s, e, llr, &it, &used);
                lst[3 * k + 0] += conv; lst[3 * k + 1] += it; lst[3 * k + 2] += used;
                memset(carry, 0, (size_t)m);
                for (int j = 0; j < ncommit[k]; ++j) {
                    if (!e[j]) continue;
                    for (int64_t q = Lptr[k][j]; q < Lptr[k][j + 1]; ++q) acc[Lidx[k][q]] ^= 1;
                    if (Uptr[k]) for (int64_t q = Uptr[k][j]; q < Uptr[k][j + 1]; ++q) carry[Uidx[k][q]] ^= 1;
                }
            }
        }
#ifdef _OPENMP
#pragma omp critical
#endif
        for (int q = 0; q < n_windows * 3; ++q) st[q] += lst[q];
        free(e); free(s); free(carry); free(llr); free(lst);
    }
    if (stats) memcpy(stats, st, (size_t)n_windows * 3 * sizeof(int64_t));
    free(st);
    return 0;
}

int qo_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
