/*
 * include/quits_b200.h -- C ABI of the B200-native QUITS Monte-Carlo engine (libquits_b200.so).
 *
 * The reference (mkangquantum/quits @ b8b3b44) has no FFI of its own on this path: its seams are Python call
 * signatures into the C++ wheels stim and ldpc.  Each entry point below names the reference interface it
 * replaces (file:line relative to the reference tree).  Plain C: opaque handles, caller-owned buffers, int
 * status codes, qb_last_error() for the message.  One context per (process, device); a context and the objects
 * created from it must not be used from two threads at once.
 *
 * Status codes: 0 ok | 1 value error (Python ValueError) | 2 not implemented (NotImplementedError)
 *               3 CUDA / runtime error | 4 bad argument.
 * Bit order everywhere: bit b of a packed row lives in word b/64 at position b%64.
 */
#ifndef QUITS_B200_H
#define QUITS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB_OK 0
#define QB_EVALUE 1
#define QB_ENOTIMPL 2
#define QB_ECUDA 3
#define QB_EARG 4

typedef struct qb_ctx qb_ctx;
typedef struct qb_circuit qb_circuit;
typedef struct qb_dem qb_dem;
typedef struct qb_plan qb_plan;
typedef struct qb_sw qb_sw;

const char* qb_last_error(void);
int qb_version(void);
/* "quits_b200 abi=.. arch=sm_100a nvcc=X.Y.Z fmad=off lineinfo=on src=<hash>": how this library was built; <hash> is the sha256
 * prefix of the sources it was compiled from (quits_b200/build.py source_hash()), so a loaded .so can be matched to a source tree. */
const char* qb_build_info(void);
/* number of CUDA devices visible (0 without a GPU; never fails) */
int qb_device_count(void);

/* ------------------------------------------------------------------------------------------------ context */
int qb_ctx_create(int device, qb_ctx** out);
void qb_ctx_destroy(qb_ctx* ctx);
int qb_ctx_synchronize(qb_ctx* ctx);

/* Page-locked host buffers for the arrays that cross the boundary (detection events, predictions): a device<->host copy
 * from pinned memory runs at PCIe speed instead of being staged through the driver's bounce buffer.  The Python layer
 * hands these out as numpy arrays and recycles them.  No reference counterpart (stim / ldpc return ordinary numpy memory). */
int qb_host_alloc(size_t bytes, void** out);
void qb_host_free(void* p);

/* ------------------------------------------------------------------------------------------------ circuit
 * Replaces stim.Circuit(text) as the reference's builders call it (src/quits/qldpc_code/bb.py:301,
 * circuit_construction/cardinal.py:267, zxcoloration.py:270).  Host only; accepts the Stim-text dialect emitted
 * by src/quits/circuit.py:58-279 (R RX H CX M MX MR, X_ERROR Z_ERROR DEPOLARIZE1 DEPOLARIZE2, DETECTOR,
 * OBSERVABLE_INCLUDE, TICK, REPEAT), also in stim's canonical (fused) printing.  PAULI_CHANNEL_1/2 -> QB_ENOTIMPL. */
typedef struct {
    int32_t n_qubits, n_measurements, n_detectors, n_observables;
    int64_t n_flat_ops;        /* instructions after REPEAT unrolling (TICK dropped) */
    int64_t n_tape_ops;        /* conflict-free slices executed by the frame kernel */
    int64_t n_noise_sites;
    int32_t ring;              /* measurement ring size used on the device */
} qb_circuit_info;

int qb_circuit_parse(const char* stim_text, size_t len, qb_circuit** out);
void qb_circuit_free(qb_circuit* c);
int qb_circuit_get_info(const qb_circuit* c, qb_circuit_info* info);
/* flattened op list (for tests and for addressing faults): kind per op, CSR of targets.  Pass NULL to skip. */
int qb_circuit_flat(const qb_circuit* c, int32_t* kind, double* arg, int64_t* tstart /*[n_flat_ops+1]*/, int32_t* targets,
                    int64_t* n_targets_out);

/* ------------------------------------------------------------------------------------------------ sampler
 * Replaces circuit.compile_detector_sampler(seed).sample(shots, separate_observables=True)
 * (src/quits/simulation.py:22-27).  Shots are numbered globally: shot s of a run with a given seed is the same
 * bits whatever the batch split or the GPU.  shot0 must be a multiple of 64.
 * qb_sample        -> det[n_shots][n_detectors], obs[n_shots][n_observables] as 0/1 bytes (numpy bool_ layout)
 * qb_sample_packed -> det_rows[n_shots][ceil(D/64)], obs_rows[n_shots][ceil(K/64)] as u64 bit rows            */
int qb_sample(qb_ctx* ctx, qb_circuit* c, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint8_t* det, uint8_t* obs);
int qb_sample_packed(qb_ctx* ctx, qb_circuit* c, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint64_t* det_rows,
                     uint64_t* obs_rows);
/* Explicit-fault mode (parity tests): the noise instructions are replaced by the listed faults.  Fault f flips
 * Pauli code[f] (bit0 X_a, bit1 Z_a, bit2 X_b, bit3 Z_b) on target/pair tgt[f] of flat op op[f] in shot shot[f]. */
int qb_sample_faults(qb_ctx* ctx, qb_circuit* c, int64_t n_faults, const int32_t* op, const int32_t* tgt, const int32_t* code,
                     const int64_t* shot, uint64_t n_shots, uint8_t* det, uint8_t* obs);

/* ------------------------------------------------------------------------------------------------ DEM
 * Replaces circuit.detector_error_model(decompose_errors=False) (src/quits/decoder/base.py:151) and, through
 * qb_dem_matrix, detector_error_model_to_matrix (base.py:74-127).  Host only. */
int qb_dem_from_circuit(const qb_circuit* c, qb_dem** out);
/* a DEM produced elsewhere (e.g. by stim itself): errors in the order given, CSR of detector / observable ids */
int qb_dem_from_errors(int32_t n_detectors, int32_t n_observables, int64_t n_errors, const double* probs, const int64_t* det_ptr,
                       const int32_t* det_idx, const int64_t* obs_ptr, const int32_t* obs_idx, qb_dem** out);
void qb_dem_free(qb_dem* d);
/* sizes: [n_detectors, n_observables, n_errors, nnz_det, nnz_obs, n_columns, nnz_H, nnz_L, n_detectorless] */
int qb_dem_sizes(const qb_dem* d, int64_t sizes[9]);
/* stim-ordered error list (CSR) + one representative circuit fault per error */
int qb_dem_errors(const qb_dem* d, double* probs, int64_t* det_ptr, int32_t* det_idx, int64_t* obs_ptr, int32_t* obs_idx,
                  int32_t* rep_op, int32_t* rep_tgt, int32_t* rep_code);
/* merged check matrix H (CSC, sorted rows), observable matrix L (CSC) and priors, columns in first-sighting order */
int qb_dem_matrix(const qb_dem* d, int64_t* h_ptr, int32_t* h_idx, int64_t* l_ptr, int32_t* l_idx, double* priors);

/* ------------------------------------------------------------------------------------------------ decoder
 * Replaces the body of sliding_window_circuit_mem (src/quits/decoder/sliding_window.py:130-188): window plan
 * (:130-141 + decoder/base.py:149-188), one BP(+OSD) decoder per window (:146-153) and the per-shot loop (:162-186)
 * with ldpc.BpOsdDecoder.decode inside (:171,182).  Option names follow the reference's kwargs (decoder/bposd.py:74-83;
 * decoder/bplsd.py:74-83 for osd_method 3 = ldpc.BpLsdDecoder's post-processing).
 * Window sizes: BP up to 8192 checks and 65534 fault columns, column weight <= 16, both bp_methods and both schedules (flooding:
 * messages in shared memory when they fit, else in a global slab; serial: messages in a global slab, row summaries in shared
 * memory); OSD-0 and LSD (any order) up to 3072 checks; osd_e / osd_cs with order > 0 up to 2304 checks; anything beyond these
 * sizes: QB_ENOTIMPL. */
typedef struct {
    int32_t bp_method;          /* 0 'minimum_sum' | 1 'product_sum' */
    int32_t schedule;           /* 0 'parallel' (flooding) | 1 'serial' (columns in index order, no random reshuffle) */
    int32_t max_iter;           /* 0 => number of columns (ldpc convention) */
    double ms_scaling_factor;   /* 0.0 => 1 - 2^-iteration */
    int32_t osd_method;         /* 0 'osd_0' | 1 'osd_e' | 2 'osd_cs' | 3 'lsd_0' | 4 'lsd_e' | 5 'lsd_cs' (BpLsdDecoder) | -1 no post-processing */
    int32_t osd_order;          /* 0 = OSD-0 / LSD-0 whatever the method; osd_e, lsd_e: <= 12; osd_cs, lsd_cs: <= 32 (for LSD this is lsd_order) */
    int32_t precision;          /* 64 (default when 0): messages in fp64 as ldpc computes; 32: fp32 messages */
    int32_t capacity;           /* shots per device batch; 0 => default */
    int32_t profile;            /* 1: time the kernel classes with CUDA events on the launching stream */
    int32_t lanes;              /* concurrent sub-batches per device batch (streams); 0 => default (1; 4 with LSD, whose long-tailed
                                   launches overlap the BP kernel of the other sub-batches) */
} qb_bp_opts;

typedef struct {
    int64_t shots;
    int64_t windows;            /* shot-windows decoded */
    int64_t bp_converged;       /* ... of which BP converged */
    int64_t bp_iterations;      /* total BP iterations run */
    int64_t osd_calls;
    int64_t bp_launches, osd_launches, frame_launches, other_launches;
    double frame_ms, bp_ms, osd_ms, total_ms;     /* CUDA-event time on the launching stream (profile = 1) */
    /* algorithmic bytes of the launches above: BP = iterations x 4 x nnz x sizeof(message) + syndrome in +
     * commit/carry out; OSD = 2 x rows x 8 ceil(cols/64) per call; frame = packed detector + observable rows written */
    double bp_alg_bytes, osd_alg_bytes, frame_alg_bytes;
    /* OSD work actually done (the elimination stops early, osd.cu): columns examined, pivots taken, worst case */
    int64_t osd_columns, osd_pivots, osd_max_columns;
    int64_t osd_overflows;      /* shots the fast OSD path handed to the full sort */
    double bp_edge_iters;       /* sum over shot-windows of BP iterations run x edges of the window (the unit of BP work) */
} qb_stats;

/* Window plan (host only): spacetime() of decoder/base.py:134-190.  n_cor < 0 derives the number of sliding windows
 * from D, m, W, F as sliding_window.py:130-141 does; n_cor >= 0 takes the caller's num_cor_rounds. */
int qb_plan_create(const qb_dem* d, int32_t m /* hz.shape[0] */, int32_t W, int32_t F, int32_t n_cor, qb_plan** out);
/* A window plan given explicitly (host only) -- used for the phenomenological sliding window, whose windows are built from
 * hz by Kronecker products instead of from a circuit DEM (src/quits/decoder/sliding_window.py:14-101, matrices :56-69).
 * dims[k] = [row0, rows, (unused), ncols, ncommit, nnz, nnz_L, nnz_U, urow0, urows] as returned by qb_plan_window; the CSC
 * arrays are the windows' arrays concatenated in order (each window's pointer array starts at 0 and has ncols+1 resp.
 * ncommit+1 entries).  Window k decodes detector rows [row0, row0+rows), XORs the carry of window k-1 into its first urows
 * rows, commits L e[:ncommit] into the observable prediction and hands U e[:ncommit] (urows rows) to window k+1. */
int qb_plan_create_explicit(int32_t m, int32_t K, int32_t D, int32_t n_windows, const int64_t* dims, const int64_t* h_ptr,
                            const int32_t* h_idx, const double* priors, const int64_t* l_ptr, const int32_t* l_idx,
                            const int64_t* u_ptr, const int32_t* u_idx, qb_plan** out);
void qb_plan_free(qb_plan* p);
/* info: [n_windows, m, K, D, W, F, num_rounds, whole_history] */
int qb_plan_info(const qb_plan* p, int64_t info[8]);
/* window k: dims = [row0, rows, col0, ncols, ncommit, nnz, nnz_L, nnz_U, urow0, urows]; arrays may be NULL */
int qb_plan_window(const qb_plan* p, int32_t k, int64_t dims[10], int64_t* h_ptr, int32_t* h_idx, double* priors, int64_t* l_ptr,
                   int32_t* l_idx, int64_t* u_ptr, int32_t* u_idx);

/* Diagnostic (host only): the shared-memory layout qb_sw_create would choose for window k at the given message precision
 * (32 / 64) -- record order[ncols], slot[nnz] of every edge inside its row (CSC order of qb_plan_window) -- and the predicted
 * shared-memory passes per request relative to the conflict-free minimum: ratios = [message gather with the naive layout,
 * row-summary gather naive, message gather chosen, row-summary gather chosen].  No reference counterpart (ldpc has no GPU). */
int qb_plan_layout(const qb_plan* p, int32_t k, int32_t precision, int32_t* order, int32_t* slot, double ratios[4]);

int qb_sw_create(qb_ctx* ctx, const qb_plan* plan, const qb_bp_opts* opts, qb_sw** out);
/* one window given explicitly: ldpc.BpOsdDecoder(pcm, channel_probs=priors, ...) as built at sliding_window.py:146-153 */
int qb_sw_create_single(qb_ctx* ctx, int32_t rows, int32_t cols, const int64_t* indptr, const int32_t* indices, const double* priors,
                        const qb_bp_opts* opts, qb_sw** out);
void qb_sw_free(qb_sw* sw);
/* det[n][D] 0/1 bytes (host) -> pred[n][K] int64 (host), the dtype the reference returns (sliding_window.py:160) */
int qb_sw_decode(qb_sw* sw, const uint8_t* det, uint64_t n, int64_t* pred, qb_stats* stats);
int qb_sw_decode_packed(qb_sw* sw, const uint64_t* det_rows, uint64_t n, uint64_t* pred_rows, qb_stats* stats);
/* seam B3 (one decode per syndrome, batched): syndromes[n][rows] bytes -> ehat[n][cols] bytes, posteriors, iterations */
int qb_bp_decode_batch(qb_sw* single, const uint8_t* syndromes, uint64_t n, uint8_t* ehat, double* llr, int32_t* iters,
                       uint8_t* converged);

/* ------------------------------------------------------------------------------------------------ fused run
 * sample -> sliding-window decode -> compare, entirely on the device (get_stim_mem_result + sliding_window_*_mem +
 * the caller's pL reduction, reference tests/test_sliding_window.py:83).  counts[0] += shots with any observable
 * mispredicted, counts[1+k] += mispredictions of observable k. */
int qb_mc_run(qb_ctx* ctx, qb_circuit* c, qb_sw* sw, uint64_t seed, uint64_t shot0, uint64_t n_shots, uint64_t* counts /*[1+K]*/,
              qb_stats* stats);

#ifdef __cplusplus
}
#endif
#endif
