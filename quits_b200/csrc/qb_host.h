// quits_b200/csrc/qb_host.h -- host-side (CPU, set-up only) data structures of the B200 engine.
//
// Everything here runs once per (circuit, window plan); the per-shot work is in frame.cu / bp.cu / osd.cu.
// The semantics follow the reference's boundary, not its code:
//   * Stim-text dialect emitted by reference src/quits/circuit.py:58-279
//   * circuit -> detector error model: what reference decoder/base.py:151 obtains from
//     stim.Circuit.detector_error_model(decompose_errors=False)
//   * DEM -> (H, L, priors): reference decoder/base.py:74-127
//   * window slicing: reference decoder/base.py:134-190 and decoder/sliding_window.py:130-141
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace qb {

enum OpKind : int32_t {
    OP_R = 0, OP_RX, OP_H, OP_CX, OP_M, OP_MX, OP_MR, OP_XERR, OP_ZERR, OP_DEP1, OP_DEP2, OP_DET, OP_OBS
};

struct value_error : std::runtime_error { using std::runtime_error::runtime_error; };          // -> Python ValueError
struct unsupported_error : std::runtime_error { using std::runtime_error::runtime_error; };    // -> NotImplementedError

// One flattened instruction (REPEAT unrolled, TICK dropped).  For OP_DET / OP_OBS the targets are ABSOLUTE
// measurement indices and arg is the detector / observable index.
struct FlatOp {
    int32_t kind;
    double arg;
    std::vector<int32_t> targets;
};

struct FlatCircuit {
    std::vector<FlatOp> ops;
    int n_qubits = 0, n_meas = 0, n_det = 0, n_obs = 0;
    int64_t n_sites = 0;          // noise sites (every noise instruction starts at a multiple of 4)
    int max_lookback = 0;         // largest rec[-k] distance at the point of use
};

void parse_flatten(const char* text, size_t len, FlatCircuit& out);

// ---- device tape --------------------------------------------------------------------------------------------
// A tape op is a conflict-free slice of a flat op (no qubit twice), so that lanes may process its targets in
// parallel; DETECTOR runs are merged into one block.
struct TapeOp {
    int32_t kind;
    int32_t n;          // targets (R,H,M..), pairs (CX, DEP2), detectors (DET), rec entries (OBS), sites (noise)
    uint32_t t0;        // offset into targets[] (gates/noise/OBS) or into detptr[] (DET)
    uint32_t aux;       // M/MX/MR: first measurement index; noise: first site id; DET: first detector id; OBS: observable id
    uint32_t thr;       // noise: level-1 threshold T1(p)
    int32_t tab;        // noise: index of the 64-entry count table
    int32_t flat;       // index of the flat op this slice came from
    int32_t foff;       // target (or pair) offset of this slice inside the flat op
};

struct Tape {
    std::vector<TapeOp> ops;
    std::vector<uint32_t> targets;      // qubit ids / absolute measurement indices (OBS)
    std::vector<uint32_t> detptr;       // CSR over detectors of all DET blocks (absolute offsets into detidx)
    std::vector<uint32_t> detidx;       // absolute measurement indices
    std::vector<uint64_t> ctab;         // [n_tables][64]
    std::vector<double> tab_p;
    int ring = 0;                       // measurement ring size (power of two > max lookback + widest measure op)
};

void build_tape(const FlatCircuit& fc, Tape& tape);
void noise_tables(double p, uint32_t* t1, uint64_t* c /*[64]*/);

// ---- detector error model -----------------------------------------------------------------------------------
struct Dem {
    int n_det = 0, n_obs = 0;
    std::vector<double> probs;
    std::vector<std::vector<int32_t>> dets, obs;
    std::vector<int32_t> rep_op, rep_tgt, rep_code;     // representative fault of each error (flat op, target/pair, Pauli code)
};

void analyze(const FlatCircuit& fc, Dem& dem);

// ---- check matrix + window plan -----------------------------------------------------------------------------
struct CheckMatrix {                    // reference decoder/base.py:74-127 (columns in first-sighting order)
    int n_det = 0, n_obs = 0;
    std::vector<std::vector<int32_t>> col_dets;    // sorted detector ids per column
    std::vector<std::vector<int32_t>> col_obs;     // sorted observable ids per column (first sighting wins)
    std::vector<double> priors;
    int n_detless = 0;                             // errors without detectors (reference prints them, base.py:114-115)
};

void dem_to_matrix(const Dem& dem, CheckMatrix& cm);

struct Window {
    int row0 = 0, rows = 0;             // detector rows [row0, row0+rows) of the global matrix
    int col0 = 0, ncols = 0;            // global columns [col0, col0+ncols)
    int ncommit = 0;                    // committed prefix of the window's columns
    int urow0 = 0, urows = 0;           // carry rows: global rows [urow0, urow0+urows) (urows == 0 for the last window)
    std::vector<int64_t> cptr;          // CSC of H_window (window-relative rows, ascending)
    std::vector<int32_t> crow;
    std::vector<double> priors;
    std::vector<int64_t> lptr;          // CSC of L over the committed columns
    std::vector<int32_t> lidx;
    std::vector<int64_t> uptr;          // CSC of U (rows relative to urow0) over the committed columns
    std::vector<int32_t> uidx;
};

struct WindowPlan {
    int m = 0, K = 0, D = 0, W = 0, F = 0, num_rounds = 0, n_cor = 0;
    bool whole_history = false;         // W larger than the number of rounds (reference warns, sliding_window.py:140)
    std::vector<Window> windows;
};

// ---- shared-memory layout of a window's BP messages (layout.cpp) ---------------------------------------------
struct BpLayout {
    std::vector<int> order;             // record r holds column order[r]; weights non-increasing
    std::vector<int> slot;              // slot of every edge inside its row, indexed like Window::crow
    long rsum_excess = 0, v_excess = 0; // residual bank-class collisions after the search (0 = conflict free)
};
// rs: row stride of the message array; precision 32 / 64 selects the bank-class geometry
// rb > 0 overrides the number of row-summary bank classes (default: 8 in fp64, 16 in fp32 -- the (min1, min2) pairs of bp_kernel_compact)
void optimize_bp_layout(const Window& hw, int rs, int precision, BpLayout& out, int rb = 0);
void layout_wavefronts(const Window& hw, int rs, int precision, const BpLayout& lay, double* v_ratio, double* r_ratio, int rb = 0);

// n_cor_override < 0: derive the number of sliding windows as sliding_window.py:130-141 does
void plan_windows(const CheckMatrix& cm, int m, int W, int F, int n_cor_override, WindowPlan& plan);

}  // namespace qb
