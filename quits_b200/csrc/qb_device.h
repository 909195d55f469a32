// quits_b200/csrc/qb_device.h -- argument blocks of the CUDA kernels (sm_100a) and their launch wrappers.
//
//   K1  frame_kernel   per-shot Pauli-frame propagation, 64 shots per bit-word      (frame.cu)
//   K3  bp_kernel_ms2 / bp_kernel_compact / bp_kernel   flooding BP of one window, one shot per CTA (bp.cu)
//   K3s bp_kernel_serial_slab   serial-schedule BP, one warp per shot, messages in a global slab (bp_serial.cu)
//   K4  osd_fast_kernel / osd_sort_kernel / osd_elim_kernel   OSD: selection or radix sort, GF(2) Gauss-Jordan, candidate sweeps (osd.cu)
//   K4L lsd_kernel, K4b osd_big_kernel   localized statistics decoding / OSD-0 of tall windows, one warp per shot (lsd.cu)
//   K2/K5 are fused into K3/K4: syndrome = slice(det) ^ carry on entry, H e == s as the stop test, L e / U e on exit.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "qb_host.h"

namespace qb {

// cudaFuncSetAttribute is per device: the "largest shared-memory size configured so far" caches are kept per device so that
// one process may drive several GPUs (the usual deployment is one process per GPU)
constexpr int kMaxDevices = 32;
inline int device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return d & (kMaxDevices - 1);
}

// ---------------------------------------------------------------------------------------------- K1
struct FrameArgs {
    const TapeOp* ops;
    int n_ops;
    const uint32_t* targets;
    const uint32_t* detptr;
    const uint32_t* detidx;
    const uint64_t* ctab;
    int n_qubits, n_det, n_obs;
    int ring;                 // power of two
    int DW, KW;               // u64 words per shot in det_rows / obs_rows
    uint64_t seed;
    uint64_t word0;           // global index of the first 64-shot word of this launch
    uint64_t n_words;
    uint64_t* det_words;      // scratch [n_words][DW*64]: word d of a 64-shot word holds detector d of its 64 shots (written once per
                              // detector by the frame kernel, transposed into det_rows at the end of the same launch)
    uint64_t* det_rows;       // [n_words*64][DW]   bit d of shot s = det_rows[s*DW + d/64] >> (d%64)
    uint64_t* obs_rows;       // [n_words*64][KW]
    // explicit-fault mode (noise instructions are replaced by the listed faults)
    int inject;
    const int32_t* inj_start; // [n_ops+1] fault range of every tape op
    const int32_t* inj_tgt;
    const int32_t* inj_code;
    const int64_t* inj_shot;  // shot index local to this launch
};
size_t frame_smem_per_warp(const FrameArgs& a);
size_t frame_scratch_bytes(const FrameArgs& a);       // bytes of the detector-word scratch a launch of a.n_words words needs
cudaError_t launch_frame(const FrameArgs& a, cudaStream_t st);

// bit rows <-> byte matrices (the reference API works on numpy bool arrays)
cudaError_t launch_unpack_bits(const uint64_t* rows, int words_per_row, int nbits, uint64_t n_rows, uint8_t* out, cudaStream_t st);
cudaError_t launch_pack_bits(const uint8_t* in, int nbits, uint64_t n_rows, uint64_t* rows, int words_per_row, cudaStream_t st);

// ---------------------------------------------------------------------------------------------- K3/K4
constexpr uint32_t kNoEdge = 0xFFFFFFFFu;
constexpr int kOsdSelCap = 512;      // columns handed to the OSD fast path per failed shot (at most)
constexpr int kOsdSelTarget = 96;    // ... and the count the selection aims for

struct WinDev {
    int rows, ncols, ncols_pad, RS, cw, ncommit;
    int row0;                 // first detector row (global) of the window
    int carry_rows;           // rows of the carry this window emits (0 for the last window)
    int KW;                   // u64 words per observable mask
    int rowsW32, nW32;
    const double* osd_wt;     // [ncols] log(1/p_j): weights of the higher-order OSD sweeps
    double bin_scale;         // OSD fast path: LLR -> selection bin scale, 10 / (smallest prior LLR of the window)
    int full_row_rank;        // GF(2) rank of the window matrix == rows (then OSD's answer does not depend on pivot-row order)
    int rank;                 // GF(2) rank of the window matrix
    const uint32_t* colE;     // [cw][ncols_pad]  (row << 8 | slot), kNoEdge when the column is shorter
    const float* llr0f;       // [ncols_pad]  prior LLRs log((1-p)/p), fp32 image (precision 32)
    const double* llr0d;      // [ncols_pad]  ... fp64 (precision 64)
    const uint64_t* lmask;    // [ncommit][KW]
    const int32_t* uptr;      // [ncommit+1]
    const uint16_t* uidx;
    const int32_t* cptr;      // [ncols+1] plain CSC of the window (OSD gather)
    const uint16_t* crow;
    const int32_t* rptr;      // [rows+1] CSR of the window, ascending columns (LSD growth candidates)
    const uint16_t* rcol;
    // compact form (cw <= 6, rows*RS < 65535, <= 1024 distinct priors): one 16-byte record per column
    //   rec = 6 x u16 message address (row*RS + slot, 0xFFFF = none), u16 prior index, u16 unused
    int compact;
    int n_ptab;
    uint32_t rs_magic;        // row = __umulhi(addr, rs_magic)  (= addr / RS exactly for addr < 65536)
    const uint4* colrec;      // [ncols_pad]
    const float* ptabf;       // [n_ptab] distinct prior LLRs
    const double* ptabd;
    const uint8_t* rlen;      // [rows] row weights
    const float2* rsum0f;     // [rows] iteration-1 row summaries (min1, min2 of the prior LLRs along the row)
    const double2* rsum0d;
    const uint8_t* neg0;      // [rows] parity of #{prior LLR <= 0} along the row
    // serial schedule on a global message slab (bp_serial.cu, bp_kernel_serial_slab): any column weight <= 16
    int nnz;
    int ss_nsteps, ss_lpc;    // steps of one sweep; lanes per column (6, 8 or 16 => 5, 4 or 2 independent columns per step)
    const uint2* ss_rec;      // [ss_nsteps + 4][32]  x = CSR position of the lane's edge (no edge: the lane's dummy slot nnz + lane),
                              //   y = row (no edge: the dummy row `rows`) | extra << 16, extra = column index on the first lane of a
                              //   column's group (0xFFFF: no column)
    const float* ss_v0f;      // [nnz] initial message of every CSR edge: the prior LLR of its column (product-sum: tanh(LLR / 2))
    const double* ss_v0d;
    const float* ss_s0f;      // product-sum: [nnz] suffix products of the initial factors along each row
    const double* ss_s0d;
    const float2* ss_rsum0f;  // min-sum: [rows] (min1, min2) of the prior LLRs' magnitudes along the row, and the parity of #{LLR <= 0}
    const double2* ss_rsum0d;
    const uint8_t* ss_neg0;
    int chunk_end[7];         // bp_kernel_ms2: chunk_end[W] = first 32-record chunk whose heaviest column is lighter than W
    int unit_alpha;           // every iteration's min-sum scaling factor is exactly 1 (the multiply is skipped)
};

struct BatchDev {
    int n_shots;
    const uint32_t* det32;    // packed detector rows viewed as u32; det_stride32 words per shot (>= 2*DW + 1)
    int det_stride32;
    int in_carry_rows;        // rows of the incoming carry (0 for the first window)
    uint32_t* carry;          // [n][carry_stride32]
    int carry_stride32;
    uint64_t* acc;            // [n][KW]  accumulated observable prediction
    void* llr_buf;            // [n][llr_stride] posteriors, float or double according to the precision
    size_t llr_stride;
    int llr_esize;            // 4 or 8
    uint16_t* order_alt;      // optional [n][llr_stride]: where the full OSD sort puts the column order; NULL => over the posterior row
    void* vscratch;           // VGLOBAL only: [grid][rows*RS] message slabs
    uint32_t* syn_buf;        // [n][syn_stride32]   post-carry syndrome of the shots handed to OSD
    int syn_stride32;
    int* fail_list;           // [n]
    int* fail_count;          // [1]
    void* sel_key;            // [n][kOsdSelCap] order keys (u32 / u64 by precision) of the least reliable columns of a failed shot
    uint16_t* sel_idx;        // [n][kOsdSelCap] their column indices (unordered; the OSD warp sorts them)
    int* sel_cnt;             // [n] tier-1 count | tier-2 count << 16 (tier 1 from the front of the row, tier 2 from its back); NULL when OSD is off
    int* fast_next;           // [1] work counter of the persistent OSD fast-path grid
    int* ovf_list;            // [n] shots the fast path hands to the full sort + elimination
    int* ovf_count;           // [1]
    int* sort_next;           // [1] work counter of the persistent OSD sort grid
    int* osd_next;            // [1] work counter of the persistent OSD elimination grid
    int* bp_next;             // [1] shot queue of the persistent bp_kernel_ms2 grid (NULL: one CTA per shot, shot = blockIdx.x)
    unsigned long long* stats;// [8] converged windows, BP iterations, OSD calls, OSD columns examined, OSD pivots, max OSD columns, fast-path overflows
    uint32_t* ehat_out;       // optional [n][ehat_stride32]  (pre-zeroed)
    int ehat_stride32;
    int32_t* iters_out;       // optional [n]
    uint8_t* conv_out;        // optional [n]
    int write_llr_always;
    int commit_unconverged;   // no OSD / LSD behind BP: commit BP's hard decision of unconverged shots too (ldpc's BpDecoder semantics)
    int osd_method, osd_order; // 0 osd_0 | 1 osd_e | 2 osd_cs; order 0 = OSD-0 | 3 lsd_0
    void* lsd_scratch;        // LSD only: [grid] slabs of lsd_slab bytes (bit owners, column-order links, operation vectors)
    size_t lsd_slab;
    int lsd_cols;             // columns of the widest window (fixes the slab layout)
    int lsd_method, lsd_order; // beyond order 0: 1 = lsd_e, 2 = lsd_cs candidate sweep inside every cluster
    void* sort_scratch;       // wide windows: [grid] slabs of osd_sort_slab_bytes for the radix sort's keys and index buffers
};

struct BpParams {
    int max_iter;
    int method;               // 0 minimum_sum | 1 product_sum
    const double* alpha;      // [max_iter+1]  scaling factor of iteration it (index it), rounded to the precision in the kernel
};

// precision: 32 or 64 (message / posterior type).  vglobal: messages in a global scratch slab instead of shared memory.
constexpr uint32_t kNoAddr = 0xFFFFu;
size_t bp_smem_bytes(const WinDev& w, int precision, bool vglobal);
size_t bp_slab_bytes(const WinDev& w, int precision);       // bytes of one CTA's global message slab (VGLOBAL)
int bp_threads(int precision);
bool bp_supports(const WinDev& w, int method, bool vglobal);
int bp_persistent_grid(const WinDev& w, int precision, bool vglobal, int method);
bool bp_ms2_enabled();       // flooding min-sum through bp_kernel_ms2 (default) or bp_kernel_compact (QB_BP_MS2=0, A/B measurements)
cudaError_t bp_configure(const WinDev& w, int precision, bool vglobal, int method);
cudaError_t launch_bp(const WinDev& w, const BatchDev& b, const BpParams& p, int precision, bool vglobal, int grid, cudaStream_t st);

// serial schedule, one warp per shot, messages in a global slab per CTA (b.vscratch: [grid] slabs of bp_serial_slab_bytes)
size_t bp_serial_slab_smem_bytes(const WinDev& w, int precision, int method);
size_t bp_serial_slab_bytes(const WinDev& w, int precision, int method);
bool bp_serial_slab_supported(const WinDev& w, int precision, int method);
cudaError_t bp_serial_slab_configure(const WinDev& w, int precision, int method);
int bp_serial_slab_ctas_per_sm(const WinDev& w, int precision, int method);
cudaError_t launch_bp_serial_slab(const WinDev& w, const BatchDev& b, const BpParams& p, int precision, int grid, cudaStream_t st);
cudaError_t launch_serial_slab_ps_tables(const WinDev& w, int precision, void* v0, void* s0, cudaStream_t st);

size_t osd_sort_smem_bytes(const WinDev& w, int precision);
size_t osd_sort_slab_bytes(const WinDev& w, int precision);
size_t osd_elim_smem_bytes(const WinDev& w, bool hi);
size_t osd_fast_smem_bytes(const WinDev& w);
bool osd_supported(const WinDev& w, int precision);
bool osd_tall_hi_supported(const WinDev& w, int precision);      // osd_e / osd_cs of order > 0 on 768 < checks <= 2304: T in a global slab
size_t osd_tall_hi_slab_bytes(const WinDev& w);
cudaError_t osd_tall_hi_configure(const WinDev& w, int precision);
cudaError_t osd_configure(const WinDev& w, int precision);
cudaError_t osd_sort_configure(const WinDev& w, int precision);
cudaError_t launch_osd_fast(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st);
cudaError_t launch_osd_sort(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st);
cudaError_t launch_osd_elim(const WinDev& w, const BatchDev& b, bool hi, int grid, cudaStream_t st);

// K4L (lsd.cu): localized statistics decoding, order 0, one warp per failed shot (persistent grid)
size_t lsd_smem_bytes(const WinDev& w);
size_t lsd_slab_bytes(int cols_cap, int max_rows);
bool lsd_supported(const WinDev& w);
cudaError_t lsd_configure(const WinDev& w, int precision, bool hi);
cudaError_t launch_lsd(const WinDev& w, const BatchDev& b, int precision, int grid, cudaStream_t st);
// OSD-0 for windows taller than the shared-memory elimination takes (768 < checks <= 3072): same slab machinery as LSD
size_t osd_big_smem_bytes(const WinDev& w);
size_t osd_big_slab_bytes(int max_rows);
bool osd_big_supported(const WinDev& w);
cudaError_t osd_big_configure(const WinDev& w);
cudaError_t launch_osd_big(const WinDev& w, const BatchDev& b, int grid, cudaStream_t st);

// ---------------------------------------------------------------------------------------------- results
// pred[n][K] int64 from acc bits; counts[0] += shots whose prediction differs from obs_rows in any observable,
// counts[1+k] += mismatches of observable k.
cudaError_t launch_expand_pred(const uint64_t* acc, int KW, int K, uint64_t n, int64_t* pred, cudaStream_t st);
cudaError_t launch_count(const uint64_t* acc, const uint64_t* obs_rows, int KW, int K, uint64_t n, unsigned long long* counts,
                         cudaStream_t st);

}  // namespace qb
