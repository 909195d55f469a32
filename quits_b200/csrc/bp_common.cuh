// quits_b200/csrc/bp_common.cuh -- device helpers shared by the BP kernels (bp.cu, bp_serial.cu): arithmetic traits, the window
// syndrome slice, the commit / hand-off to OSD at the end of a shot.  Included inside an anonymous namespace per translation unit.
#pragma once
#include <cfloat>

#include "qb_device.h"

namespace qb {

namespace {

template <typename R> struct Real;
template <> struct Real<float> {
    using pair = float2;
    static __device__ __forceinline__ float prior(const WinDev& w, int j) { return __ldg(w.llr0f + j); }
    static __device__ __forceinline__ float big() { return FLT_MAX; }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ pair mk(float a, float b) { return make_float2(a, b); }
};
template <> struct Real<double> {
    using pair = double2;
    static __device__ __forceinline__ double prior(const WinDev& w, int j) { return __ldg(w.llr0d + j); }
    static __device__ __forceinline__ double big() { return DBL_MAX; }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ pair mk(double a, double b) { return make_double2(a, b); }
};

__host__ __device__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- hand-off to OSD: the posteriors' 32-bin histogram (built while the last iteration stores them) selects the least
// reliable columns -- tier 1: the smallest bin prefix with >= kOsdSelTarget columns; tier 2: the further bins that still fit
// kOsdSelCap columns in total -- and the CTA writes their
// (order key, column) pairs to HBM, so the OSD warp never scans the full posterior vector (osd.cu, fast path).
constexpr int kSelWords = 36;        // 32 bins, tier-1 count, tier-2 count, the two boundary bins

template <typename R>
__device__ __forceinline__ int llr_bin(const R v, const R scale) {
    return v > R(0) ? 1 + static_cast<int>(fmin(static_cast<double>(v * scale), 30.0)) : 0;
}
__device__ __forceinline__ uint32_t order_key_of(float f) {
    const uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t order_key_of(double f) {
    const uint64_t u = static_cast<uint64_t>(__double_as_longlong(f + 0.0));
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

template <typename R, int NT>
__device__ __forceinline__ void select_for_osd(const WinDev& w, const BatchDev& b, int shot, int tid, uint32_t* hist) {
    using KeyT = decltype(order_key_of(R(0)));
    if (tid < 32) {
        uint32_t cum = hist[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, cum, o);
            if (tid >= o) cum += t;
        }
        // tier 1: the smallest bin prefix with >= kOsdSelTarget columns; tier 2: everything else that still fits the buffer
        const uint32_t fits = __ballot_sync(0xFFFFFFFFu, cum <= static_cast<uint32_t>(kOsdSelCap));
        const int b2 = fits ? 31 - __clz(fits) : -1;
        const uint32_t enough = __ballot_sync(0xFFFFFFFFu, cum >= static_cast<uint32_t>(kOsdSelTarget));
        int b1 = enough ? __ffs(enough) - 1 : 31;
        if (b1 > b2) b1 = b2;
        if (tid == 0) { hist[32] = 0; hist[33] = 0; hist[34] = static_cast<uint32_t>(b1 + 1); hist[35] = static_cast<uint32_t>(b2 + 1); }
    }
    __syncthreads();
    const int b1 = static_cast<int>(hist[34]) - 1, b2 = static_cast<int>(hist[35]) - 1;
    const R scale = static_cast<R>(w.bin_scale);
    const R* llr = reinterpret_cast<const R*>(b.llr_buf) + static_cast<size_t>(shot) * b.llr_stride;
    KeyT* gkey = reinterpret_cast<KeyT*>(b.sel_key) + static_cast<size_t>(shot) * kOsdSelCap;
    uint16_t* gidx = b.sel_idx + static_cast<size_t>(shot) * kOsdSelCap;
    for (int j = tid; j < w.ncols; j += NT) {
        const R v = llr[j];
        const int bin = llr_bin<R>(v, scale);
        if (bin <= b2) {                               // tier 1 fills the buffer from the front, tier 2 from the back
            const uint32_t pos = bin <= b1 ? atomicAdd(&hist[32], 1u) : static_cast<uint32_t>(kOsdSelCap - 1) - atomicAdd(&hist[33], 1u);
            gkey[pos] = order_key_of(v);
            gidx[pos] = static_cast<uint16_t>(j);
        }
    }
    __syncthreads();
    if (tid == 0) b.sel_cnt[shot] = static_cast<int>(hist[32] | (hist[33] << 16));
}

// ---- window syndrome: detector bits [row0, row0+rows) of this shot, first rows XOR the carry (sliding_window.py:168-169)
__device__ __forceinline__ void load_syndrome(const WinDev& w, const BatchDev& b, int shot, int tid, uint32_t* syn, uint32_t* accs,
                                              uint32_t* car) {
    const int carryW = (w.carry_rows + 31) / 32;
    const int nt = static_cast<int>(blockDim.x);             // strided: the warp-per-shot kernels have fewer threads than words
    const uint32_t* d = b.det32 + static_cast<size_t>(shot) * b.det_stride32;
    for (int i = tid; i < w.rowsW32; i += nt) {
        const int bit = w.row0 + 32 * i;
        const int wd = bit >> 5, sh = bit & 31;
        uint32_t v = __ldg(d + wd) >> sh;
        if (sh) v |= __ldg(d + wd + 1) << (32 - sh);
        const int left = w.rows - 32 * i;
        if (left < 32) v &= (1u << left) - 1u;
        if (32 * i < b.in_carry_rows) v ^= b.carry[static_cast<size_t>(shot) * b.carry_stride32 + i];
        syn[i] = v;
    }
    for (int i = tid; i < 2 * w.KW; i += nt) accs[i] = 0;
    for (int i = tid; i <= carryW; i += nt) car[i] = 0;
}

// ---- after BP: commit acc ^= L e[:ncommit], carry = U e[:ncommit] (sliding_window.py:172-175), or hand the shot to OSD
// hard decisions come either as a per-thread mask (bit k <-> column / record tid + k*NT) or, when `ebits` is given, as a bit
// array over the columns in shared memory (serial kernel)
template <typename R, int NT, bool RECORDS = false>
__device__ __forceinline__ void finish_shot(const WinDev& w, const BatchDev& b, int shot, int tid, bool conv, int it, uint32_t hmask,
                                            const uint32_t* syn, uint32_t* accs, uint32_t* car, uint32_t* hist,
                                            const uint32_t* ebits = nullptr) {
    const int carryW = (w.carry_rows + 31) / 32;
    if (conv || b.commit_unconverged) {          // without post-processing BP's last hard decision is the answer, converged or not
        int wd0 = ebits ? tid : 0;
        uint32_t hm = ebits ? (wd0 < w.nW32 ? ebits[wd0] : 0u) : hmask;
        for (;;) {
            if (!hm) {
                if (!ebits) break;
                wd0 += NT;
                if (wd0 >= w.nW32) break;
                hm = ebits[wd0];
                continue;
            }
            const int kk = __ffs(hm) - 1;
            hm &= hm - 1;
            int j = ebits ? 32 * wd0 + kk : tid + kk * NT;
            if (RECORDS) j = static_cast<int>(__ldg(&w.colrec[j].w) & 0xFFFFu);      // record -> original column
            if (b.ehat_out) atomicOr(&b.ehat_out[static_cast<size_t>(shot) * b.ehat_stride32 + (j >> 5)], 1u << (j & 31));
            if (j < w.ncommit) {
                for (int wd = 0; wd < w.KW; ++wd) {
                    const uint64_t lm = __ldg(w.lmask + static_cast<size_t>(j) * w.KW + wd);
                    if (static_cast<uint32_t>(lm)) atomicXor(&accs[2 * wd], static_cast<uint32_t>(lm));
                    if (static_cast<uint32_t>(lm >> 32)) atomicXor(&accs[2 * wd + 1], static_cast<uint32_t>(lm >> 32));
                }
                if (w.carry_rows) {
                    for (int q = __ldg(w.uptr + j); q < __ldg(w.uptr + j + 1); ++q) {
                        const uint32_t r = __ldg(w.uidx + q);
                        atomicXor(&car[r >> 5], 1u << (r & 31));
                    }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < w.KW; i += NT) {
            const uint64_t v = (static_cast<uint64_t>(accs[2 * i + 1]) << 32) | accs[2 * i];
            b.acc[static_cast<size_t>(shot) * w.KW + i] ^= v;
        }
        for (int i = tid; i < carryW; i += NT) b.carry[static_cast<size_t>(shot) * b.carry_stride32 + i] = car[i];
    } else {
        // post-carry syndrome for OSD; the posteriors are already in llr_buf
        for (int i = tid; i < w.rowsW32; i += NT) b.syn_buf[static_cast<size_t>(shot) * b.syn_stride32 + i] = syn[i];
        if (tid == 0) {
            const int slot = atomicAdd(b.fail_count, 1);
            b.fail_list[slot] = shot;
        }
        if (b.sel_cnt) {
            __threadfence_block();
            __syncthreads();                                  // the posteriors of this shot are complete
            select_for_osd<R, NT>(w, b, shot, tid, hist);
        }
    }
    if (tid == 0) {
        if (conv) atomicAdd(&b.stats[0], 1ull);
        atomicAdd(&b.stats[1], static_cast<unsigned long long>(it));
        if (b.iters_out) b.iters_out[shot] = it;
        if (b.conv_out) b.conv_out[shot] = conv ? 1 : 0;
    }
}

template <typename R> struct Compact;
template <> struct Compact<float> {
    static __device__ __forceinline__ float2 sum0(const WinDev& w, int i) { return __ldg(w.rsum0f + i); }
    static __device__ __forceinline__ const float* ptab(const WinDev& w) { return w.ptabf; }
    static __device__ __forceinline__ float signed_by(float m, uint32_t neg) { return __uint_as_float(__float_as_uint(m) | (neg << 31)); }
    static __device__ __forceinline__ float mag(float m) { return __uint_as_float(__float_as_uint(m) & 0x7FFFFFFFu); }
    // x with its sign flipped when (sign bit of s) xor neg
    static __device__ __forceinline__ float flip(float x, float s, bool neg) {
        return __uint_as_float(__float_as_uint(x) ^ ((__float_as_uint(s) & 0x80000000u) ^ (neg ? 0x80000000u : 0u)));
    }
};
template <> struct Compact<double> {
    static __device__ __forceinline__ double2 sum0(const WinDev& w, int i) { return __ldg(w.rsum0d + i); }
    static __device__ __forceinline__ const double* ptab(const WinDev& w) { return w.ptabd; }
    static __device__ __forceinline__ double signed_by(double m, uint32_t neg) {
        return __hiloint2double(__double2hiint(m) | static_cast<int>(neg << 31), __double2loint(m));
    }
    static __device__ __forceinline__ double mag(double m) { return __hiloint2double(__double2hiint(m) & 0x7FFFFFFF, __double2loint(m)); }
    static __device__ __forceinline__ double flip(double x, double s, bool neg) {
        const uint32_t f = (static_cast<uint32_t>(__double2hiint(s)) & 0x80000000u) ^ (neg ? 0x80000000u : 0u);
        return __hiloint2double(static_cast<int>(static_cast<uint32_t>(__double2hiint(x)) ^ f), __double2loint(x));
    }
};

// ---- product-sum (ldpc bp_method 'product_sum'): check->bit message = 2 atanh( prod_{others} tanh(v/2) ), sign from the syndrome.
// The message array holds tanh(v/2) instead of v (one tanh per edge and iteration, taken when the message is written).
// The row summary is (s * prod of the non-zero tanh(v/2), number of zero factors); "the others" is obtained by dividing the
// row product by the edge's own factor, which differs from ldpc's prefix/suffix products by rounding only (x is formed
// explicitly, as ldpc does, so that saturation -- x rounding to exactly 1, message +inf -- happens at the same places) (no bit parity with
// the CPU here anyway: libm and CUDA tanh/log differ in the last ulps).
template <typename R> struct Trans;
template <> struct Trans<float> {
    static __device__ __forceinline__ float th(float x) { return tanhf(x); }
    static __device__ __forceinline__ float lg(float x) { return logf(x); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
};
template <> struct Trans<double> {
    static __device__ __forceinline__ double th(double x) { return tanh(x); }
    static __device__ __forceinline__ double lg(double x) { return log(x); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
};

// warp minimum of non-negative reals through their bit patterns (monotone for x >= 0)
__device__ __forceinline__ float warp_min_mag(float x) {
    return __uint_as_float(__reduce_min_sync(0xFFFFFFFFu, __float_as_uint(x)));
}
__device__ __forceinline__ double warp_min_mag(double x) {
    const uint32_t hi = static_cast<uint32_t>(__double2hiint(x)), lo = static_cast<uint32_t>(__double2loint(x));
    const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
    const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, hi == mh ? lo : 0xFFFFFFFFu);
    return __hiloint2double(static_cast<int>(mh), static_cast<int>(ml));
}

template <typename R> __device__ __forceinline__ uint32_t sign_bit(R x);
template <> __device__ __forceinline__ uint32_t sign_bit<float>(float x) { return __float_as_uint(x) >> 31; }
template <> __device__ __forceinline__ uint32_t sign_bit<double>(double x) { return static_cast<uint32_t>(__double2hiint(x)) >> 31; }

template <typename R> __device__ __forceinline__ R shfl_real(R x, int src);
template <> __device__ __forceinline__ float shfl_real<float>(float x, int src) { return __shfl_sync(0xFFFFFFFFu, x, src); }
template <> __device__ __forceinline__ double shfl_real<double>(double x, int src) { return __shfl_sync(0xFFFFFFFFu, x, src); }


}  // namespace

}  // namespace qb
