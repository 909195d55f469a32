// quits_b200/csrc/frame.cu -- K1: Pauli-frame propagation of the syndrome-extraction circuit (sm_100a).
//
// Replaces what the reference obtains from stim's compile_detector_sampler().sample()
// (reference src/quits/simulation.py:22-27).  One warp owns one 64-shot bit-word: the X/Z frame of every
// qubit, a ring of the most recent measurement words and the detector/observable words live in shared
// memory as uint64; the lanes of the warp walk the targets of each tape op in parallel (tape ops are
// conflict-free slices, see qb_host.cpp build_tape) with a __syncwarp between ops.  No block-level barrier.
// Noise is counter based (Philox4x32-10 keyed by the run seed, counter = (site, word)), so a shot's faults do
// not depend on the launch geometry, the batch size or the GPU that ran it.
// Output is written shot-major (one packed row of detector bits per shot) after an in-place 64x64 bit
// transpose, which is the layout the decoder kernels read.
#include "qb_device.h"

namespace qb {

namespace {

struct Ph4 { uint32_t v[4]; };

__device__ __forceinline__ Ph4 philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    Ph4 o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}

__device__ __forceinline__ void sm_xor(uint64_t* p, uint64_t m) {
    if (m) atomicXor(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(m));
}
// the lane is the only one of the warp that touches *p during this tape op (targets of the op are distinct qubits)
__device__ __forceinline__ void sm_xor_owned(uint64_t* p, uint64_t m, bool owned) {
    if (!m) return;
    if (owned) *p ^= m;
    else atomicXor(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(m));
}

// levels 2 and 3 of the sampling scheme: how many of the 64 shots of this word are hit at this site, which
// ones, and with which Pauli.  m[0..3] = masks to XOR into x_a, z_a, x_b, z_b.
__device__ void site_faults(uint32_t k0, uint32_t k1, uint32_t s, uint32_t wl, uint32_t wh, const uint64_t* __restrict__ ctab,
                            int npauli, int fixed_code, uint64_t m[4]) {
    Ph4 r = philox4x32_10(k0, k1, s, wl, wh, 1u);
    const uint64_t u = (static_cast<uint64_t>(r.v[0]) << 32) | r.v[1];
    int n = 1;
    while (n < 64 && u >= __ldg(&ctab[n])) ++n;
    uint64_t used = 0;
    m[0] = m[1] = m[2] = m[3] = 0;
    for (int j = 0; j < n; ++j) {
        Ph4 q = philox4x32_10(k0, k1, s, wl, wh, 2u + static_cast<uint32_t>(j));
        const int code = fixed_code ? fixed_code : 1 + static_cast<int>(__umulhi(q.v[0], static_cast<uint32_t>(npauli)));
        int pos = -1, last = 0;
        for (int i = 0; i < 15; ++i) {
            const uint32_t word = i < 5 ? q.v[1] : (i < 10 ? q.v[2] : q.v[3]);
            const int cand = static_cast<int>((word >> (6 * (i % 5))) & 63u);
            last = cand;
            if (!((used >> cand) & 1ull)) { pos = cand; break; }
        }
        if (pos < 0) {
            pos = last;
            while ((used >> pos) & 1ull) pos = (pos + 1) & 63;
        }
        const uint64_t bit = 1ull << pos;
        used |= bit;
        if (code & 1) m[0] |= bit;
        if (code & 2) m[1] |= bit;
        if (code & 4) m[2] |= bit;
        if (code & 8) m[3] |= bit;
    }
}

__global__ void __launch_bounds__(128) frame_kernel(const FrameArgs a, const int warps_per_block, const int words_per_warp_smem) {
    extern __shared__ uint64_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t widx = static_cast<uint64_t>(blockIdx.x) * warps_per_block + warp;
    if (widx >= a.n_words) return;                       // whole warp leaves; there is no block barrier below
    const uint64_t w = a.word0 + widx;
    const uint32_t wl = static_cast<uint32_t>(w), wh = static_cast<uint32_t>(w >> 32);
    const uint32_t k0 = static_cast<uint32_t>(a.seed), k1 = static_cast<uint32_t>(a.seed >> 32);
    const int nq = a.n_qubits, DW = a.DW, KW = a.KW, rmask = a.ring - 1;
    uint64_t* x = smem + static_cast<size_t>(warp) * words_per_warp_smem;
    uint64_t* z = x + nq;
    uint64_t* ring = z + nq;
    uint64_t* obs = ring + a.ring;
    // detector words go straight to a global scratch row of this 64-shot word (each is written exactly once, 32 lanes x 8 bytes
    // coalesced) instead of 7.7 KB of shared memory per warp: shared memory per warp is what bounds this kernel's occupancy
    uint64_t* det = a.det_words + widx * static_cast<uint64_t>(DW) * 64ull;
    for (int i = lane; i < 2 * nq; i += 32) x[i] = 0;
    for (int i = lane; i < KW * 64; i += 32) obs[i] = 0;
    for (int i = a.n_det + lane; i < DW * 64; i += 32) det[i] = 0;       // padding detectors of the last word
    __syncwarp();

    // the 32-byte op headers are read one op ahead (the tape is the same for every warp: L1 / L2 resident, but its latency would
    // otherwise sit between every two ops of the dependent chain); the tape is padded by one no-op
    int4 n0 = __ldg(reinterpret_cast<const int4*>(a.ops));
    int4 n1 = __ldg(reinterpret_cast<const int4*>(a.ops) + 1);
    for (int i = 0; i < a.n_ops; ++i) {
        const int4 h0 = n0, h1 = n1;
        n0 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1));
        n1 = __ldg(reinterpret_cast<const int4*>(a.ops + i + 1) + 1);
        const int kind = h0.x, n = h0.y;
        const uint32_t t0 = static_cast<uint32_t>(h0.z), aux = static_cast<uint32_t>(h0.w);
        const uint32_t thr = static_cast<uint32_t>(h1.x);
        const uint32_t* __restrict__ tg = a.targets + t0;
        switch (kind) {
        case OP_R: case OP_RX:
            for (int j = lane; j < n; j += 32) { const uint32_t q = __ldg(tg + j); x[q] = 0; z[q] = 0; }
            break;
        case OP_H:
            for (int j = lane; j < n; j += 32) { const uint32_t q = __ldg(tg + j); const uint64_t t = x[q]; x[q] = z[q]; z[q] = t; }
            break;
        case OP_CX:
            for (int j = lane; j < n; j += 32) {
                const uint2 ct = __ldg(reinterpret_cast<const uint2*>(a.targets) + (t0 >> 1) + j);
                x[ct.y] ^= x[ct.x];
                z[ct.x] ^= z[ct.y];
            }
            break;
        case OP_M:
            for (int j = lane; j < n; j += 32) ring[(aux + j) & rmask] = x[__ldg(tg + j)];
            break;
        case OP_MX:
            for (int j = lane; j < n; j += 32) ring[(aux + j) & rmask] = z[__ldg(tg + j)];
            break;
        case OP_MR:
            for (int j = lane; j < n; j += 32) { const uint32_t q = __ldg(tg + j); ring[(aux + j) & rmask] = x[q]; x[q] = 0; z[q] = 0; }
            break;
        case OP_DET:
            for (int j = lane; j < n; j += 32) {
                const uint32_t b = __ldg(a.detptr + t0 + j), e = __ldg(a.detptr + t0 + j + 1);
                uint64_t acc = 0;
                for (uint32_t q = b; q < e; ++q) acc ^= ring[__ldg(a.detidx + q) & rmask];
                __stcg(det + aux + j, acc);
            }
            break;
        case OP_OBS: {
            uint64_t acc = 0;
            for (int j = lane; j < n; j += 32) acc ^= ring[__ldg(tg + j) & rmask];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc ^= __shfl_xor_sync(0xFFFFFFFFu, acc, o);
            if (lane == 0) obs[aux] ^= acc;
        } break;
        default: {                                      // noise channels
            const bool two = kind == OP_DEP2;
            if (a.inject) {
                const int fb = __ldg(a.inj_start + i), fe = __ldg(a.inj_start + i + 1);
                for (int f = fb + lane; f < fe; f += 32) {
                    const int64_t sh = __ldg(a.inj_shot + f);
                    if (static_cast<uint64_t>(sh >> 6) != widx) continue;
                    const uint64_t bit = 1ull << (sh & 63);
                    const int code = __ldg(a.inj_code + f), j = __ldg(a.inj_tgt + f);
                    const uint32_t qa = __ldg(tg + (two ? 2 * j : j));
                    if (code & 1) sm_xor(&x[qa], bit);
                    if (code & 2) sm_xor(&z[qa], bit);
                    if (two) {
                        const uint32_t qb = __ldg(tg + 2 * j + 1);
                        if (code & 4) sm_xor(&x[qb], bit);
                        if (code & 8) sm_xor(&z[qb], bit);
                    }
                }
                break;
            }
            if (thr == 0) break;
            const int npauli = kind == OP_DEP1 ? 3 : (two ? 15 : 0);
            const int fixed = kind == OP_XERR ? 1 : (kind == OP_ZERR ? 2 : 0);
            const uint64_t* ctab = a.ctab + static_cast<size_t>(h1.y) * 64;
            const int groups = (n + 3) >> 2;
            const bool owned = h1.w != 0;
            for (int g = lane; g < groups; g += 32) {
                const uint32_t s0 = aux + 4u * static_cast<uint32_t>(g);
                const Ph4 r = philox4x32_10(k0, k1, s0 >> 2, wl, wh, 0u);
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const int j = 4 * g + l;
                    if (j >= n) break;
                    if (!(r.v[l] < thr || thr == 0xFFFFFFFFu)) continue;
                    uint64_t m[4];
                    site_faults(k0, k1, s0 + static_cast<uint32_t>(l), wl, wh, ctab, npauli, fixed, m);
                    const uint32_t qa = __ldg(tg + (two ? 2 * j : j));
                    sm_xor_owned(&x[qa], m[0], owned);
                    sm_xor_owned(&z[qa], m[1], owned);
                    if (two) {
                        const uint32_t qb = __ldg(tg + 2 * j + 1);
                        sm_xor_owned(&x[qb], m[2], owned);
                        sm_xor_owned(&z[qb], m[3], owned);
                    }
                }
            }
        } break;
        }
        __syncwarp();
    }

    // 64x64 bit transposes: word d of block c (bit b = shot b) -> word c of shot b (bit d = detector c*64 + d).  The detector words
    // come back from the scratch row (L2: this warp wrote them; every lane reads the same word), the shot rows are written in place.
    __syncwarp();
    uint64_t* drow = a.det_rows + widx * 64ull * DW;
    for (int c = 0; c < DW; ++c) {
        const uint64_t* blk = det + c * 64;
        uint64_t o0 = 0, o1 = 0;
#pragma unroll 16
        for (int d = 0; d < 64; ++d) {
            const uint64_t v = __ldcg(blk + d);
            o0 |= ((v >> lane) & 1ull) << d;
            o1 |= ((v >> (lane + 32)) & 1ull) << d;
        }
        drow[static_cast<size_t>(lane) * DW + c] = o0;
        drow[static_cast<size_t>(lane + 32) * DW + c] = o1;
    }
    uint64_t* orow = a.obs_rows + widx * 64ull * KW;
    for (int c = 0; c < KW; ++c) {
        const uint64_t* blk = obs + c * 64;
        uint64_t o0 = 0, o1 = 0;
#pragma unroll 8
        for (int d = 0; d < 64; ++d) {
            const uint64_t v = blk[d];
            o0 |= ((v >> lane) & 1ull) << d;
            o1 |= ((v >> (lane + 32)) & 1ull) << d;
        }
        orow[static_cast<size_t>(lane) * KW + c] = o0;
        orow[static_cast<size_t>(lane + 32) * KW + c] = o1;
    }
}

__global__ void unpack_bits_kernel(const uint64_t* __restrict__ rows, int wpr, int nbits, uint64_t n_rows, uint8_t* __restrict__ out) {
    const uint64_t total = n_rows * static_cast<uint64_t>(nbits);
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t r = i / nbits;
        const int b = static_cast<int>(i - r * nbits);
        out[i] = static_cast<uint8_t>((rows[r * wpr + (b >> 6)] >> (b & 63)) & 1ull);
    }
}

__global__ void pack_bits_kernel(const uint8_t* __restrict__ in, int nbits, uint64_t n_rows, uint64_t* __restrict__ rows, int wpr) {
    // one warp per (row, 64-bit word): two ballots
    const int lane = threadIdx.x & 31;
    const uint64_t nwarps = (static_cast<uint64_t>(gridDim.x) * blockDim.x) >> 5;
    const uint64_t total = n_rows * static_cast<uint64_t>(wpr);
    for (uint64_t wi = (blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x) >> 5; wi < total; wi += nwarps) {
        const uint64_t r = wi / wpr;
        const int c = static_cast<int>(wi - r * wpr);
        const int b0 = c * 64 + lane, b1 = b0 + 32;
        const uint8_t v0 = b0 < nbits ? in[r * nbits + b0] : 0, v1 = b1 < nbits ? in[r * nbits + b1] : 0;
        const uint32_t lo = __ballot_sync(0xFFFFFFFFu, v0 & 1), hi = __ballot_sync(0xFFFFFFFFu, v1 & 1);
        if (lane == 0) rows[wi] = (static_cast<uint64_t>(hi) << 32) | lo;
    }
}

__global__ void expand_pred_kernel(const uint64_t* __restrict__ acc, int KW, int K, uint64_t n, int64_t* __restrict__ pred) {
    const uint64_t total = n * static_cast<uint64_t>(K);
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t s = i / K;
        const int k = static_cast<int>(i - s * K);
        pred[i] = static_cast<int64_t>((acc[s * KW + (k >> 6)] >> (k & 63)) & 1ull);
    }
}

__global__ void count_kernel(const uint64_t* __restrict__ acc, const uint64_t* __restrict__ obs, int KW, int K, uint64_t n,
                             unsigned long long* counts) {
    extern __shared__ unsigned int cnt[];             // [1 + K]
    for (int i = threadIdx.x; i <= K; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (uint64_t s = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; s < n; s += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        bool any = false;
        for (int wd = 0; wd < KW; ++wd) {
            uint64_t diff = acc[s * KW + wd] ^ obs[s * KW + wd];
            if (wd == KW - 1 && (K & 63)) diff &= (1ull << (K & 63)) - 1;
            any = any || diff != 0;
            while (diff) {
                const int b = __ffsll(static_cast<long long>(diff)) - 1;
                diff &= diff - 1;
                atomicAdd(&cnt[1 + wd * 64 + b], 1u);
            }
        }
        if (any) atomicAdd(&cnt[0], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= K; i += blockDim.x)
        if (cnt[i]) atomicAdd(&counts[i], static_cast<unsigned long long>(cnt[i]));
}

}  // namespace

size_t frame_smem_per_warp(const FrameArgs& a) {
    return (static_cast<size_t>(2) * a.n_qubits + a.ring + static_cast<size_t>(a.KW) * 64) * sizeof(uint64_t);
}
size_t frame_scratch_bytes(const FrameArgs& a) { return static_cast<size_t>(a.n_words) * a.DW * 64 * sizeof(uint64_t) + 64; }

cudaError_t launch_frame(const FrameArgs& a, cudaStream_t st) {
    if (a.n_words == 0) return cudaSuccess;
    const size_t per_warp = frame_smem_per_warp(a);
    const size_t budget = 200 * 1024;
    if (per_warp > budget) return cudaErrorInvalidValue;
    if (!a.det_words) return cudaErrorInvalidValue;
    int wpb = 4;
    while (wpb > 1 && per_warp * wpb > budget / 6) --wpb;          // keep >= 6 CTAs per SM when possible
    const size_t smem = per_warp * wpb;
    static size_t configured_dev[kMaxDevices] = {};
    size_t& configured = configured_dev[device_slot()];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(frame_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const uint64_t blocks = (a.n_words + wpb - 1) / wpb;
    frame_kernel<<<static_cast<unsigned>(blocks), wpb * 32, smem, st>>>(a, wpb, static_cast<int>(per_warp / sizeof(uint64_t)));
    return cudaGetLastError();
}

static unsigned grid_for(uint64_t work, int per_block) {
    uint64_t b = (work + per_block - 1) / per_block;
    const uint64_t cap = 148ull * 16;
    return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

cudaError_t launch_unpack_bits(const uint64_t* rows, int wpr, int nbits, uint64_t n_rows, uint8_t* out, cudaStream_t st) {
    if (n_rows == 0 || nbits == 0) return cudaSuccess;
    unpack_bits_kernel<<<grid_for(n_rows * nbits, 256 * 8), 256, 0, st>>>(rows, wpr, nbits, n_rows, out);
    return cudaGetLastError();
}

cudaError_t launch_pack_bits(const uint8_t* in, int nbits, uint64_t n_rows, uint64_t* rows, int wpr, cudaStream_t st) {
    if (n_rows == 0 || wpr == 0) return cudaSuccess;
    pack_bits_kernel<<<grid_for(n_rows * wpr, 8 * 4), 256, 0, st>>>(in, nbits, n_rows, rows, wpr);
    return cudaGetLastError();
}

cudaError_t launch_expand_pred(const uint64_t* acc, int KW, int K, uint64_t n, int64_t* pred, cudaStream_t st) {
    if (n == 0 || K == 0) return cudaSuccess;
    expand_pred_kernel<<<grid_for(n * K, 256 * 4), 256, 0, st>>>(acc, KW, K, n, pred);
    return cudaGetLastError();
}

cudaError_t launch_count(const uint64_t* acc, const uint64_t* obs_rows, int KW, int K, uint64_t n, unsigned long long* counts,
                         cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    count_kernel<<<grid_for(n, 256 * 4), 256, (1 + KW * 64) * sizeof(unsigned int), st>>>(acc, obs_rows, KW, K, n, counts);
    return cudaGetLastError();
}

}  // namespace qb
