// sa_mlp_train.cu -- TRAIN-MODE SetAbstraction shared MLP on tcgen05 tensor cores (sm_100a): forward with
// batch-statistic BatchNorm + max-pool, and the full backward (dgrad, wgrad, BatchNorm and ReLU / max-pool gradients).
//
// Replaces, for module.train() / autograd, what models/pointnet2_encoder.py:106-112 runs through cuDNN / cuBLAS:
//     for conv, bn in zip(mlp_convs, mlp_bns): x = relu(bn(conv(x)))     # nn.BatchNorm2d, batch statistics (:74,110)
//     x = max(x, dim=K)
// BatchNorm in training mode needs the per-channel mean / variance over ALL B*S*K rows between the GEMM and the ReLU,
// so a layer cannot be fused with the next one inside a tile; every layer is a pass over the rows:
//   forward, layer l:   Z_l = X_{l-1} W_l^T + bias_l     (tcgen05 GEMM; X_{l-1} = relu(bn(Z_{l-1})) applied on the fly in
//                                                          the operand build, or the grouping gather for l = 0)  -> bf16
//                       column sums / sums of squares of Z_l (fp64)  ->  mean, 1/sqrt(var + eps), running-stat update
//   forward, pooling:   out[g, c] = max_k relu(a_c Z_2[g, k, c] + b_c), argmax kept for the backward
//   backward, layer l:  per-channel sums  s1 = sum dY_l, s2 = sum dY_l * xhat_l   (sparse for the pooled layer)
//                       dZ_l = gamma / sigma * (dY_l - s1 / n - xhat_l * s2 / n)   (never materialised: rebuilt in the
//                                                          operand build of the two GEMMs below)
//                       dgrad:  dY_{l-1} = (dZ_l W_l) * [X_{l-1} > 0]             (tcgen05 GEMM, rows x Cin)
//                       wgrad:  dW_l = dZ_l^T X_{l-1}                            (tcgen05 GEMM, K = rows, both operands
//                                                          written TRANSPOSED into shared memory, TMEM accumulates across tiles)
// The pre-BatchNorm activations Z_l between passes live in HBM as FP32 [rows, C] (the "saved" blob, also what autograd
// keeps for the backward).  fp32, not bf16: BatchNorm subtracts the batch mean, so rounding Z to 8 mantissa bits BEFORE the
// subtraction costs |mean| / sigma times the bf16 step in the normalised value (measured: 5-20 % feature error on the
// reference's encoder when Z was stored as bf16); the operands the tensor cores read are rounded to bf16 AFTER
// normalisation + ReLU, where the values are O(1).  Gradients dY between passes are fp32 as well.  The path is HBM-bound:
// ~4 * C bytes written and ~8 * C read per row and layer.
// Precision modes (argument `precision` of the entry points):
//   1 = bf16 operands, one tcgen05.mma per K step (the autocast / config-4 mode): forward within 1e-2 relative L2 per stage;
//   0 = SPLIT operands ("bf16x3"): every fp32 operand x is fed as hi = bf16(x), lo = bf16(x - hi) and a K step issues
//       hi*hi + lo*hi + hi*lo into the same fp32 accumulator (16 mantissa bits per operand, the dropped lo*lo term is
//       2^-18) -- fp32-faithful GEMMs on the bf16 tensor pipe, so that the train-mode result tracks the reference's fp32
//       autograd closely enough that max-pool / ReLU selections do not flip (tests: rtol 2e-3).  Tensor work triples,
//       which is immaterial here: the passes are HBM-bound.  A operands wider than 256 channels are processed in K
//       panels of 256 so that both copies fit shared memory.
// fp32 accumulate, activations and inter-layer gradients; fp64 batch statistics in both modes.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pcst {

typedef __nv_bfloat16 bf16;

#define PCST_CUDA_NAMED(call, what)                         \
    do {                                                    \
        int _st = pcst::check_cuda((call), what);           \
        if (_st != PCST_OK) return _st;                     \
    } while (0)

constexpr int kTrMaxC = 512;              // widest layer (input or output channels, padded)
constexpr int kTrStageBytes = 16 * 1024;  // weight ring stage of the row GEMM
constexpr int kTrStages = 3;

// ---- small helpers ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float (&v)[8]) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void load_f32x8(const float* p, float (&v)[8]) {
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(p + 4));
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// hi = bf16(v), lo = bf16(v - hi): the two bf16 terms of the split ("bf16x3") operand
__device__ __forceinline__ void split_bf16x8(const float (&v)[8], uint4& hi, uint4& lo) {
    hi = pack_bf16x8(v);
    float h[8], l[8];
    unpack_bf16x8(hi, h);
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = v[i] - h[i];
    lo = pack_bf16x8(l);
}

// What a row's operand is built from.  One struct serves the row GEMM (A operand) and the wgrad kernel (both operands).
struct RowSrc {
    int kind;  // 0 = grouping gather (layer-0 input), 1 = relu(a * Z + b) (input of layers 1, 2),
               // 2 = dZ from a dense dY, 3 = dZ of the pooled layer (dY is nonzero only at the argmax rows)
    // kind 0
    const float* xyz;
    const float* feats;
    const float* new_xyz;
    const int64_t* idx;
    int N, S, K, D;
    // kinds 1-3: Z [rows, C] fp32; per-channel fp32 vectors (global): kind 1: vec = a | b; kinds 2, 3: vec = g | m1 | m2 | mean | invstd
    const float* Z;
    const float* vec;
    int C;
    // kind 2
    const float* dY;
    // kind 3
    const int* argmax;  // [G, C]; -1 = the pooled value was clipped by the ReLU (no gradient)
    const float* dout;  // [G, C]
};

// Operand order of the gathered layer-0 input: the D feature channels first (16-byte aligned groups), then the three
// relative coordinates, then zero padding -- the same order sa_mlp_tc.cu uses (reference order is xyz first, :99).
template <typename Emit>
__device__ __forceinline__ void build_gather(const RowSrc& s, int r, bool valid, int k_begin, int k_end, Emit&& emit) {
    int j = 0, bs = 0;
    if (valid) {
        bs = r / s.K;
        if (s.idx) {
            const int64_t jj = s.idx[r];
            j = jj < 0 ? 0 : (jj >= s.N ? s.N - 1 : (int)jj);
        } else {
            j = r % s.K;
        }
    }
    const int b = bs / s.S;
    float rel[3] = {0.f, 0.f, 0.f};
    if (valid) {
        const float* p = s.xyz + ((size_t)b * s.N + j) * 3;
        rel[0] = p[0]; rel[1] = p[1]; rel[2] = p[2];
        if (s.new_xyz) {
            const float* c = s.new_xyz + (size_t)bs * 3;
            rel[0] = __fsub_rn(rel[0], c[0]); rel[1] = __fsub_rn(rel[1], c[1]); rel[2] = __fsub_rn(rel[2], c[2]);
        }
    }
    const int D = s.D;
    const float* frow = D > 0 ? s.feats + ((size_t)b * s.N + j) * D : nullptr;
    const bool vec = D > 0 && (D % 4) == 0 && ((reinterpret_cast<uintptr_t>(s.feats) & 15) == 0);
    for (int kc = k_begin / 8; kc < k_end / 8; ++kc) {
        float v[8];
        const int k0 = kc * 8;
        if (valid && vec && k0 + 8 <= D) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(frow + k0));
            const float4 q1 = __ldg(reinterpret_cast<const float4*>(frow + k0 + 4));
            v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = k0 + i;
                float x = 0.f;
                if (valid) {
                    if (k < D) x = __ldg(frow + k);
                    else if (k < D + 3) x = rel[k - D];
                }
                v[i] = x;
            }
        }
        emit(kc - k_begin / 8, v);
    }
}

// channels [c0, c1) (multiples of 8) of relu(a * Z + b)
template <typename Emit>
__device__ __forceinline__ void build_bnrelu(const RowSrc& s, const float* vec_sm, int r, bool valid, int c0, int c1, Emit&& emit) {
    const float* a = vec_sm;
    const float* b = vec_sm + s.C;
    const float* zrow = s.Z + (size_t)r * s.C;
    for (int c = c0; c < c1; c += 8) {
        float v[8];
        if (valid) {
            float z[8];
            load_f32x8(zrow + c, z);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(__fmaf_rn(a[c + i], z[i], b[c + i]), 0.f);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        emit((c - c0) / 8, v);
    }
}

// channels [c0, c1) of dZ = g * (dY - m1 - xhat * m2), xhat = (Z - mean) * invstd
template <typename Emit>
__device__ __forceinline__ void build_dz(const RowSrc& s, const float* vec_sm, int r, bool valid, int c0, int c1, Emit&& emit) {
    const int C = s.C;
    const float *g = vec_sm, *m1 = vec_sm + C, *m2 = vec_sm + 2 * C, *mean = vec_sm + 3 * C, *istd = vec_sm + 4 * C;
    const float* zrow = s.Z + (size_t)r * C;
    int grp = 0, kk = 0;
    if (valid && s.kind == 3) {
        grp = r / s.K;
        kk = r % s.K;
    }
    for (int c = c0; c < c1; c += 8) {
        float v[8];
        if (valid) {
            float z[8], dy[8];
            load_f32x8(zrow + c, z);
            if (s.kind == 2) {
                load_f32x8(s.dY + (size_t)r * C + c, dy);
            } else {
                const int4 a0 = __ldg(reinterpret_cast<const int4*>(s.argmax + (size_t)grp * C + c));
                const int4 a1 = __ldg(reinterpret_cast<const int4*>(s.argmax + (size_t)grp * C + c + 4));
                const float4 d0 = __ldg(reinterpret_cast<const float4*>(s.dout + (size_t)grp * C + c));
                const float4 d1 = __ldg(reinterpret_cast<const float4*>(s.dout + (size_t)grp * C + c + 4));
                dy[0] = a0.x == kk ? d0.x : 0.f; dy[1] = a0.y == kk ? d0.y : 0.f;
                dy[2] = a0.z == kk ? d0.z : 0.f; dy[3] = a0.w == kk ? d0.w : 0.f;
                dy[4] = a1.x == kk ? d1.x : 0.f; dy[5] = a1.y == kk ? d1.y : 0.f;
                dy[6] = a1.z == kk ? d1.z : 0.f; dy[7] = a1.w == kk ? d1.w : 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float xh = (z[i] - mean[c + i]) * istd[c + i];
                v[i] = g[c + i] * (dy[i] - m1[c + i] - xh * m2[c + i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
        }
        emit((c - c0) / 8, v);
    }
}

// ---- row GEMM: ACC[128 rows, n] = A[128, kp] * Wp[n, kp]^T, one tile per CTA ------------------------------------------------
// K is walked in panels (256 channels in split mode, else the whole K): build the panel's A operand, issue its MMAs for
// every 256-column chunk of N, next panel (the accumulators of all chunks stay in TMEM).
struct TgArgs {
    RowSrc src;          // the A operand
    int rows, kp, n;     // kp = padded K (multiple of 16), n = output columns (multiple of 16, <= 512)
    int split;           // 1 = split operands (hi + lo copies of A and of the weights, three MMAs per K step)
    const unsigned char* wblob;  // packed B operand: per 256-column chunk [kp/8][nc][8] bf16 (hi), then the same for lo when split
    int epi;             // 0: Z = acc + bias -> fp32 [rows, n];  1: dY_prev = acc * [a * Zp + b > 0] -> fp32 [rows, n];
                         // 2: grad of the gathered input -> fp32 [rows, 3 + D] in the reference's channel order
    const float* bias;   // epi 0: [n]
    float* out_z;        // epi 0, 1
    const float* zprev;  // epi 1: Z_{l-1} [rows, n]
    const float* ab;     // epi 1: a | b of layer l-1 [2][n]
    float* out_f32;      // epi 2
    int D;               // epi 2
    uint32_t off_a, a_bytes, off_ring, off_vec, off_ab, tmem_cols;
};

__device__ __forceinline__ int tg_chunk_rows(int nc, int klen, int split) {
    int ck = kTrStageBytes / (nc * 2 * (split ? 2 : 1)) / 16 * 16;
    if (ck < 16) ck = 16;
    return ck < klen ? ck : klen;
}

__global__ void __launch_bounds__(kTcThreads)
train_gemm_kernel(const __grid_constant__ TgArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kTrStages];
    __shared__ __align__(8) uint64_t empty_bar[kTrStages];
    __shared__ __align__(8) uint64_t a_bar;    // a K panel of the A operand is in shared memory
    __shared__ __align__(8) uint64_t mma_bar;  // the MMAs of a K panel have completed
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (uniform role branches)
    const int row0 = blockIdx.x * kTcM;
    if (tid == 0) {
        for (int s = 0; s < kTrStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&a_bar, kTcEpiThreads);
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_sh, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    const int nchunks = (a.n + 255) / 256;
    const int kpan = a.split ? 256 : kTrMaxC;
    const int npan = (a.kp + kpan - 1) / kpan;
    const int copies = a.split ? 2 : 1;

    if (warp == 4) {
        // ---- weight producer: panel by panel, chunk by chunk, K rows per stage = as many as fit ----
        if (lane == 0) {
            uint32_t it = 0;
            for (int pan = 0; pan < npan; ++pan) {
                const int kbeg = pan * kpan, klen = min(kpan, a.kp - kbeg);
                size_t woff = 0;
                for (int ch = 0; ch < nchunks; ++ch) {
                    const int nc = min(256, a.n - ch * 256);
                    const int ck = tg_chunk_rows(nc, klen, a.split);
                    for (int k0 = 0; k0 < klen; k0 += ck, ++it) {
                        const uint32_t stage = it % kTrStages;
                        if (it >= kTrStages) mbar_wait(&empty_bar[stage], ((it / kTrStages) - 1u) & 1u);
                        const int rowsk = min(ck, klen - k0);
                        const uint32_t bytes = (uint32_t)rowsk * nc * 2u;
                        unsigned char* dst = smem + a.off_ring + stage * kTrStageBytes;
                        const unsigned char* src = a.wblob + woff + (size_t)(kbeg + k0) * nc * 2u;
                        mbar_arrive_expect_tx(&full_bar[stage], bytes * copies);
                        tma_load_1d(dst, src, bytes, &full_bar[stage]);
                        if (a.split) tma_load_1d(dst + bytes, src + (size_t)a.kp * nc * 2u, bytes, &full_bar[stage]);
                    }
                    woff += (size_t)a.kp * nc * 2u * copies;
                }
            }
        }
    } else if (warp == 5) {
        // ---- MMA issuer: the whole warp converged, one elected lane issues (uniform datapath; see sa_mlp_tc.cu) ----
        const bool leader = elect_one_sync();
        const uint32_t smem_base = tc_opaque_u32(smem_u32(smem));
        const uint32_t full0 = tc_opaque_u32(smem_u32(&full_bar[0])), empty0 = tc_opaque_u32(smem_u32(&empty_bar[0]));
        const uint32_t a_bar_a = tc_opaque_u32(smem_u32(&a_bar)), mma_bar_a = tc_opaque_u32(smem_u32(&mma_bar));
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);     // SBO = 128 B; descriptor version 1 (bit 46)
        const uint32_t a_step = (2u * kTcM * 16u) >> 4;        // one K = 16 slice of A
        const uint32_t a_base = (((smem_base + a.off_a) >> 4) & 0x3FFFu) | (((kTcM * 16u) >> 4) << 16);
        const uint32_t a_lo_copy = a.a_bytes >> 4;             // the lo copy of the operand (split mode)
        const uint32_t ring_lo = (smem_base + a.off_ring) >> 4;
        const bool split = a.split != 0;
        uint32_t stage = 0, round = 0;
        for (int pan = 0; pan < npan; ++pan) {
            const int klen = min(kpan, a.kp - pan * kpan);
            mbar_wait_addr(a_bar_a, pan & 1);
            tc_fence_after();
            for (int ch = 0; ch < nchunks; ++ch) {
                const uint32_t nc = (uint32_t)min(256, a.n - ch * 256);
                const uint32_t ck = (uint32_t)tg_chunk_rows((int)nc, klen, a.split);
                const uint32_t idesc = umma_idesc_bf16(kTcM, (int)nc);
                const uint32_t w_lbo = nc << 16, w_step = nc * 2u;
                const uint32_t d_addr = tmem_base + (uint32_t)ch * 256u;
                uint32_t a_lo = a_base;
                for (uint32_t k0 = 0; k0 < (uint32_t)klen; k0 += ck) {
                    mbar_wait_addr(full0 + stage * 8u, round & 1u);
                    tc_fence_after();
                    const uint32_t rowsk = min(ck, (uint32_t)klen - k0), nk = rowsk >> 4;
                    if (leader) {
                        const uint32_t w_lo = ((ring_lo + stage * (kTrStageBytes >> 4)) & 0x3FFFu) | w_lbo;
                        const uint32_t w_lo_copy = (rowsk * nc * 2u) >> 4;  // the lo copy follows the hi copy inside the stage
                        const bool first = pan == 0 && k0 == 0;
                        for (uint32_t kk = 0; kk < nk; ++kk) {
                            const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_lo + kk * a_step);
                            const uint64_t bd = ((uint64_t)desc_hi << 32) | (w_lo + kk * w_step);
                            umma_bf16(d_addr, ad, bd, idesc, !(first && kk == 0));
                            if (split) {
                                umma_bf16(d_addr, ad + a_lo_copy, bd, idesc, true);
                                umma_bf16(d_addr, ad, bd + w_lo_copy, idesc, true);
                            }
                        }
                        umma_commit_addr(empty0 + stage * 8u);
                    }
                    __syncwarp();
                    a_lo += nk * a_step;
                    if (++stage == kTrStages) { stage = 0; ++round; }
                }
            }
            if (leader) umma_commit_addr(mma_bar_a);
            __syncwarp();
        }
    } else {
        // ---- operand build + epilogue: thread = row = TMEM lane ----
        const int m = tid, r = row0 + m;
        const bool valid = r < a.rows;
        float* vec_sm = reinterpret_cast<float*>(smem + a.off_vec);
        float* ab_sm = reinterpret_cast<float*>(smem + a.off_ab);
        if (a.src.kind >= 1) {
            const int nv = (a.src.kind == 1 ? 2 : 5) * a.src.C;
            for (int i = tid; i < nv; i += kTcEpiThreads) vec_sm[i] = a.src.vec[i];
        }
        if (a.epi == 1)
            for (int i = tid; i < 2 * a.n; i += kTcEpiThreads) ab_sm[i] = a.ab[i];
        epi_bar_sync();
        unsigned char* A = smem + a.off_a;
        auto emit = [&](int kc, const float (&v)[8]) {
            uint4* dst = reinterpret_cast<uint4*>(A + ((size_t)kc * kTcM + m) * 16);
            if (a.split) {
                uint4 hi, lo;
                split_bf16x8(v, hi, lo);
                *dst = hi;
                *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dst) + a.a_bytes) = lo;
            } else {
                *dst = pack_bf16x8(v);
            }
        };
        for (int pan = 0; pan < npan; ++pan) {
            const int kbeg = pan * kpan, kend = min(a.kp, kbeg + kpan);
            if (pan > 0) mbar_wait(&mma_bar, (pan - 1) & 1);  // the previous panel's MMAs have read the operand
            if (a.src.kind == 0) build_gather(a.src, r, valid, kbeg, kend, emit);
            else if (a.src.kind == 1) build_bnrelu(a.src, vec_sm, r, valid, kbeg, kend, emit);
            else build_dz(a.src, vec_sm, r, valid, kbeg, kend, emit);
            fence_proxy_async();
            mbar_arrive(&a_bar);
        }
        mbar_wait(&mma_bar, (npan - 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < a.n; c0 += 16) {
            uint32_t rr[16];
            tmem_ld16_issue(taddr + (uint32_t)c0, rr);
            tmem_ld_wait(rr);
            if (!valid) continue;
            if (a.epi == 0) {
                float4* dst = reinterpret_cast<float4*>(a.out_z + (size_t)r * a.n + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bq = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + 4 * i));
                    dst[i] = make_float4(__uint_as_float(rr[4 * i]) + bq.x, __uint_as_float(rr[4 * i + 1]) + bq.y,
                                         __uint_as_float(rr[4 * i + 2]) + bq.z, __uint_as_float(rr[4 * i + 3]) + bq.w);
                }
            } else if (a.epi == 1) {
                const float* zp = a.zprev + (size_t)r * a.n + c0;
                float4* dst = reinterpret_cast<float4*>(a.out_z + (size_t)r * a.n + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 z4 = __ldg(reinterpret_cast<const float4*>(zp + 4 * i));
                    const int c = c0 + 4 * i;
                    float4 o;
                    o.x = __fmaf_rn(ab_sm[c], z4.x, ab_sm[a.n + c]) > 0.f ? __uint_as_float(rr[4 * i]) : 0.f;
                    o.y = __fmaf_rn(ab_sm[c + 1], z4.y, ab_sm[a.n + c + 1]) > 0.f ? __uint_as_float(rr[4 * i + 1]) : 0.f;
                    o.z = __fmaf_rn(ab_sm[c + 2], z4.z, ab_sm[a.n + c + 2]) > 0.f ? __uint_as_float(rr[4 * i + 2]) : 0.f;
                    o.w = __fmaf_rn(ab_sm[c + 3], z4.w, ab_sm[a.n + c + 3]) > 0.f ? __uint_as_float(rr[4 * i + 3]) : 0.f;
                    dst[i] = o;
                }
            } else {
                // operand column j: feature j (j < D) -> reference channel 3 + j; D <= j < D + 3 -> coordinate j - D
                float* orow = a.out_f32 + (size_t)r * (3 + a.D);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = c0 + i;
                    if (j < a.D) orow[3 + j] = __uint_as_float(rr[i]);
                    else if (j < a.D + 3) orow[j - a.D] = __uint_as_float(rr[i]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ---- wgrad: dW[mb*128 .. +128, 0 .. kp) += sum over this CTA's row tiles of dZ^T X ---------------------------------------
// Row tiles of `tr` rows (128; 64 in split mode so that the hi and lo copies of both operands fit shared memory).
struct WgArgs {
    RowSrc dz;       // kind 2 or 3: the layer's dZ (A operand = its channels [mb*128, mb*128+128) transposed)
    RowSrc x;        // kind 0 or 1: the layer's input (B operand, all kp channels, transposed)
    int rows, kp, cl;  // kp = padded input channels (multiple of 16, <= 512), cl = output channels of the layer
    int split, tr;     // split operands; rows per tile (= the K extent of a tile's MMAs)
    float* partial;  // [gridDim.x][mblocks * 128][kp] fp32
    uint32_t off_a, a_bytes, off_b, b_bytes, off_vec_dz, off_vec_x, tmem_cols;
};

__global__ void __launch_bounds__(kTcThreads)
train_wgrad_kernel(const __grid_constant__ WgArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t built_bar;  // both operands of a row tile are in shared memory
    __shared__ __align__(8) uint64_t mma_bar;    // the tile's MMAs have read them
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler (uniform role branches)
    const int mb = blockIdx.y;
    const int tiles = (a.rows + a.tr - 1) / a.tr;
    const int my_tiles = ((int)blockIdx.x < tiles) ? (tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid == 0) {
        mbar_init(&built_bar, kTcEpiThreads);
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_sh, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;
    // K-major operands with K = the tile's rows: [tr/8 K-groups][rows of M or N][8 x bf16]; the K-group stride is padded
    // by one 16-byte row so that the transposing 2-byte stores of a warp (32 consecutive K) hit 32 distinct banks
    const uint32_t lbo_a = (kTcM + 1) * 16, lbo_b = (uint32_t)(a.kp + 1) * 16;

    if (warp == 5) {
        // whole warp converged, one elected lane issues (uniform datapath; see sa_mlp_tc.cu)
        const bool leader = elect_one_sync();
        const uint32_t smem_base = tc_opaque_u32(smem_u32(smem));
        const uint32_t built_a = tc_opaque_u32(smem_u32(&built_bar)), mma_bar_a = tc_opaque_u32(smem_u32(&mma_bar));
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);     // SBO = 128 B; descriptor version 1 (bit 46)
        const uint32_t a_base = (((smem_base + a.off_a) >> 4) & 0x3FFFu) | ((lbo_a >> 4) << 16);
        const uint32_t b_base = (((smem_base + a.off_b) >> 4) & 0x3FFFu) | ((lbo_b >> 4) << 16);
        const uint32_t a_step = (2u * lbo_a) >> 4, b_step = (2u * lbo_b) >> 4;
        const uint32_t a_copy = a.a_bytes >> 4, b_copy = a.b_bytes >> 4;
        const bool split = a.split != 0;
        const uint32_t nq = (uint32_t)a.tr / 16u;
        for (int t = 0; t < my_tiles; ++t) {
            mbar_wait_addr(built_a, t & 1);
            tc_fence_after();
            if (leader) {
                for (int n0 = 0; n0 < a.kp; n0 += 256) {
                    const int nc = min(256, a.kp - n0);
                    const uint32_t idesc = umma_idesc_bf16(kTcM, nc);
                    const uint32_t d_addr = tmem_base + (uint32_t)n0;
                    const uint32_t b_lo = b_base + (uint32_t)n0;   // n0 rows of 16 bytes
                    for (uint32_t q = 0; q < nq; ++q) {
                        const uint64_t ad = ((uint64_t)desc_hi << 32) | (a_base + q * a_step);
                        const uint64_t bd = ((uint64_t)desc_hi << 32) | (b_lo + q * b_step);
                        umma_bf16(d_addr, ad, bd, idesc, t > 0 || q > 0);
                        if (split) {
                            umma_bf16(d_addr, ad + a_copy, bd, idesc, true);
                            umma_bf16(d_addr, ad, bd + b_copy, idesc, true);
                        }
                    }
                }
                umma_commit_addr(mma_bar_a);
            }
            __syncwarp();
        }
    } else if (warp < 4) {
        const int m = tid;
        float* vdz = reinterpret_cast<float*>(smem + a.off_vec_dz);
        float* vx = reinterpret_cast<float*>(smem + a.off_vec_x);
        for (int i = tid; i < 5 * a.dz.C; i += kTcEpiThreads) vdz[i] = a.dz.vec[i];
        if (a.x.kind == 1)
            for (int i = tid; i < 2 * a.x.C; i += kTcEpiThreads) vx[i] = a.x.vec[i];
        // channels of this M block that do not exist (cl not a multiple of 128): their operand rows stay zero
        {
            uint4* A4 = reinterpret_cast<uint4*>(smem + a.off_a);
            const int n16 = (int)(a.a_bytes / 16) * (a.split ? 2 : 1);
            for (int i = tid; i < n16; i += kTcEpiThreads) A4[i] = make_uint4(0, 0, 0, 0);
        }
        epi_bar_sync();
        bf16* A = reinterpret_cast<bf16*>(smem + a.off_a);
        bf16* Bm = reinterpret_cast<bf16*>(smem + a.off_b);
        const size_t a_lo = a.a_bytes / 2, b_lo = a.b_bytes / 2;  // element offsets of the lo copies
        const int c_lo = mb * kTcM, c_hi = min(a.cl, c_lo + kTcM);
        const int kg = m >> 3, ke = m & 7;  // this row's K group and position inside it
        const bool split = a.split != 0;
        auto emit_a = [&](int kc, const float (&v)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const size_t e = ((size_t)kg * (kTcM + 1) + kc * 8 + i) * 8 + ke;
                const bf16 h = __float2bfloat16_rn(v[i]);
                A[e] = h;
                if (split) A[a_lo + e] = __float2bfloat16_rn(v[i] - __bfloat162float(h));
            }
        };
        auto emit_b = [&](int kc, const float (&v)[8]) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const size_t e = ((size_t)kg * (a.kp + 1) + kc * 8 + i) * 8 + ke;
                const bf16 h = __float2bfloat16_rn(v[i]);
                Bm[e] = h;
                if (split) Bm[b_lo + e] = __float2bfloat16_rn(v[i] - __bfloat162float(h));
            }
        };
        for (int t = 0; t < my_tiles; ++t) {
            const int tile = (int)blockIdx.x + t * (int)gridDim.x;
            const int r = tile * a.tr + m;
            const bool valid = r < a.rows;
            if (t > 0) mbar_wait(&mma_bar, (t - 1) & 1);  // the previous tile's MMAs have read the operands
            if (m < a.tr) {
                build_dz(a.dz, vdz, r, valid, c_lo, c_hi, emit_a);
                if (a.x.kind == 0) build_gather(a.x, r, valid, 0, a.kp, emit_b);
                else build_bnrelu(a.x, vx, r, valid, 0, a.kp, emit_b);
            }
            fence_proxy_async();
            mbar_arrive(&built_bar);
        }
        if (my_tiles > 0) {
            mbar_wait(&mma_bar, (my_tiles - 1) & 1);
            tc_fence_after();
        }
        // accumulator row m = output channel c_lo + m, columns = the layer's (operand-order) input channels
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        float* prow = a.partial + ((size_t)blockIdx.x * gridDim.y * kTcM + (size_t)mb * kTcM + m) * a.kp;
        for (int c0 = 0; c0 < a.kp; c0 += 16) {
            uint32_t rr[16];
            if (my_tiles > 0) {
                tmem_ld16_issue(taddr + (uint32_t)c0, rr);
                tmem_ld_wait(rr);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) rr[i] = 0u;
            }
            float4* dst = reinterpret_cast<float4*>(prow + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                dst[i] = make_float4(__uint_as_float(rr[4 * i]), __uint_as_float(rr[4 * i + 1]), __uint_as_float(rr[4 * i + 2]),
                                     __uint_as_float(rr[4 * i + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc(tmem_base, a.tmem_cols);
}

// dW[c, ref_k] = sum over splits of partial[split][c][op_k];  layer 0: operand column j < D is reference channel 3 + j,
// D <= j < D + 3 is coordinate j - D (feat_first = D);  other layers: identity (feat_first = -1)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int mrows, int kp, int cl, int cin,
                                    int feat_first, float* __restrict__ dw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= cl * cin) return;
    const int c = e / cin, k = e % cin;
    int j = k;
    if (feat_first >= 0) j = k < 3 ? feat_first + k : k - 3;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += partial[((size_t)s * mrows + c) * kp + j];
    dw[e] = acc;
}

// ---- weight packing (every step: the parameters change) -------------------------------------------------------------------
// forward operand of layer l: B[n, k] = W[n, k] -> chunks of <= 256 rows n, each [kp/8][nc][8] bf16, K zero padded;
// layer 0 uses the features-first K order.  transposed = the dgrad operand: B[n = input channel, k = output channel] = W[k, n].
// split: every chunk is followed by its lo copy (bf16(w - bf16(w))).
__global__ void train_pack_kernel(const float* __restrict__ w, int cout, int cin, int kp, int n, int feat_first, int transposed,
                                  int split, bf16* __restrict__ out) {
    const int total = kp * n;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        // e enumerates chunk-major: chunk ch (256 columns), then [kp/8][nc][8]
        int ch = 0, rem = e;
        int nc = min(256, n);
        size_t chunk_base = 0;  // element offset of the chunk's hi copy in `out`
        while (rem >= kp * nc) {
            rem -= kp * nc;
            chunk_base += (size_t)kp * nc * (split ? 2 : 1);
            ++ch;
            nc = min(256, n - ch * 256);
        }
        const int kc = rem / (nc * 8);
        const int r2 = rem % (nc * 8);
        const int nn = ch * 256 + r2 / 8, k = kc * 8 + (r2 % 8);
        float v = 0.f;
        if (!transposed) {
            // nn = output channel (< cout by construction), k = operand input channel
            int src = k;
            if (feat_first >= 0) src = k < feat_first ? 3 + k : k - feat_first;
            if (nn < cout && k < (feat_first >= 0 ? feat_first + 3 : cin)) v = w[(size_t)nn * cin + src];
        } else {
            // nn = operand input channel (layer 0: features-first order), k = output channel
            int src = nn;
            bool ok = nn < cin;
            if (feat_first >= 0) {
                ok = nn < feat_first + 3;
                src = nn < feat_first ? 3 + nn : nn - feat_first;
            }
            if (ok && k < cout) v = w[(size_t)k * cin + src];
        }
        const bf16 h = __float2bfloat16_rn(v);
        out[chunk_base + rem] = h;
        if (split) out[chunk_base + (size_t)kp * nc + rem] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

// ---- per-channel sums over the rows ------------------------------------------------------------------------------------------
// mode 0: (sum Z, sum Z^2);  mode 1: (sum dY, sum dY * xhat), xhat = (Z - mean) * invstd.   sums [2][C] fp64, pre-zeroed.
__global__ void __launch_bounds__(256)
col_sums_kernel(const float* __restrict__ Z, const float* __restrict__ dY, const float* __restrict__ mean_istd /*[2][C], mode 1*/,
                int rows, int C, int mode, double* __restrict__ sums) {
    extern __shared__ float sm[];  // [2][C]
    const int oct = C / 8;         // 16-byte groups per row
    const int lanes_r = blockDim.x / oct > 0 ? blockDim.x / oct : 1;
    const int o = threadIdx.x % oct, rl = threadIdx.x / oct;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
    if (rl < lanes_r) {
        float mu[8], is[8];
        if (mode == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                mu[i] = mean_istd[o * 8 + i];
                is[i] = mean_istd[C + o * 8 + i];
            }
        }
        const int rows_per_cta = (rows + gridDim.x - 1) / gridDim.x;
        const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
        for (int r = r0 + rl; r < r1; r += lanes_r) {
            float z[8];
            load_f32x8(Z + (size_t)r * C + o * 8, z);
            if (mode == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s1[i] += z[i];
                    s2[i] = __fmaf_rn(z[i], z[i], s2[i]);
                }
            } else {
                float d[8];
                load_f32x8(dY + (size_t)r * C + o * 8, d);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    s1[i] += d[i];
                    s2[i] = __fmaf_rn(d[i], (z[i] - mu[i]) * is[i], s2[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            atomicAdd(&sm[o * 8 + i], s1[i]);
            atomicAdd(&sm[C + o * 8 + i], s2[i]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], (double)sm[i]);
}

// pooled layer: dY is nonzero only at (argmax row, channel): s1 = sum dOut, s2 = sum dOut * xhat(argmax row)
__global__ void __launch_bounds__(512)
pooled_sums_kernel(const float* __restrict__ Z, const int* __restrict__ argmax, const float* __restrict__ dout,
                   const float* __restrict__ mean_istd, int G, int K, int C, double* __restrict__ sums) {
    extern __shared__ float sm[];
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
    const int c = threadIdx.x % C;
    const int gl = threadIdx.x / C, gstep = max(1, (int)blockDim.x / C);
    float s1 = 0.f, s2 = 0.f;
    if (gl < gstep) {
        const int per = (G + gridDim.x - 1) / gridDim.x;
        const int g0 = blockIdx.x * per, g1 = min(G, g0 + per);
        const float mu = mean_istd[c], is = mean_istd[C + c];
        for (int g = g0 + gl; g < g1; g += gstep) {
            const int k = argmax[(size_t)g * C + c];
            if (k < 0) continue;
            const float d = dout[(size_t)g * C + c];
            const float z = Z[((size_t)g * K + k) * C + c];
            s1 += d;
            s2 = __fmaf_rn(d, (z - mu) * is, s2);
        }
        atomicAdd(&sm[c], s1);
        atomicAdd(&sm[C + c], s2);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], (double)sm[i]);
}

// forward: sums -> stat = mean | invstd | a | b; running statistics exactly as nn.BatchNorm2d (momentum, unbiased variance)
__global__ void bn_finalize_fwd_kernel(const double* __restrict__ sums, int C, double n, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, float momentum, float* __restrict__ stat,
                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                       long long* __restrict__ num_batches_tracked) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && num_batches_tracked) *num_batches_tracked += 1;
    if (c >= C) return;
    const double mean = sums[c] / n;
    double var = sums[C + c] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float istd = (float)(1.0 / sqrt(var + (double)eps));
    const float aa = gamma[c] * istd;
    stat[c] = (float)mean;
    stat[C + c] = istd;
    stat[2 * C + c] = aa;
    stat[3 * C + c] = beta[c] - (float)mean * aa;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// backward: sums (s1, s2) -> vec = g | m1 | m2 | mean | invstd for the dZ build; grad_gamma = s2, grad_beta = s1
__global__ void bn_finalize_bwd_kernel(const double* __restrict__ sums, int C, double n, const float* __restrict__ gamma,
                                       const float* __restrict__ stat, float* __restrict__ vec, float* __restrict__ grad_gamma,
                                       float* __restrict__ grad_beta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float istd = stat[C + c];
    vec[c] = gamma[c] * istd;
    vec[C + c] = (float)(sums[c] / n);
    vec[2 * C + c] = (float)(sums[C + c] / n);
    vec[3 * C + c] = stat[c];
    vec[4 * C + c] = istd;
    if (grad_gamma) grad_gamma[c] = (float)sums[C + c];
    if (grad_beta) grad_beta[c] = (float)sums[c];
}

// out[g, c] = max_k relu(a_c Z[g, k, c] + b_c); argmax = the first k attaining a POSITIVE maximum, else -1
__global__ void __launch_bounds__(256)
pool_argmax_kernel(const float* __restrict__ Z, const float* __restrict__ stat, int G, int K, int C, float* __restrict__ out,
                   int* __restrict__ argmax) {
    const int oct = C / 8;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)G * oct) return;
    const int g = (int)(e / oct), o = (int)(e % oct);
    float aa[8], bb[8], best[8];
    int arg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        aa[i] = stat[2 * C + o * 8 + i];
        bb[i] = stat[3 * C + o * 8 + i];
        best[i] = 0.f;
        arg[i] = -1;
    }
    const float* base = Z + (size_t)g * K * C + o * 8;
    for (int k = 0; k < K; ++k) {
        float z[8];
        load_f32x8(base + (size_t)k * C, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float y = __fmaf_rn(aa[i], z[i], bb[i]);
            if (y > best[i]) {
                best[i] = y;
                arg[i] = k;
            }
        }
    }
    float* orow = out + (size_t)g * C + o * 8;
    int* arow = argmax + (size_t)g * C + o * 8;
    *reinterpret_cast<float4*>(orow) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<float4*>(orow + 4) = make_float4(best[4], best[5], best[6], best[7]);
    *reinterpret_cast<int4*>(arow) = make_int4(arg[0], arg[1], arg[2], arg[3]);
    *reinterpret_cast<int4*>(arow + 4) = make_int4(arg[4], arg[5], arg[6], arg[7]);
}

// ---- host side -------------------------------------------------------------------------------------------------------------
struct TrainPlan {
    bool ok;
    int rows, G, K;
    int cin[3], kp[3], c[3];   // real input channels, padded operand K, output channels
    size_t z_off[3], stat_off[3], argmax_off, saved_bytes;
    // forward workspace
    size_t wf_off[3], sums_off[3], ws_fwd;
    // backward workspace
    size_t wt_off[3], dy_off[2], bsums_off[3], vec_off[3], partial_off, ws_bwd;
    int wg_splits, split, wg_tile_rows;
};

static TrainPlan train_plan(int B, int S, int K, int D, const int* cout, int precision) {
    TrainPlan p = {};
    p.ok = false;
    if (precision != 0 && precision != 1) return p;
    p.split = precision == 0;
    const size_t copies = p.split ? 2 : 1;
    const size_t rows_sz = (size_t)B * S * K;
    if (B <= 0 || S <= 0 || K <= 0 || D < 0 || !cout || rows_sz == 0 || rows_sz >= (1u << 30)) return p;
    p.rows = (int)rows_sz;
    p.G = B * S;
    p.K = K;
    for (int l = 0; l < 3; ++l) {
        if (cout[l] <= 0 || cout[l] > kTrMaxC || (cout[l] % 16) != 0) return p;
        p.c[l] = cout[l];
        p.cin[l] = l == 0 ? 3 + D : cout[l - 1];
        p.kp[l] = (int)align_up((size_t)p.cin[l], 16);
        if (p.kp[l] > kTrMaxC) return p;
    }
    size_t off = 0;
    for (int l = 0; l < 3; ++l) {
        p.z_off[l] = off;
        off += align_up(rows_sz * p.c[l] * sizeof(float), 256);
    }
    for (int l = 0; l < 3; ++l) {
        p.stat_off[l] = off;
        off += align_up((size_t)4 * p.c[l] * sizeof(float), 256);
    }
    p.argmax_off = off;
    off += align_up((size_t)p.G * p.c[2] * sizeof(int), 256);
    p.saved_bytes = off;

    off = 0;
    for (int l = 0; l < 3; ++l) {
        p.wf_off[l] = off;
        off += align_up((size_t)p.kp[l] * p.c[l] * sizeof(bf16) * copies, 256);
    }
    for (int l = 0; l < 3; ++l) {
        p.sums_off[l] = off;
        off += align_up((size_t)2 * p.c[l] * sizeof(double), 256);
    }
    p.ws_fwd = off;

    off = 0;
    for (int l = 0; l < 3; ++l) {  // dgrad operand of layer l: [n = kp_l][k = c_l]
        p.wt_off[l] = off;
        off += align_up((size_t)p.c[l] * p.kp[l] * sizeof(bf16) * copies, 256);
    }
    for (int l = 0; l < 2; ++l) {  // dY of layers 0 and 1
        p.dy_off[l] = off;
        off += align_up(rows_sz * p.c[l] * sizeof(float), 256);
    }
    for (int l = 0; l < 3; ++l) {
        p.bsums_off[l] = off;
        off += align_up((size_t)2 * p.c[l] * sizeof(double), 256);
        p.vec_off[l] = off;
        off += align_up((size_t)5 * p.c[l] * sizeof(float), 256);
    }
    p.wg_tile_rows = p.split ? 64 : kTcM;
    const int tiles = (p.rows + p.wg_tile_rows - 1) / p.wg_tile_rows;
    p.wg_splits = tiles < num_sms() ? tiles : num_sms();
    size_t pmax = 0;
    for (int l = 0; l < 3; ++l) {
        const size_t mrows = align_up((size_t)p.c[l], kTcM);
        const size_t b = (size_t)p.wg_splits * mrows * p.kp[l] * sizeof(float);
        if (b > pmax) pmax = b;
    }
    p.partial_off = off;
    off += align_up(pmax, 256);
    p.ws_bwd = off;
    p.ok = true;
    return p;
}

static int launch_gemm(TgArgs& a, cudaStream_t stream) {
    // shared memory: A operand (hi, then lo when split; one K panel) | weight ring | A-build vectors | epilogue vectors
    const int kpan = a.split ? 256 : kTrMaxC;
    a.off_a = 0;
    a.a_bytes = (uint32_t)align_up((size_t)kTcM * (a.kp < kpan ? a.kp : kpan) * 2, 128);
    a.off_ring = a.a_bytes * (a.split ? 2 : 1);
    a.off_vec = a.off_ring + kTrStages * kTrStageBytes;
    a.off_ab = a.off_vec + (uint32_t)align_up((size_t)5 * kTrMaxC * sizeof(float), 128);
    const uint32_t smem = a.off_ab + (uint32_t)align_up((size_t)2 * kTrMaxC * sizeof(float), 128);
    a.tmem_cols = tmem_cols_pow2(a.n);  // chunk ch of <= 256 columns sits at column 256 * ch
    PCST_CUDA(cudaFuncSetAttribute(train_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles = (a.rows + kTcM - 1) / kTcM;
    train_gemm_kernel<<<tiles, kTcThreads, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "train_gemm_kernel");
}

static int launch_wgrad(WgArgs& a, int splits, cudaStream_t stream) {
    const int groups = a.tr / 8;
    const uint32_t copies = a.split ? 2 : 1;
    a.off_a = 0;
    a.a_bytes = (uint32_t)align_up((size_t)groups * (kTcM + 1) * 16, 128);
    a.off_b = a.a_bytes * copies;
    a.b_bytes = (uint32_t)align_up((size_t)groups * (a.kp + 1) * 16, 128);
    a.off_vec_dz = a.off_b + a.b_bytes * copies;
    a.off_vec_x = a.off_vec_dz + (uint32_t)align_up((size_t)5 * kTrMaxC * sizeof(float), 128);
    const uint32_t smem = a.off_vec_x + (uint32_t)align_up((size_t)2 * kTrMaxC * sizeof(float), 128);
    a.tmem_cols = tmem_cols_pow2(a.kp);
    PCST_CUDA(cudaFuncSetAttribute(train_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int mblocks = (a.cl + kTcM - 1) / kTcM;
    train_wgrad_kernel<<<dim3(splits, mblocks), kTcThreads, smem, stream>>>(a);
    return check_cuda(cudaGetLastError(), "train_wgrad_kernel");
}

static RowSrc gather_src(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int N, int S, int K,
                         int D) {
    RowSrc s = {};
    s.kind = 0;
    s.xyz = xyz; s.feats = feats; s.new_xyz = new_xyz; s.idx = idx;
    s.N = N; s.S = S; s.K = K; s.D = D;
    return s;
}

}  // namespace pcst

using namespace pcst;

extern "C" size_t pcst_sa_mlp_train_saved_bytes(int B, int S, int K, int D, const int* cout) {
    const TrainPlan p = train_plan(B, S, K, D, cout, 1);  // the saved blob does not depend on the precision mode
    return p.ok ? p.saved_bytes : 0;
}
extern "C" size_t pcst_sa_mlp_train_workspace_bytes(int B, int S, int K, int D, const int* cout, int precision, int backward) {
    const TrainPlan p = train_plan(B, S, K, D, cout, precision);
    if (!p.ok) return 0;
    return backward ? p.ws_bwd : p.ws_fwd;
}

extern "C" int pcst_sa_mlp_max_bnstats_bf16(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx,
                                            int B, int N, int S, int K, int D, const pcst_mlp3_train_t* mlp, int precision,
                                            float* out, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes,
                                            pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && mlp && out && saved && ws, "null pointer");
    PCST_CHECK_ARG(D == 0 || feats, "feats is NULL but D > 0");
    PCST_CHECK_ARG(idx || (S == 1 && K == N && !new_xyz), "idx == NULL means group_all: S = 1, K = N, new_xyz = NULL");
    PCST_CHECK_ARG(!idx || new_xyz, "new_xyz is required with idx");
    PCST_CHECK_ARG(precision == 0 || precision == 1, "precision must be 0 (split bf16x3 operands) or 1 (bf16 operands)");
    const TrainPlan p = train_plan(B, S, K, D, mlp->cout, precision);
    if (!p.ok) {
        set_error("pcst_sa_mlp_max_bnstats_bf16: unsupported shape (Cout must be a multiple of 16, <= 512; 3 + D <= 512)");
        return PCST_ERR_UNSUPPORTED;
    }
    if (saved_bytes < p.saved_bytes || ws_bytes < p.ws_fwd || ((uintptr_t)saved & 255) || ((uintptr_t)ws & 255)) {
        set_error("pcst_sa_mlp_max_bnstats_bf16: saved / workspace buffer too small or misaligned");
        return PCST_ERR_WORKSPACE;
    }
    for (int l = 0; l < 3; ++l) PCST_CHECK_ARG(mlp->w[l] && mlp->bias[l] && mlp->gamma[l] && mlp->beta[l], "null layer pointer");
    char* sv = (char*)saved;
    char* w = (char*)ws;
    PCST_CUDA(cudaMemsetAsync(w + p.sums_off[0], 0, p.ws_fwd - p.sums_off[0], stream));
    for (int l = 0; l < 3; ++l) {
        const int total = p.kp[l] * p.c[l];
        train_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(mlp->w[l], p.c[l], p.cin[l], p.kp[l], p.c[l],
                                                                   l == 0 ? D : -1, 0, p.split, (bf16*)(w + p.wf_off[l]));
        PCST_CUDA_NAMED(cudaGetLastError(), "train_pack_kernel");
    }
    const double n = (double)p.rows;
    for (int l = 0; l < 3; ++l) {
        TgArgs a = {};
        if (l == 0) {
            a.src = gather_src(xyz, feats, new_xyz, idx, N, S, K, D);
        } else {
            a.src.kind = 1;
            a.src.Z = (const float*)(sv + p.z_off[l - 1]);
            a.src.vec = (const float*)(sv + p.stat_off[l - 1]) + 2 * p.c[l - 1];  // a | b
            a.src.C = p.c[l - 1];
        }
        a.rows = p.rows; a.kp = p.kp[l]; a.n = p.c[l];
        a.split = p.split;
        a.wblob = (const unsigned char*)(w + p.wf_off[l]);
        a.epi = 0;
        a.bias = mlp->bias[l];
        a.out_z = (float*)(sv + p.z_off[l]);
        int st = launch_gemm(a, stream);
        if (st != PCST_OK) return st;
        const int C = p.c[l];
        int ctas = (p.rows + 255) / 256;
        if (ctas > 4 * num_sms()) ctas = 4 * num_sms();
        col_sums_kernel<<<ctas, 256, 2 * C * sizeof(float), stream>>>((const float*)(sv + p.z_off[l]), nullptr, nullptr, p.rows, C, 0,
                                                                       (double*)(w + p.sums_off[l]));
        PCST_CUDA_NAMED(cudaGetLastError(), "col_sums_kernel");
        bn_finalize_fwd_kernel<<<(C + 127) / 128, 128, 0, stream>>>(
            (const double*)(w + p.sums_off[l]), C, n, mlp->gamma[l], mlp->beta[l], mlp->eps, mlp->momentum,
            (float*)(sv + p.stat_off[l]), mlp->running_mean[l], mlp->running_var[l], (long long*)mlp->num_batches_tracked[l]);
        PCST_CUDA_NAMED(cudaGetLastError(), "bn_finalize_fwd_kernel");
    }
    const long items = (long)p.G * (p.c[2] / 8);
    pool_argmax_kernel<<<(unsigned)((items + 255) / 256), 256, 0, stream>>>(
        (const float*)(sv + p.z_off[2]), (const float*)(sv + p.stat_off[2]), p.G, K, p.c[2], out, (int*)(sv + p.argmax_off));
    return check_cuda(cudaGetLastError(), "pool_argmax_kernel");
}

extern "C" int pcst_sa_mlp_max_bwd_bf16(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B,
                                        int N, int S, int K, int D, const pcst_mlp3_train_t* mlp, int precision,
                                        const void* saved, size_t saved_bytes, const float* grad_out,
                                        const pcst_mlp3_grads_t* grads, float* grad_grouped, void* ws, size_t ws_bytes,
                                        pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && mlp && saved && grad_out && grads && ws, "null pointer");
    PCST_CHECK_ARG(D == 0 || feats, "feats is NULL but D > 0");
    PCST_CHECK_ARG(precision == 0 || precision == 1, "precision must be 0 (split bf16x3 operands) or 1 (bf16 operands)");
    const TrainPlan p = train_plan(B, S, K, D, mlp->cout, precision);
    if (!p.ok) {
        set_error("pcst_sa_mlp_max_bwd_bf16: unsupported shape");
        return PCST_ERR_UNSUPPORTED;
    }
    if (saved_bytes < p.saved_bytes || ws_bytes < p.ws_bwd || ((uintptr_t)saved & 255) || ((uintptr_t)ws & 255)) {
        set_error("pcst_sa_mlp_max_bwd_bf16: saved / workspace buffer too small or misaligned");
        return PCST_ERR_WORKSPACE;
    }
    const char* sv = (const char*)saved;
    char* w = (char*)ws;
    const double n = (double)p.rows;
    for (int l = 0; l < 3; ++l) PCST_CUDA(cudaMemsetAsync(w + p.bsums_off[l], 0, (size_t)2 * p.c[l] * sizeof(double), stream));
    for (int l = 0; l < 3; ++l) {
        if (l == 0 && !grad_grouped) continue;
        const int total = p.c[l] * p.kp[l];
        train_pack_kernel<<<(total + 255) / 256, 256, 0, stream>>>(mlp->w[l], p.c[l], p.cin[l], p.c[l], p.kp[l], l == 0 ? D : -1, 1,
                                                                   p.split, (bf16*)(w + p.wt_off[l]));
        PCST_CUDA_NAMED(cudaGetLastError(), "train_pack_kernel");
    }
    for (int l = 2; l >= 0; --l) {
        const int C = p.c[l];
        const float* Zl = (const float*)(sv + p.z_off[l]);
        const float* statl = (const float*)(sv + p.stat_off[l]);
        double* sums = (double*)(w + p.bsums_off[l]);
        float* vec = (float*)(w + p.vec_off[l]);
        // 1. the two per-channel sums of the BatchNorm backward
        if (l == 2) {
            int threads = 256;
            if (C > threads) threads = C;  // one thread per channel at least (C <= 512)
            int ctas = (p.G + 63) / 64;
            if (ctas > 2 * num_sms()) ctas = 2 * num_sms();
            pooled_sums_kernel<<<ctas, threads, 2 * C * sizeof(float), stream>>>(Zl, (const int*)(sv + p.argmax_off), grad_out, statl,
                                                                                   p.G, K, C, sums);
        } else {
            int ctas = (p.rows + 255) / 256;
            if (ctas > 4 * num_sms()) ctas = 4 * num_sms();
            col_sums_kernel<<<ctas, 256, 2 * C * sizeof(float), stream>>>(Zl, (const float*)(w + p.dy_off[l]), statl, p.rows, C, 1, sums);
        }
        PCST_CUDA_NAMED(cudaGetLastError(), "col_sums_kernel");
        bn_finalize_bwd_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sums, C, n, mlp->gamma[l], statl, vec, grads->gamma[l],
                                                                    grads->beta[l]);
        PCST_CUDA_NAMED(cudaGetLastError(), "bn_finalize_bwd_kernel");
        if (grads->bias[l]) PCST_CUDA(cudaMemsetAsync(grads->bias[l], 0, (size_t)C * sizeof(float), stream));  // sum_r dZ = 0 exactly

        RowSrc dz = {};
        dz.kind = l == 2 ? 3 : 2;
        dz.Z = Zl; dz.vec = vec; dz.C = C; dz.K = K;
        dz.dY = l == 2 ? nullptr : (const float*)(w + p.dy_off[l]);
        dz.argmax = (const int*)(sv + p.argmax_off);
        dz.dout = grad_out;
        RowSrc xin = {};
        if (l == 0) {
            xin = gather_src(xyz, feats, new_xyz, idx, N, S, K, D);
        } else {
            xin.kind = 1;
            xin.Z = (const float*)(sv + p.z_off[l - 1]);
            xin.vec = (const float*)(sv + p.stat_off[l - 1]) + 2 * p.c[l - 1];
            xin.C = p.c[l - 1];
        }
        // 2. wgrad
        if (grads->w[l]) {
            WgArgs wa = {};
            wa.dz = dz; wa.x = xin;
            wa.rows = p.rows; wa.kp = p.kp[l]; wa.cl = C;
            wa.split = p.split; wa.tr = p.wg_tile_rows;
            wa.partial = (float*)(w + p.partial_off);
            int st = launch_wgrad(wa, p.wg_splits, stream);
            if (st != PCST_OK) return st;
            const int mrows = (int)align_up((size_t)C, kTcM);
            const int total = C * p.cin[l];
            wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, stream>>>(wa.partial, p.wg_splits, mrows, p.kp[l], C, p.cin[l],
                                                                         l == 0 ? D : -1, grads->w[l]);
            PCST_CUDA_NAMED(cudaGetLastError(), "wgrad_reduce_kernel");
        }
        // 3. dgrad
        if (l > 0 || grad_grouped) {
            TgArgs a = {};
            a.src = dz;
            a.rows = p.rows; a.kp = C; a.n = p.kp[l];
            a.split = p.split;
            a.wblob = (const unsigned char*)(w + p.wt_off[l]);
            if (l > 0) {
                a.epi = 1;
                a.out_z = (float*)(w + p.dy_off[l - 1]);
                a.zprev = (const float*)(sv + p.z_off[l - 1]);
                a.ab = (const float*)(sv + p.stat_off[l - 1]) + 2 * p.c[l - 1];
            } else {
                a.epi = 2;
                a.out_f32 = grad_grouped;
                a.D = D;
            }
            int st = launch_gemm(a, stream);
            if (st != PCST_OK) return st;
        }
    }
    return PCST_OK;
}
