// noise_mlp_tc.cu -- NoisePredictor (models/diffusion_model.py:38-61) as ONE fused tcgen05 kernel per call (sm_100a).
//
// The reference evaluates, per point, a chain of nn.Linear layers on cuBLAS with separate bias / ReLU / residual kernels:
//   point_feat = Linear(3,128) -> ReLU -> Linear(128,256) -> ReLU -> Linear(256,F)
//   x = point_feat + time_proj(TimeEmbedding(t)) + style_proj(style)              (per batch element vectors, :55-57)
//   6 x:  x = Linear(2F,F)(ReLU(Linear(F,2F)(x))) + x                             (Dropout is identity in eval, :58-59)
//   out = Linear(F,256) -> ReLU -> Linear(256,128) -> ReLU -> Linear(128,3)
// = 3.5 MFLOP per point and ~20 activation round trips through HBM.  Here a CTA owns 128 points and runs the whole chain
// with activations in shared memory (bf16 K-major UMMA operands) and TMEM:
//   * the residual stream x lives in TMEM columns [0, F) as the fp32 ACCUMULATOR for the whole chain: the second Linear of
//     every block accumulates straight onto it (tcgen05.mma with accumulate), so the residual add costs nothing and x is
//     never rounded to bf16 except when it is read as the next GEMM's operand;
//   * every bias that lands on x (point_encoder.4, time / style projections, the blocks' second biases) is folded into a
//     per-(batch element, stage) shift vector computed once per call by a tiny prep kernel; the operand read of stage i
//     adds shift_i;
//   * the hidden 2F activations are produced and consumed in two halves of F columns (TMEM columns [256, 256 + F)), so
//     the widest operand in shared memory is 128 x 256 bf16;
//   * weights are packed once (bf16 [K/8][N][8] blocks in step order) and streamed through a 4-stage TMA ring; they stay
//     L2-resident (3.3 MB for F = 256).  Every 128-row tile needs ALL of them (3.3 MB per tile, 64 bytes per SM and clock at
//     the tensor floor -- more than the L2 delivers to 148 SMs at once), so the kernel can run as thread-block clusters of 2 or 4 CTAs
//     that walk their tiles in lockstep and share every stage: each CTA fetches 1/C of it and MULTICASTS it into all
//     shared memories (cp.async.bulk ... .multicast::cluster), and a stage is refilled once all MMA issuers have
//     committed it (tcgen05.commit.multicast::cluster onto every CTA's empty barrier).  Tuning key noise.cluster.
// Warp roles as in sa_mlp_tc.cu: warps 0-3 = operand build + epilogues (thread = row = TMEM lane), warp 4 = TMEM
// allocation + weight producer, warp 5 = MMA issuer.
// Precision: bf16 operands, fp32 accumulation / residual stream / biases -> within rtol 2e-2 of the reference's fp32 module.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pcst {

constexpr int kNpMaxSteps = 40;
constexpr int kNpStages = 4;
constexpr int kNpStageBytes = 16 * 1024;
constexpr int kNpMaxBlocks = 8;

struct NpStep {
    uint32_t a_off;    // shared-memory offset of the A operand
    uint32_t out_off;  // epi 1, 2: operand buffer written
    uint32_t w_off;    // offset of the step's weight block [kp/8][n][8] in the blob
    uint32_t bias_off; // epi 1, 3: first bias (float index into the bias table)
    uint16_t kp, n, tmem_col;
    uint8_t acc;       // accumulate onto the TMEM contents
    uint8_t epi;       // 0 none; 1 relu(acc + bias) -> bf16 operand; 2 (acc + shift[stage]) -> bf16 operand; 3 acc + bias -> out
    uint8_t stage;     // epi 2: which shift vector
};

struct NpArgs {
    const float* pts;   // [B, N, 3]
    float* out;         // [B, N, 3]
    int B, N, F, tiles_per_b;
    const unsigned char* blob;
    const float* bias;   // bias table inside the blob
    const float* shift;  // [B][nstage][F]
    int nstage, nsteps;
    NpStep st[kNpMaxSteps];
    uint32_t off_ring;
    uint32_t cluster;    // CTAs sharing every weight stage by multicast (1, 2 or 4)
    int ntiles;          // real row tiles; CTAs past them (the grid is padded to whole clusters) only keep the lockstep
};

__device__ __forceinline__ int np_chunk_rows(int n, int kp) {
    int ck = kNpStageBytes / (n * 2) / 16 * 16;
    if (ck < 16) ck = 16;
    return ck < kp ? ck : kp;
}

__global__ void __launch_bounds__(kTcThreads)
noise_mlp_kernel(const __grid_constant__ NpArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kNpStages];
    __shared__ __align__(8) uint64_t empty_bar[kNpStages];
    __shared__ __align__(8) uint64_t mma_bar;  // accumulator of an epilogue-bearing step is complete
    __shared__ __align__(8) uint64_t a_bar;    // the input operand / an epilogue's operand is written
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t C = a.cluster;
    const uint32_t rank = C > 1 ? cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << C) - 1u);
    const bool real_tile = (int)blockIdx.x < a.ntiles;
    const int b = real_tile ? blockIdx.x / a.tiles_per_b : 0;
    const int row0 = real_tile ? (blockIdx.x % a.tiles_per_b) * kTcM : a.N;   // a padding CTA owns no valid row
    if (tid == 0) {
        for (int s = 0; s < kNpStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], C);   // every CTA of the cluster commits to every CTA's empty barrier
        }
        mbar_init(&mma_bar, 1);
        mbar_init(&a_bar, kTcEpiThreads);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc(&tmem_base_sh, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (C > 1) cluster_sync_all();  // every peer's barriers exist before anything is multicast to them
    const uint32_t tmem_base = tmem_base_sh;

    if (warp == 4) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int s = 0; s < a.nsteps; ++s) {
                const NpStep& st = a.st[s];
                const int ck = np_chunk_rows(st.n, st.kp);
                for (int k0 = 0; k0 < st.kp; k0 += ck, ++it) {
                    const uint32_t stage = it % kNpStages;
                    if (it >= kNpStages) mbar_wait(&empty_bar[stage], ((it / kNpStages) - 1u) & 1u);
                    const int rowsk = min(ck, st.kp - k0);
                    const uint32_t bytes = (uint32_t)rowsk * st.n * 2u;
                    mbar_arrive_expect_tx(&full_bar[stage], bytes);   // the whole stage: this CTA's slice + the peers'
                    unsigned char* dst = smem + a.off_ring + stage * kNpStageBytes;
                    const unsigned char* src = a.blob + st.w_off + (size_t)k0 * st.n * 2u;
                    if (C > 1) {
                        const uint32_t slice = bytes / C;   // bytes is a multiple of 512
                        tma_load_1d_multicast(dst + rank * slice, src + rank * slice, slice, &full_bar[stage], cmask);
                    } else {
                        tma_load_1d(dst, src, bytes, &full_bar[stage]);
                    }
                }
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            uint32_t it = 0, seen = 0, need = 1;  // the input operand, then one event per operand-writing epilogue
            const uint32_t lbo_a = kTcM * 16;
            for (int s = 0; s < a.nsteps; ++s) {
                const NpStep& st = a.st[s];
                while (seen < need) {
                    mbar_wait(&a_bar, seen & 1u);
                    ++seen;
                }
                tc_fence_after();
                const int ck = np_chunk_rows(st.n, st.kp);
                const uint32_t idesc = umma_idesc_bf16(kTcM, st.n);
                const uint32_t a_addr = smem_u32(smem + st.a_off);
                const uint32_t lbo_w = (uint32_t)st.n * 16;
                const uint32_t d_addr = tmem_base + st.tmem_col;
                for (int k0 = 0; k0 < st.kp; k0 += ck, ++it) {
                    const uint32_t stage = it % kNpStages;
                    mbar_wait(&full_bar[stage], (it / kNpStages) & 1u);
                    tc_fence_after();
                    const uint32_t w_addr = smem_u32(smem + a.off_ring + stage * kNpStageBytes);
                    const int rowsk = min(ck, st.kp - k0);
                    for (int kk = 0; kk < rowsk / 16; ++kk) {
                        const int q = k0 / 16 + kk;
                        const uint64_t ad = umma_smem_desc(a_addr + (uint32_t)q * 2u * lbo_a, lbo_a, 128);
                        const uint64_t bd = umma_smem_desc(w_addr + (uint32_t)kk * 2u * lbo_w, lbo_w, 128);
                        umma_bf16(d_addr, ad, bd, idesc, st.acc || q > 0);
                    }
                    if (C > 1) umma_commit_multicast(&empty_bar[stage], cmask);
                    else umma_commit(&empty_bar[stage]);
                }
                if (st.epi) {
                    umma_commit(&mma_bar);
                    if (st.epi != 3) ++need;
                }
            }
        }
    } else {
        const int m = tid, row = row0 + m;
        const bool valid = row < a.N;
        // input operand: K = 16 = (x, y, z, 0 ...), two 16-byte K groups per row
        {
            float x = 0.f, y = 0.f, z = 0.f;
            if (valid) {
                const float* p = a.pts + ((size_t)b * a.N + row) * 3;
                x = p[0]; y = p[1]; z = p[2];
            }
            unsigned char* A0 = smem + a.st[0].a_off;
            __nv_bfloat162 xy = __floats2bfloat162_rn(x, y), z0 = __floats2bfloat162_rn(z, 0.f);
            *reinterpret_cast<uint4*>(A0 + (size_t)m * 16) =
                make_uint4(*reinterpret_cast<uint32_t*>(&xy), *reinterpret_cast<uint32_t*>(&z0), 0u, 0u);
            *reinterpret_cast<uint4*>(A0 + ((size_t)kTcM + m) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        mbar_arrive(&a_bar);

        const uint32_t taddr_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        uint32_t phase = 0;
        for (int s = 0; s < a.nsteps; ++s) {
            const NpStep& st = a.st[s];
            if (!st.epi) continue;
            mbar_wait(&mma_bar, phase & 1u);
            ++phase;
            tc_fence_after();
            const uint32_t taddr = taddr_lane + st.tmem_col;
            if (st.epi == 3) {
                uint32_t rr[16];
                tmem_ld16_issue(taddr, rr);
                tmem_ld_wait(rr);
                if (valid) {
                    float* o = a.out + ((size_t)b * a.N + row) * 3;
                    const float* bs = a.bias + st.bias_off;
                    o[0] = __uint_as_float(rr[0]) + __ldg(bs);
                    o[1] = __uint_as_float(rr[1]) + __ldg(bs + 1);
                    o[2] = __uint_as_float(rr[2]) + __ldg(bs + 2);
                }
                continue;
            }
            // per-channel additive vector of this epilogue: a bias (epi 1) or the batch element's shift of the stage (epi 2)
            const float* add = st.epi == 1 ? a.bias + st.bias_off : a.shift + ((size_t)b * a.nstage + st.stage) * a.F;
            unsigned char* outp = smem + st.out_off;
            for (int c0 = 0; c0 < st.n; c0 += 32) {
                uint32_t r[2][16];
                tmem_ld16_issue(taddr + (uint32_t)c0, r[0]);
                const bool two = c0 + 16 < st.n;
                if (two) tmem_ld16_issue(taddr + (uint32_t)c0 + 16u, r[1]);
                tmem_ld_wait(r[0]);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h == 1 && !two) break;
                    const int cb = c0 + 16 * h;
                    uint32_t packed[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 ad = __ldg(reinterpret_cast<const float4*>(add + cb + 4 * i));
                        float y0 = __uint_as_float(r[h][4 * i]) + ad.x, y1 = __uint_as_float(r[h][4 * i + 1]) + ad.y;
                        float y2 = __uint_as_float(r[h][4 * i + 2]) + ad.z, y3 = __uint_as_float(r[h][4 * i + 3]) + ad.w;
                        if (st.epi == 1) {
                            y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f);
                        }
                        __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
                        packed[2 * i] = *reinterpret_cast<uint32_t*>(&p0);
                        packed[2 * i + 1] = *reinterpret_cast<uint32_t*>(&p1);
                    }
                    *reinterpret_cast<uint4*>(outp + ((size_t)(cb >> 3) * kTcM + m) * 16) =
                        make_uint4(packed[0], packed[1], packed[2], packed[3]);
                    *reinterpret_cast<uint4*>(outp + ((size_t)((cb >> 3) + 1) * kTcM + m) * 16) =
                        make_uint4(packed[4], packed[5], packed[6], packed[7]);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&a_bar);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast or commit into it
    if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ---- prep: per batch element, shift[b][0] = b_pe2 + time_proj(TimeEmbedding(t_b)) + style_proj(style_b);
//            shift[b][i] = shift[b][i-1] + b2_{i-1}    (models/diffusion_model.py:15-26,55-59)
__global__ void noise_prep_kernel(const long long* __restrict__ timestep, const float* __restrict__ style, int F, int T, int nblocks,
                                  const float* __restrict__ time_w, const float* __restrict__ time_b,
                                  const float* __restrict__ style_w, const float* __restrict__ style_b,
                                  const float* __restrict__ pe2_b, const float* __restrict__ b2 /*[nblocks][F]*/,
                                  float* __restrict__ shift /*[B][nblocks + 1][F]*/) {
    extern __shared__ float emb[];  // [T]
    const int b = blockIdx.x;
    const int half = T / 2;
    const float t = (float)timestep[b];
    const float scale = logf(10000.0f) / (float)(half - 1);
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float f = expf((float)j * -scale);
        const float v = t * f;
        emb[j] = sinf(v);
        emb[half + j] = cosf(v);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < F; c += blockDim.x) {
        float acc = time_b[c];
        for (int j = 0; j < T; ++j) acc = fmaf(time_w[(size_t)c * T + j], emb[j], acc);
        float acs = style_b[c];
        for (int j = 0; j < F; ++j) acs = fmaf(style_w[(size_t)c * F + j], style[(size_t)b * F + j], acs);
        float s = pe2_b[c] + acc + acs;
        float* o = shift + (size_t)b * (nblocks + 1) * F + c;
        o[0] = s;
        for (int i = 0; i < nblocks; ++i) {
            s += b2[(size_t)i * F + c];
            o[(size_t)(i + 1) * F] = s;
        }
    }
}

// fp32 W[n0 + n, k0 + k] (row-major [rows, cin]) -> bf16 [kp/8][nlen][8], zero padded in K and N
__global__ void noise_pack_block_kernel(const float* __restrict__ w, int rows, int cin, int n0, int nlen, int k0, int klen, int kp,
                                        __nv_bfloat16* __restrict__ out) {
    const int total = kp * nlen;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kc = e / (nlen * 8);
        const int rem = e % (nlen * 8);
        const int n = rem / 8, k = kc * 8 + (rem % 8);
        float v = 0.f;
        if (n0 + n < rows && k < klen) v = w[(size_t)(n0 + n) * cin + k0 + k];
        out[e] = __float2bfloat16_rn(v);
    }
}
__global__ void noise_copy_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int pad_to) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pad_to; i += gridDim.x * blockDim.x) dst[i] = i < n ? src[i] : 0.f;
}

// ---- host: the step table and the blob layout -------------------------------------------------------------------------------
struct NpPlan {
    bool ok;
    int F, T, nblocks, nsteps;
    NpStep st[kNpMaxSteps];
    // which reference matrix a step's weight block is cut from
    struct Src { int which, block, n0, nlen, k0, klen; } src[kNpMaxSteps];  // which: 0-2 pe, 3 blk_w1, 4 blk_w2, 5-7 out
    uint32_t bias_floats;  // bias table: pe0 | pe1 | blk b1 x nblocks | out0 | out1 | out2(16)
    uint32_t off_bias, off_time_w, off_time_b, off_style_w, off_style_b, off_pe2_b, off_b2;
    size_t blob_bytes;
    uint32_t off_x, off_h, off_ring, smem_bytes;
};

static NpPlan np_plan(int F, int T, int nblocks) {
    NpPlan p = {};
    p.ok = false;
    if (F < 16 || F > 256 || (F % 16) != 0 || T < 4 || (T % 2) != 0 || T > 1024 || nblocks < 0 || nblocks > kNpMaxBlocks) return p;
    p.F = F; p.T = T; p.nblocks = nblocks;
    p.off_x = 0;
    p.off_h = kTcM * 256 * 2;
    p.off_ring = 2 * kTcM * 256 * 2;
    p.smem_bytes = p.off_ring + kNpStages * kNpStageBytes;
    uint32_t woff = 0, boff = 0;
    const uint32_t HC = 256;  // TMEM column of the hidden accumulators (x occupies [0, F))
    auto add = [&](uint32_t a_off, uint32_t out_off, int kp, int n, int col, int acc, int epi, int stage, uint32_t bias_off,
                   int which, int block, int n0, int nlen, int k0, int klen) {
        NpStep& s = p.st[p.nsteps];
        s.a_off = a_off; s.out_off = out_off; s.w_off = woff; s.bias_off = bias_off;
        s.kp = (uint16_t)kp; s.n = (uint16_t)n; s.tmem_col = (uint16_t)col;
        s.acc = (uint8_t)acc; s.epi = (uint8_t)epi; s.stage = (uint8_t)stage;
        p.src[p.nsteps] = {which, block, n0, nlen, k0, klen};
        woff += (uint32_t)align_up((size_t)kp * n * 2, 128);
        ++p.nsteps;
    };
    // point encoder
    add(p.off_x, p.off_h, 16, 128, HC, 0, 1, 0, boff, 0, 0, 0, 128, 0, 3);            boff += 128;
    add(p.off_h, p.off_x, 128, 256, HC, 0, 1, 0, boff, 1, 0, 0, 256, 0, 128);          boff += 256;
    add(p.off_x, p.off_x, 256, F, 0, 0, 2, 0, 0, 2, 0, 0, F, 0, 256);                   // x = pe2(.) ; operand read adds shift_0
    // residual blocks: the hidden layer in two halves of F columns
    for (int i = 0; i < nblocks; ++i) {
        for (int h = 0; h < 2; ++h) {
            add(p.off_x, p.off_h, F, F, HC, 0, 1, 0, boff + h * F, 3, i, h * F, F, 0, F);   // relu(W1[hF:(h+1)F, :] x + b1)
            add(p.off_h, p.off_x, F, F, 0, 1, h == 1 ? 2 : 0, i + 1, 0, 4, i, 0, F, h * F, F);  // x += W2[:, hF:(h+1)F] h
        }
        boff += 2 * F;
    }
    // output MLP
    add(p.off_x, p.off_h, F, 256, HC, 0, 1, 0, boff, 5, 0, 0, 256, 0, F);             boff += 256;
    add(p.off_h, p.off_x, 256, 128, HC, 0, 1, 0, boff, 6, 0, 0, 128, 0, 256);         boff += 128;
    add(p.off_x, 0, 128, 16, HC, 0, 3, 0, boff, 7, 0, 0, 16, 0, 128);                 boff += 16;
    p.bias_floats = boff;
    uint32_t off = woff;
    auto take = [&](size_t floats) { const uint32_t o = off; off += (uint32_t)align_up(floats * sizeof(float), 256); return o; };
    p.off_bias = take(boff);
    p.off_time_w = take((size_t)F * T);
    p.off_time_b = take(F);
    p.off_style_w = take((size_t)F * F);
    p.off_style_b = take(F);
    p.off_pe2_b = take(F);
    p.off_b2 = take((size_t)(nblocks > 0 ? nblocks : 1) * F);
    p.blob_bytes = off;
    p.ok = true;
    return p;
}

}  // namespace pcst

using namespace pcst;

extern "C" size_t pcst_noise_predictor_packed_bytes(int feature_dim, int time_dim, int nblocks) {
    const NpPlan p = np_plan(feature_dim, time_dim, nblocks);
    return p.ok ? p.blob_bytes : 0;
}
extern "C" size_t pcst_noise_predictor_workspace_bytes(int B, int feature_dim, int nblocks) {
    if (B <= 0 || feature_dim <= 0 || nblocks < 0) return 0;
    return align_up((size_t)B * (nblocks + 1) * feature_dim * sizeof(float), 256);
}

extern "C" int pcst_noise_predictor_pack_f32(const pcst_noise_mlp_t* m, void* packed, size_t packed_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(m && packed, "null pointer");
    const NpPlan p = np_plan(m->feature_dim, m->time_dim, m->nblocks);
    if (!p.ok) {
        set_error("pcst_noise_predictor_pack_f32: unsupported sizes (feature_dim a multiple of 16 in [16, 256], nblocks <= 8)");
        return PCST_ERR_UNSUPPORTED;
    }
    PCST_CHECK_ARG(packed_bytes >= p.blob_bytes && ((uintptr_t)packed & 255) == 0, "packed buffer too small or misaligned");
    const int F = p.F;
    char* blob = (char*)packed;
    for (int s = 0; s < p.nsteps; ++s) {
        const auto& src = p.src[s];
        const float* w = nullptr;
        int rows = 0, cin = 0;
        switch (src.which) {
            case 0: w = m->pe_w[0]; rows = 128; cin = 3; break;
            case 1: w = m->pe_w[1]; rows = 256; cin = 128; break;
            case 2: w = m->pe_w[2]; rows = F; cin = 256; break;
            case 3: w = m->blk_w1[src.block]; rows = 2 * F; cin = F; break;
            case 4: w = m->blk_w2[src.block]; rows = F; cin = 2 * F; break;
            case 5: w = m->out_w[0]; rows = 256; cin = F; break;
            case 6: w = m->out_w[1]; rows = 128; cin = 256; break;
            default: w = m->out_w[2]; rows = 3; cin = 128; break;
        }
        PCST_CHECK_ARG(w, "null weight pointer");
        const int total = p.st[s].kp * p.st[s].n;
        noise_pack_block_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, rows, cin, src.n0, src.nlen, src.k0, src.klen, p.st[s].kp,
                                                                         (__nv_bfloat16*)(blob + p.st[s].w_off));
        PCST_CUDA(cudaGetLastError());
    }
    // bias table in step order
    float* bias = (float*)(blob + p.off_bias);
    uint32_t at = 0;
    auto put = [&](const float* src, int n, int pad) -> int {
        if (!src) {
            set_error("pcst_noise_predictor_pack_f32: null bias pointer");
            return PCST_ERR_INVALID;
        }
        noise_copy_kernel<<<(pad + 255) / 256, 256, 0, stream>>>(src, n, bias + at, pad);
        at += pad;
        return check_cuda(cudaGetLastError(), "noise_copy_kernel");
    };
    int st;
    if ((st = put(m->pe_b[0], 128, 128)) != PCST_OK) return st;
    if ((st = put(m->pe_b[1], 256, 256)) != PCST_OK) return st;
    for (int i = 0; i < p.nblocks; ++i)
        if ((st = put(m->blk_b1[i], 2 * F, 2 * F)) != PCST_OK) return st;
    if ((st = put(m->out_b[0], 256, 256)) != PCST_OK) return st;
    if ((st = put(m->out_b[1], 128, 128)) != PCST_OK) return st;
    if ((st = put(m->out_b[2], 3, 16)) != PCST_OK) return st;
    auto copyf = [&](const float* src, size_t n, uint32_t off) -> int {
        if (!src) {
            set_error("pcst_noise_predictor_pack_f32: null pointer");
            return PCST_ERR_INVALID;
        }
        return check_cuda(cudaMemcpyAsync(blob + off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, stream), "cudaMemcpyAsync");
    };
    if ((st = copyf(m->time_w, (size_t)F * p.T, p.off_time_w)) != PCST_OK) return st;
    if ((st = copyf(m->time_b, F, p.off_time_b)) != PCST_OK) return st;
    if ((st = copyf(m->style_w, (size_t)F * F, p.off_style_w)) != PCST_OK) return st;
    if ((st = copyf(m->style_b, F, p.off_style_b)) != PCST_OK) return st;
    if ((st = copyf(m->pe_b[2], F, p.off_pe2_b)) != PCST_OK) return st;
    for (int i = 0; i < p.nblocks; ++i)
        if ((st = copyf(m->blk_b2[i], F, p.off_b2 + (uint32_t)(i * F * sizeof(float)))) != PCST_OK) return st;
    return PCST_OK;
}

extern "C" int pcst_noise_predictor_f32(const float* points, const int64_t* timestep, const float* style, int B, int N,
                                        int feature_dim, int time_dim, int nblocks, const void* packed, float* out, void* ws,
                                        size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(points && timestep && style && packed && out && ws, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0, "B and N must be positive");
    const NpPlan p = np_plan(feature_dim, time_dim, nblocks);
    if (!p.ok) {
        set_error("pcst_noise_predictor_f32: unsupported sizes");
        return PCST_ERR_UNSUPPORTED;
    }
    if (ws_bytes < pcst_noise_predictor_workspace_bytes(B, feature_dim, nblocks) || ((uintptr_t)ws & 255) || ((uintptr_t)packed & 255)) {
        set_error("pcst_noise_predictor_f32: workspace too small or misaligned");
        return PCST_ERR_WORKSPACE;
    }
    const char* blob = (const char*)packed;
    const int F = p.F;
    noise_prep_kernel<<<B, 256, p.T * sizeof(float), stream>>>(
        (const long long*)timestep, style, F, p.T, p.nblocks, (const float*)(blob + p.off_time_w), (const float*)(blob + p.off_time_b),
        (const float*)(blob + p.off_style_w), (const float*)(blob + p.off_style_b), (const float*)(blob + p.off_pe2_b),
        (const float*)(blob + p.off_b2), (float*)ws);
    PCST_CUDA(cudaGetLastError());
    NpArgs a = {};
    a.pts = points; a.out = out;
    a.B = B; a.N = N; a.F = F;
    a.tiles_per_b = (N + kTcM - 1) / kTcM;
    a.blob = (const unsigned char*)blob;
    a.bias = (const float*)(blob + p.off_bias);
    a.shift = (const float*)ws;
    a.nstage = p.nblocks + 1;
    a.nsteps = p.nsteps;
    for (int s = 0; s < p.nsteps; ++s) a.st[s] = p.st[s];
    a.off_ring = p.off_ring;
    PCST_CUDA(cudaFuncSetAttribute(noise_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    const long tiles = (long)B * a.tiles_per_b;
    PCST_CHECK_ARG(tiles < (1L << 30), "too many rows");
    int C = tuning("noise.cluster", 0);
    if (C != 2 && C != 4) C = 1;   // measured (profiles/r02/noise_cluster_ab.log): the serial MMA -> epilogue chain per tile, not
                                   // the L2, bounds the kernel today, and the cluster's lockstep costs 18 %: opt-in
    a.cluster = (uint32_t)C;
    a.ntiles = (int)tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((tiles + C - 1) / C * C));
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PCST_CUDA(cudaLaunchKernelEx(&cfg, noise_mlp_kernel, a));
    return PCST_OK;
}
