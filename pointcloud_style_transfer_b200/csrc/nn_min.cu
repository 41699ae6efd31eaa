// nn_min.cu -- nearest-neighbour minimum reduction (Chamfer loss / cdist metrics), sm_100a.
//
// Replaces the chunked bmm + clamp + min loops of models/losses.py:29-59 and the full
// torch.cdist matrix + min of evaluation/metrics.py:32-40,99-103.
//
// Bound: FP32 CUDA cores.  Algorithmic work = 5 FP32-pipe instructions (mul, fma, fma, add, fma)
// + 1 min per (row, candidate) pair -- "8 flop per pair" in SURVEY.md §8(d).  HBM traffic is
// negligible ((N+M)*16 B packed in, N*4 B out); candidates are streamed through shared memory in
// 16 KiB tiles by 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) behind a 3-stage mbarrier ring.
//
// Layout: each thread keeps R rows (x, y, z, |a|^2) in registers and scans a range of candidate
// tiles; candidates are read from shared memory as broadcast float4 (one LDS.128 per candidate per
// warp).  The grid is (row tiles) x (candidate splits) x B; partial minima of the splits are merged
// with atomicMin on the IEEE bits (all values are >= +0 after the clamp), or on a 64-bit
// (bits << 32 | index) key when the argmin is requested, which also yields the lowest-index tie-break.
#include "common.cuh"

namespace pcst {

constexpr int kNNThreads = 256;
constexpr int kNNStages = 3;

template <int R>
struct RowRegs {
    float x[R], y[R], z[R], n[R];
};

// One pair in the reference's rounding order.  FORM 0: losses.py:38 ((|a|^2 + |b|^2) + (-2 dot)).
// FORM 1/2: ATen _euclidean_dist K=5 chain; x,y,z are pre-scaled by -2 (exact), the two norms are
// added last, x1's first (FORM 1: rows are x1; FORM 2: candidates are x1).
template <int FORM>
__device__ __forceinline__ float pair_dist(float ax, float ay, float az, float an, const float4& c) {
    if (FORM == 0) {
        float t = __fadd_rn(an, c.w);
        float dot = dot3_chain(ax, ay, az, c.x, c.y, c.z);
        return __fmaf_rn(-2.0f, dot, t);
    } else {
        float r = __fmul_rn(ax, c.x);
        r = __fmaf_rn(ay, c.y, r);
        r = __fmaf_rn(az, c.z, r);
        if (FORM == 1) {
            r = __fadd_rn(r, an);
            r = __fadd_rn(r, c.w);
        } else {
            r = __fadd_rn(r, c.w);
            r = __fadd_rn(r, an);
        }
        return r;
    }
}

template <int FORM, int R, bool ARG, bool VEC2>
__global__ void __launch_bounds__(kNNThreads, 2)
nn_min_kernel(const float4* __restrict__ A, const float4* __restrict__ Bp, int N, int Npad, int Mpad,
              int tiles_per_split, unsigned int* __restrict__ rowmin_bits,
              unsigned long long* __restrict__ rowkey) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tiles = reinterpret_cast<float4*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[kNNStages];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int total_tiles = Mpad / kTilePoints;
    const int tile0 = blockIdx.y * tiles_per_split;
    int ntiles = total_tiles - tile0;
    if (ntiles > tiles_per_split) ntiles = tiles_per_split;
    if (ntiles <= 0) return;  // uniform per CTA
    const float4* cand = Bp + (size_t)b * Mpad + (size_t)tile0 * kTilePoints;

    if (tid == 0) {
        for (int s = 0; s < kNNStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int pre = ntiles < kNNStages ? ntiles : kNNStages;
        for (int s = 0; s < pre; ++s) {
            mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
            tma_load_1d(tiles + (size_t)s * kTilePoints, cand + (size_t)s * kTilePoints, kTileBytes, &full_bar[s]);
        }
    }

    // rows of this thread (coalesced float4 loads; A is sentinel-padded to a whole row tile)
    const int row0 = blockIdx.x * (kNNThreads * R) + tid;
    const float4* arow = A + (size_t)b * Npad;
    float ax[R], ay[R], az[R], an[R], best[R];
    int bj[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float4 v = arow[row0 + r * kNNThreads];
        if (FORM == 0) {
            ax[r] = v.x; ay[r] = v.y; az[r] = v.z;
        } else {
            ax[r] = -2.0f * v.x; ay[r] = -2.0f * v.y; az[r] = -2.0f * v.z;  // x1.mul(-2): exact
        }
        an[r] = v.w;
        best[r] = __int_as_float(0x7f800000);
        bj[r] = 0;
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kNNStages;
        mbar_wait(&full_bar[s], (uint32_t)((t / kNNStages) & 1));
        const float4* tile = tiles + (size_t)s * kTilePoints;
        const int jbase = (tile0 + t) * kTilePoints;
        if (!ARG && VEC2 && (R % 2 == 0)) {
            // packed fp32x2 path: two rows per FFMA2 / FADD2 / FMUL2 issue slot
#pragma unroll 4
            for (int j = 0; j < kTilePoints; ++j) {
                const float4 c = tile[j];
                const float2 cx = make_float2(c.x, c.x), cy = make_float2(c.y, c.y), cz = make_float2(c.z, c.z),
                             cw = make_float2(c.w, c.w);
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    const float2 vx = make_float2(ax[r], ax[r + 1]), vy = make_float2(ay[r], ay[r + 1]),
                                 vz = make_float2(az[r], az[r + 1]), vn = make_float2(an[r], an[r + 1]);
                    float2 d;
                    if (FORM == 0) {
                        float2 tt = __fadd2_rn(vn, cw);
                        float2 dot = __ffma2_rn(vz, cz, __ffma2_rn(vy, cy, __fmul2_rn(vx, cx)));
                        d = __ffma2_rn(make_float2(-2.0f, -2.0f), dot, tt);
                    } else {
                        float2 rr = __ffma2_rn(vz, cz, __ffma2_rn(vy, cy, __fmul2_rn(vx, cx)));
                        if (FORM == 1) {
                            rr = __fadd2_rn(rr, vn);
                            d = __fadd2_rn(rr, cw);
                        } else {
                            rr = __fadd2_rn(rr, cw);
                            d = __fadd2_rn(rr, vn);
                        }
                    }
                    best[r] = fminf(best[r], d.x);
                    best[r + 1] = fminf(best[r + 1], d.y);
                }
            }
        } else {
#pragma unroll 4
            for (int j = 0; j < kTilePoints; ++j) {
                const float4 c = tile[j];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float d = pair_dist<FORM>(ax[r], ay[r], az[r], an[r], c);
                    if (ARG) {
                        d = fmaxf(d, 0.0f);  // clamp per pair so ties at 0 keep the first index
                        if (d < best[r]) {
                            best[r] = d;
                            bj[r] = jbase + j;
                        }
                    } else {
                        best[r] = fminf(best[r], d);
                    }
                }
            }
        }
        __syncthreads();  // all warps are done with stage s
        if (tid == 0 && t + kNNStages < ntiles) {
            mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
            tma_load_1d(tiles + (size_t)s * kTilePoints, cand + (size_t)(t + kNNStages) * kTilePoints, kTileBytes,
                        &full_bar[s]);
        }
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r * kNNThreads;
        if (i < N) {
            const float v = fmaxf(best[r], 0.0f);  // clamp(min=0); min and clamp commute
            if (ARG) {
                unsigned long long key = ((unsigned long long)__float_as_uint(v) << 32) | (unsigned int)bj[r];
                atomicMin(rowkey + (size_t)b * N + i, key);
            } else {
                atomicMin(rowmin_bits + (size_t)b * N + i, __float_as_uint(v));
            }
        }
    }
}

// init: rowmin = +inf bits (or key = all ones)
__global__ void nn_min_init_kernel(unsigned int* rowmin_bits, unsigned long long* rowkey, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        if (rowkey) rowkey[i] = ~0ull;
        else rowmin_bits[i] = 0x7f800000u;
    }
}

// finalize: sqrt for the cdist forms; unpack (value, index) keys
__global__ void nn_min_finalize_kernel(float* rowmin, int64_t* rowarg, const unsigned long long* rowkey, size_t n,
                                       int do_sqrt) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float v;
        if (rowkey) {
            unsigned long long k = rowkey[i];
            v = __uint_as_float((unsigned int)(k >> 32));
            rowarg[i] = (int64_t)(unsigned int)(k & 0xffffffffull);
        } else {
            v = rowmin[i];
        }
        rowmin[i] = do_sqrt ? __fsqrt_rn(v) : v;
    }
}

struct NNPlan {
    int Npad, Mpad, row_tiles, splits, tiles_per_split, R;
    size_t off_a, off_b, off_key, total;
};

static NNPlan make_plan(int B, int N, int M) {
    NNPlan p;
    p.R = 4;
    const int rows_per_cta = kNNThreads * p.R;
    p.Npad = (int)align_up(align_up((size_t)N, rows_per_cta), kTilePoints);
    p.Mpad = padded_points(M);
    p.row_tiles = (N + rows_per_cta - 1) / rows_per_cta;
    const int cand_tiles = p.Mpad / kTilePoints;
    // choose the number of candidate splits that best fills whole waves of 2 CTAs/SM x 148 SMs
    const long wave = 2L * kNumSMs;
    int forced = tuning("nn_min.splits", 0);
    int best_s = 1;
    double best_eff = -1.0;
    for (int s = 1; s <= cand_tiles && s <= 256; ++s) {
        if (forced) s = forced < cand_tiles ? forced : cand_tiles;
        const int tps = (cand_tiles + s - 1) / s;
        const int s_eff = (cand_tiles + tps - 1) / tps;
        const long ctas = (long)p.row_tiles * s_eff * B;
        const long waves = (ctas + wave - 1) / wave;
        // per-CTA cost = tps tiles + ~0.35 tile of fixed overhead (row load, atomics, launch tail)
        const double eff = ((double)p.row_tiles * cand_tiles * B) / ((double)waves * wave * (tps + 0.35));
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_s = s_eff;
        }
        if (forced) break;
    }
    p.tiles_per_split = (cand_tiles + best_s - 1) / best_s;
    p.splits = (cand_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
    p.off_a = 0;
    p.off_b = align_up(p.off_a + (size_t)B * p.Npad * sizeof(float4), 256);
    p.off_key = align_up(p.off_b + (size_t)B * p.Mpad * sizeof(float4), 256);
    p.total = align_up(p.off_key + (size_t)B * N * sizeof(unsigned long long), 256);
    return p;
}

template <int FORM, bool ARG, bool VEC2>
static int launch_main(const NNPlan& p, const float4* A, const float4* Bp, int B, int N, unsigned int* bits,
                       unsigned long long* keys, cudaStream_t stream) {
    auto kern = nn_min_kernel<FORM, 4, ARG, VEC2>;
    const int smem = kNNStages * kTileBytes;
    PCST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(p.row_tiles, p.splits, B);
    kern<<<grid, kNNThreads, smem, stream>>>(A, Bp, N, p.Npad, p.Mpad, p.tiles_per_split, bits, keys);
    return check_cuda(cudaGetLastError(), "nn_min_kernel");
}

}  // namespace pcst

using namespace pcst;

extern "C" size_t pcst_nn_min_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return make_plan(B, N, M).total;
}

extern "C" int pcst_nn_min_f32(const float* a, const float* b, int B, int N, int M, int form, float* rowmin,
                               int64_t* rowarg, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(a && b && rowmin, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && M > 0, "B, N, M must be positive");
    PCST_CHECK_ARG(form >= 0 && form <= 2, "form must be 0, 1 or 2");
    const NNPlan p = make_plan(B, N, M);
    if (!ws || ws_bytes < p.total || ((uintptr_t)ws & 255)) {
        set_error("pcst_nn_min_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, p.total);
        return PCST_ERR_WORKSPACE;
    }
    char* w = (char*)ws;
    float4* A = (float4*)(w + p.off_a);
    float4* Bp = (float4*)(w + p.off_b);
    unsigned long long* keys = rowarg ? (unsigned long long*)(w + p.off_key) : nullptr;
    int st;
    if ((st = launch_pack(a, B, N, p.Npad, A, stream)) != PCST_OK) return st;
    if ((st = launch_pack(b, B, M, p.Mpad, Bp, stream)) != PCST_OK) return st;
    const size_t n = (size_t)B * N;
    nn_min_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>((unsigned int*)rowmin, keys, n);
    PCST_CUDA(cudaGetLastError());
    const bool vec2 = tuning("nn_min.variant", 2) == 2;
    unsigned int* bits = (unsigned int*)rowmin;
#define PCST_NN_DISPATCH(F)                                                                          \
    if (rowarg) st = launch_main<F, true, false>(p, A, Bp, B, N, bits, keys, stream);                \
    else if (vec2) st = launch_main<F, false, true>(p, A, Bp, B, N, bits, keys, stream);             \
    else st = launch_main<F, false, false>(p, A, Bp, B, N, bits, keys, stream);
    if (form == 0) { PCST_NN_DISPATCH(0) }
    else if (form == 1) { PCST_NN_DISPATCH(1) }
    else { PCST_NN_DISPATCH(2) }
#undef PCST_NN_DISPATCH
    if (st != PCST_OK) return st;
    if (rowarg || form != 0) {
        nn_min_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(rowmin, rowarg, keys, n, form != 0);
        PCST_CUDA(cudaGetLastError());
    }
    return PCST_OK;
}

// ---- backward of the Chamfer loss (models/losses.py:24-61 under autograd) ---------------------
// chamfer[b] = mean_i D(p_i, t_{a(i)}) + mean_j D(t_j, p_{c(j)}),  dD/dp = 2 (p - t), dD/dt = -2 (p - t).
__global__ void chamfer_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                   const int64_t* __restrict__ arg_pt, const int64_t* __restrict__ arg_tp,
                                   const float* __restrict__ grad_out, int N, int M, float* __restrict__ gp,
                                   float* __restrict__ gt) {
    const int b = blockIdx.y;
    const float go = grad_out[b];
    const float* P = pred + (size_t)b * N * 3;
    const float* T = target + (size_t)b * M * 3;
    float* GP = gp + (size_t)b * N * 3;
    float* GT = gt + (size_t)b * M * 3;
    const int total = N + M;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        if (e < N) {
            const int i = e;
            const int j = (int)arg_pt[(size_t)b * N + i];
            const float s = 2.0f * go / (float)N;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float g = s * (P[3 * i + c] - T[3 * j + c]);
                atomicAdd(&GP[3 * i + c], g);
                atomicAdd(&GT[3 * j + c], -g);
            }
        } else {
            const int j = e - N;
            const int i = (int)arg_tp[(size_t)b * M + j];
            const float s = 2.0f * go / (float)M;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float g = s * (T[3 * j + c] - P[3 * i + c]);
                atomicAdd(&GT[3 * j + c], g);
                atomicAdd(&GP[3 * i + c], -g);
            }
        }
    }
}

extern "C" int pcst_chamfer_bwd_f32(const float* pred, const float* target, const int64_t* arg_pt,
                                    const int64_t* arg_tp, const float* grad_out, int B, int N, int M,
                                    float* grad_pred, float* grad_target, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(pred && target && arg_pt && arg_tp && grad_out && grad_pred && grad_target, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && M > 0, "B, N, M must be positive");
    PCST_CUDA(cudaMemsetAsync(grad_pred, 0, (size_t)B * N * 3 * sizeof(float), stream));
    PCST_CUDA(cudaMemsetAsync(grad_target, 0, (size_t)B * M * 3 * sizeof(float), stream));
    int blocks = (N + M + 255) / 256;
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    chamfer_bwd_kernel<<<dim3(blocks, B), 256, 0, stream>>>(pred, target, arg_pt, arg_tp, grad_out, N, M, grad_pred,
                                                             grad_target);
    return check_cuda(cudaGetLastError(), "chamfer_bwd_kernel");
}
