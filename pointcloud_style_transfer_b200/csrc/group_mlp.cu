// group_mlp.cu -- SetAbstraction grouping + shared MLP + max-pool, fp32 CUDA-core path (sm_100a).
//
// Replaces index_points x2 + subtract + cat (models/pointnet2_encoder.py:94-99) and
// apply_mlp (:106-112: permute, 3 x relu(BatchNorm2d(Conv2d 1x1)), max over nsample) in eval mode.
// precision 0 (this file): one tiled fp32 SGEMM-style launch per layer; layer 0 gathers its rows
// on the fly (the grouped [B,S,K,3+D] tensor is never written), the last layer reduces the max over
// each group's K rows in its epilogue (atomicMax on the IEEE bits of post-ReLU values, which are
// >= +0).  Intermediate activations (rows x C fp32, a few MB) round-trip through L2.  Stages with few
// rows (the group_all stage: 128 rows per scan) use 32 x 32 tiles so that the launch still covers
// dozens of SMs.
// precision 1 (tcgen05 bf16 tensor-core chain) lives in sa_mlp_tc.cu.
//
// Bound: the dense contraction is tensor-core work; this fp32 path exists as the exact-parity
// reference (rtol 1e-4 against the reference's fp32 CPU result) and is FP32-pipe bound.
#include "common.cuh"

namespace pcst {

constexpr int kKC = 16;   // reduction chunk
constexpr int kMlpThreads = 256;

struct LayerArgs {
    // input: dense X [rows, Cin] (gather == 0) or gathered on the fly (gather == 1)
    const float* X;
    const float* xyz;      // [B,N,3]
    const float* feats;    // [B,N,D] or null
    const float* new_xyz;  // [B,S,3] or null (group_all)
    const int64_t* idx;    // [B,S,K] or null (group_all: row k of batch b is point k)
    int N, S, K, D;
    int gather;
    // layer
    const float* W;      // [Cout, Cin]
    const float* scale;  // [Cout]
    const float* shift;  // [Cout]
    int rows, Cin, Cout;
    // output: dense Y [rows, Cout] (pool == 0) or pooled out [B, Cout, S] via atomicMax (pool == 1)
    float* Y;
    unsigned int* out_bits;
    int pool;
};

__device__ __forceinline__ float load_x(const LayerArgs& a, int row, int ci) {
    if (row >= a.rows || ci >= a.Cin) return 0.f;
    if (!a.gather) return a.X[(size_t)row * a.Cin + ci];
    const int k = row % a.K;
    const int bs = row / a.K;
    const int b = bs / a.S;
    int j;
    if (a.idx) {
        const int64_t jj = a.idx[row];
        j = jj < 0 ? 0 : (jj >= a.N ? a.N - 1 : (int)jj);
    } else {
        j = k;
    }
    if (ci < 3) {
        const float p = a.xyz[((size_t)b * a.N + j) * 3 + ci];
        return a.new_xyz ? __fsub_rn(p, a.new_xyz[(size_t)bs * 3 + ci]) : p;
    }
    return a.feats[((size_t)b * a.N + j) * a.D + (ci - 3)];
}

// T x T output tile per CTA (T = 64: 4 x 4 register micro-tile per thread; T = 32: 2 x 2), 256 threads.
template <int T>
__global__ void __launch_bounds__(kMlpThreads)
mlp_layer_kernel(const LayerArgs a) {
    constexpr int MI = T / 16;  // micro-tile edge
    __shared__ __align__(16) float Xs[kKC][T + 4];
    __shared__ __align__(16) float Ws[kKC][T + 4];
    __shared__ float Ys[T][T + 1];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * T;
    const int col0 = blockIdx.y * T;
    float acc[MI][MI];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < MI; ++j) acc[i][j] = 0.f;

    // loader: T rows x 16 reduction columns of X and of W per chunk, T * 16 / 256 elements per thread
    constexpr int LPT = T * kKC / kMlpThreads;  // 4 (T = 64) or 2 (T = 32)
    const int lr = tid / (kKC / LPT);            // row (X) / channel (W) loaded by this thread
    const int lc = (tid % (kKC / LPT)) * LPT;    // first ci of the LPT it loads
    for (int k0 = 0; k0 < a.Cin; k0 += kKC) {
#pragma unroll
        for (int u = 0; u < LPT; ++u) {
            const int ci = k0 + lc + u;
            Xs[lc + u][lr] = load_x(a, row0 + lr, ci);
            const int co = col0 + lr;
            Ws[lc + u][lr] = (co < a.Cout && ci < a.Cin) ? __ldg(a.W + (size_t)co * a.Cin + ci) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
            float xr[MI], wr[MI];
#pragma unroll
            for (int i = 0; i < MI; ++i) {
                xr[i] = Xs[kk][ty * MI + i];
                wr[i] = Ws[kk][tx * MI + i];
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < MI; ++j) acc[i][j] = __fmaf_rn(xr[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: y = relu(scale * acc + shift)
    float sc[MI], sh[MI];
#pragma unroll
    for (int j = 0; j < MI; ++j) {
        const int co = col0 + tx * MI + j;
        sc[j] = co < a.Cout ? a.scale[co] : 0.f;
        sh[j] = co < a.Cout ? a.shift[co] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int row = row0 + ty * MI + i;
#pragma unroll
        for (int j = 0; j < MI; ++j) {
            const float y = __fmaf_rn(acc[i][j], sc[j], sh[j]);
            const float v = y > 0.f ? y : 0.f;
            if (a.pool) {
                Ys[ty * MI + i][tx * MI + j] = v;
            } else if (row < a.rows && col0 + tx * MI + j < a.Cout) {
                a.Y[(size_t)row * a.Cout + col0 + tx * MI + j] = v;
            }
        }
    }
    if (!a.pool) return;
    __syncthreads();
    // max over the rows of each group segment inside this tile, one (segment, channel) per thread;
    // pooled rows are point-major: out[(b * S + s), co]
    int rend = row0 + T;
    if (rend > a.rows) rend = a.rows;
    const int g0 = row0 / a.K, g1 = (rend - 1) / a.K;
    const int nseg = g1 - g0 + 1;
    for (int item = tid; item < nseg * T; item += kMlpThreads) {
        const int seg = item / T, c = item % T;
        const int co = col0 + c;
        if (co >= a.Cout) continue;
        const int g = g0 + seg;
        int ra = g * a.K, rb = ra + a.K;
        if (ra < row0) ra = row0;
        if (rb > rend) rb = rend;
        float m = 0.f;
        for (int r = ra; r < rb; ++r) m = fmaxf(m, Ys[r - row0][c]);
        atomicMax(a.out_bits + (size_t)g * a.Cout + co, __float_as_uint(m));
    }
}

bool sa_mlp_tc_supported(int D, const int* cout);
int sa_mlp_tc_pick_cluster(long tiles, int D, const int* cout);
size_t sa_mlp_tc_blob_bytes(int D, const int* cout, int C);
int sa_mlp_tc_pack(const pcst_mlp3_t* mlp, int D, int C, void* blob, cudaStream_t stream);
int sa_mlp_tc_run(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                  int K, int D, const int* cout, int C, const void* blob, float* out, void* ws, size_t ws_bytes,
                  cudaStream_t stream);
size_t sa_mlp_tc_workspace_bytes(int B, int N, int S, int K, int D);
int sa_mlp_tc_launches(int B, int N, int S, int K, int D);
void sa_mlp_tc_set_probe(unsigned long long* buf, int tiles);

// fp32 blob: w0 [n0, 3+D] | w1 [n1, n0] | w2 [n2, n1] | scale0 shift0 scale1 shift1 scale2 shift2, each 256-byte aligned
struct F32Blob {
    size_t w[3], scale[3], shift[3], total;
};
static F32Blob f32_blob(int D, const int* cout) {
    F32Blob b;
    size_t off = 0;
    int cin = 3 + D;
    for (int l = 0; l < 3; ++l) {
        b.w[l] = off;
        off += align_up((size_t)cout[l] * cin * sizeof(float), 256);
        cin = cout[l];
    }
    for (int l = 0; l < 3; ++l) {
        b.scale[l] = off;
        off += align_up((size_t)cout[l] * sizeof(float), 256);
        b.shift[l] = off;
        off += align_up((size_t)cout[l] * sizeof(float), 256);
    }
    b.total = off;
    return b;
}

}  // namespace pcst

using namespace pcst;

static int check_cout(const int* cout) {
    if (!cout) return 0;
    for (int l = 0; l < 3; ++l)
        if (cout[l] <= 0 || cout[l] > 1024 || (cout[l] % 32) != 0) return 0;
    return 1;
}
// the tensor-core kernel covers widths up to 512 with the first layer <= 256; anything else runs on the
// fp32 path, which is the more precise of the two
static bool use_tc(int D, const int* cout, int precision) { return precision == 1 && sa_mlp_tc_supported(D, cout); }

extern "C" int pcst_sa_mlp_pick_cluster(int B, int S, int K, int D, const int* cout, int precision) {
    if (B <= 0 || S <= 0 || K <= 0 || D < 0 || !check_cout(cout) || !use_tc(D, cout, precision)) return 1;
    const long tiles = ((long)B * S * K + 127) / 128;
    return sa_mlp_tc_pick_cluster(tiles, D, cout);
}

extern "C" size_t pcst_sa_mlp_packed_bytes(int D, const int* cout, int precision, int cluster) {
    if (D < 0 || !check_cout(cout)) return 0;
    return use_tc(D, cout, precision) ? sa_mlp_tc_blob_bytes(D, cout, cluster) : f32_blob(D, cout).total;
}

extern "C" int pcst_sa_mlp_pack_f32(const pcst_mlp3_t* mlp, int D, int precision, int cluster, void* packed,
                                    size_t packed_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(mlp && packed, "null pointer");
    PCST_CHECK_ARG(D >= 0 && check_cout(mlp->cout), "Cout must be a multiple of 32 in [32, 1024]");
    for (int l = 0; l < 3; ++l) PCST_CHECK_ARG(mlp->w[l] && mlp->scale[l] && mlp->shift[l], "null layer pointer");
    PCST_CHECK_ARG(precision == 0 || precision == 1, "precision must be 0 (fp32) or 1 (bf16 tensor cores)");
    const size_t need_bytes = pcst_sa_mlp_packed_bytes(D, mlp->cout, precision, cluster);
    PCST_CHECK_ARG(need_bytes > 0, "unsupported layer widths / cluster size");
    PCST_CHECK_ARG(packed_bytes >= need_bytes && ((uintptr_t)packed & 255) == 0,
                   "packed buffer too small or not 256-byte aligned");
    if (use_tc(D, mlp->cout, precision)) return sa_mlp_tc_pack(mlp, D, cluster, packed, stream);
    const F32Blob b = f32_blob(D, mlp->cout);
    int cin = 3 + D;
    for (int l = 0; l < 3; ++l) {
        const size_t n = (size_t)mlp->cout[l];
        PCST_CUDA(cudaMemcpyAsync((char*)packed + b.w[l], mlp->w[l], n * cin * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        PCST_CUDA(cudaMemcpyAsync((char*)packed + b.scale[l], mlp->scale[l], n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        PCST_CUDA(cudaMemcpyAsync((char*)packed + b.shift[l], mlp->shift[l], n * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        cin = mlp->cout[l];
    }
    return PCST_OK;
}

extern "C" size_t pcst_sa_mlp_max_workspace_bytes(int B, int N, int S, int K, int D, const int* cout, int precision) {
    if (B <= 0 || N <= 0 || S <= 0 || K <= 0 || D < 0 || !check_cout(cout)) return 0;
    if (use_tc(D, cout, precision)) return sa_mlp_tc_workspace_bytes(B, N, S, K, D);  // activations stay in shared memory /
                                                                                      // TMEM; a bf16 copy of feats at throughput shapes
    const size_t rows = (size_t)B * S * K;
    return align_up(rows * cout[0] * sizeof(float), 256) + align_up(rows * cout[1] * sizeof(float), 256);
}

extern "C" int pcst_sa_mlp_max_kernel_launches(int B, int N, int S, int K, int D, const int* cout, int precision) {
    if (B <= 0 || N <= 0 || S <= 0 || K <= 0 || D < 0 || !check_cout(cout)) return 0;
    return use_tc(D, cout, precision) ? sa_mlp_tc_launches(B, N, S, K, D) : 3;
}

extern "C" void pcst_sa_mlp_set_probe(unsigned long long* stamps, int tiles) { pcst::sa_mlp_tc_set_probe(stamps, tiles); }

int pcst_sa_mlp_max_f32(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx,
                                   int B, int N, int S, int K, int D, const int* cout, int precision, int cluster,
                                   const void* packed, float* out, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && out && packed, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, "bad sizes");
    PCST_CHECK_ARG(D == 0 || feats, "feats is NULL but D > 0");
    PCST_CHECK_ARG(check_cout(cout), "Cout must be a multiple of 32 in [32, 1024]");
    PCST_CHECK_ARG(idx || (S == 1 && K == N && !new_xyz), "idx == NULL means group_all: S = 1, K = N, new_xyz = NULL");
    PCST_CHECK_ARG(!idx || new_xyz, "new_xyz is required with idx");
    PCST_CHECK_ARG(precision == 0 || precision == 1, "precision must be 0 (fp32) or 1 (bf16 tensor cores)");
    PCST_CHECK_ARG(((uintptr_t)packed & 255) == 0, "packed must be 256-byte aligned");
    if (use_tc(D, cout, precision))
        return sa_mlp_tc_run(xyz, feats, new_xyz, idx, B, N, S, K, D, cout, cluster, packed, out, ws, ws_bytes, stream);

    const size_t need = pcst_sa_mlp_max_workspace_bytes(B, N, S, K, D, cout, precision);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
        set_error("pcst_sa_mlp_max_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
        return PCST_ERR_WORKSPACE;
    }
    const size_t rows_sz = (size_t)B * S * K;
    PCST_CHECK_ARG(rows_sz < (1u << 30), "B*S*K too large");
    const int rows = (int)rows_sz;
    const F32Blob blob = f32_blob(D, cout);
    float* act0 = (float*)ws;
    float* act1 = (float*)((char*)ws + align_up(rows_sz * cout[0] * sizeof(float), 256));
    PCST_CUDA(cudaMemsetAsync(out, 0, (size_t)B * S * cout[2] * sizeof(float), stream));

    LayerArgs a = {};
    a.xyz = xyz; a.feats = feats; a.new_xyz = new_xyz; a.idx = idx;
    a.N = N; a.S = S; a.K = K; a.D = D;
    a.rows = rows;
    int cin = 3 + D;
    const float* x_in = nullptr;
    const bool small = rows <= 2048;  // few rows: smaller tiles so that the layer still spreads over many SMs
    for (int l = 0; l < 3; ++l) {
        a.gather = (l == 0);
        a.X = x_in;
        a.W = (const float*)((const char*)packed + blob.w[l]);
        a.scale = (const float*)((const char*)packed + blob.scale[l]);
        a.shift = (const float*)((const char*)packed + blob.shift[l]);
        a.Cin = cin; a.Cout = cout[l];
        a.pool = (l == 2);
        a.Y = l == 0 ? act0 : act1;
        a.out_bits = (unsigned int*)out;
        if (small) {
            dim3 grid((rows + 31) / 32, (a.Cout + 31) / 32);
            mlp_layer_kernel<32><<<grid, kMlpThreads, 0, stream>>>(a);
        } else {
            dim3 grid((rows + 63) / 64, (a.Cout + 63) / 64);
            mlp_layer_kernel<64><<<grid, kMlpThreads, 0, stream>>>(a);
        }
        PCST_CUDA(cudaGetLastError());
        x_in = a.Y;
        cin = a.Cout;
    }
    return PCST_OK;
}
