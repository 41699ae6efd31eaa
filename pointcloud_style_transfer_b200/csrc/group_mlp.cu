// group_mlp.cu -- SetAbstraction grouping + shared MLP + max-pool, fp32 CUDA-core path (sm_100a).
//
// Replaces index_points x2 + subtract + cat (models/pointnet2_encoder.py:94-99) and
// apply_mlp (:106-112: permute, 3 x relu(BatchNorm2d(Conv2d 1x1)), max over nsample) in eval mode.
// precision 0 (this file): one tiled fp32 SGEMM-style launch per layer; layer 0 gathers its rows
// on the fly (the grouped [B,S,K,3+D] tensor is never written), the last layer reduces the max over
// each group's K rows in its epilogue (atomicMax on the IEEE bits of post-ReLU values, which are
// >= +0).  Intermediate activations (rows x C fp32, a few MB) round-trip through L2.
// precision 1 (tcgen05 bf16 tensor-core chain) lives in sa_mlp_tc.cu.
//
// Bound: the dense contraction is tensor-core work; this fp32 path exists as the exact-parity
// reference (rtol 1e-4 against the reference's fp32 CPU result) and is FP32-pipe bound.
#include "common.cuh"

namespace pcst {

constexpr int kMT = 64;   // rows per CTA tile
constexpr int kNT = 64;   // output channels per CTA tile
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

__global__ void __launch_bounds__(kMlpThreads)
mlp_layer_kernel(const LayerArgs a) {
    __shared__ __align__(16) float Xs[kKC][kMT + 4];
    __shared__ __align__(16) float Ws[kKC][kNT + 4];
    __shared__ float Ys[kMT][kNT + 1];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.x * kMT;
    const int col0 = blockIdx.y * kNT;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lr = tid >> 2;         // 0..63: row (X) / channel (W) loaded by this thread
    const int lc = (tid & 3) * 4;    // 0,4,8,12: first ci of the 4 it loads
    for (int k0 = 0; k0 < a.Cin; k0 += kKC) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int ci = k0 + lc + u;
            Xs[lc + u][lr] = load_x(a, row0 + lr, ci);
            const int co = col0 + lr;
            Ws[lc + u][lr] = (co < a.Cout && ci < a.Cin) ? __ldg(a.W + (size_t)co * a.Cin + ci) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kKC; ++kk) {
            const float4 xv = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
            const float4 wv = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
            const float xr[4] = {xv.x, xv.y, xv.z, xv.w};
            const float wr[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(xr[i], wr[j], acc[i][j]);
        }
        __syncthreads();
    }

    // epilogue: y = relu(scale * acc + shift)
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int co = col0 + tx * 4 + j;
        sc[j] = co < a.Cout ? a.scale[co] : 0.f;
        sh[j] = co < a.Cout ? a.shift[co] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int row = row0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float y = __fmaf_rn(acc[i][j], sc[j], sh[j]);
            const float v = y > 0.f ? y : 0.f;
            if (a.pool) {
                Ys[ty * 4 + i][tx * 4 + j] = v;
            } else if (row < a.rows && col0 + tx * 4 + j < a.Cout) {
                a.Y[(size_t)row * a.Cout + col0 + tx * 4 + j] = v;
            }
        }
    }
    if (!a.pool) return;
    __syncthreads();
    // max over the rows of each group segment inside this tile, one (segment, channel) per thread
    int rend = row0 + kMT;
    if (rend > a.rows) rend = a.rows;
    const int g0 = row0 / a.K, g1 = (rend - 1) / a.K;
    const int nseg = g1 - g0 + 1;
    for (int item = tid; item < nseg * kNT; item += kMlpThreads) {
        const int seg = item / kNT, c = item % kNT;
        const int co = col0 + c;
        if (co >= a.Cout) continue;
        const int g = g0 + seg;
        int ra = g * a.K, rb = ra + a.K;
        if (ra < row0) ra = row0;
        if (rb > rend) rb = rend;
        float m = 0.f;
        for (int r = ra; r < rb; ++r) m = fmaxf(m, Ys[r - row0][c]);
        const int b = g / a.S, s = g % a.S;
        atomicMax(a.out_bits + ((size_t)b * a.Cout + co) * a.S + s, __float_as_uint(m));
    }
}

int sa_mlp_max_tc(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                  int K, int D, const pcst_mlp3_t* mlp, float* out, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t sa_mlp_max_tc_workspace(int B, int N, int S, int K, int D, const pcst_mlp3_t* mlp);
bool sa_mlp_max_tc_supported(int D, const pcst_mlp3_t* mlp);

}  // namespace pcst

using namespace pcst;

static int check_mlp(const pcst_mlp3_t* mlp) {
    if (!mlp) return 0;
    for (int l = 0; l < 3; ++l) {
        if (!mlp->w[l] || !mlp->scale[l] || !mlp->shift[l]) return 0;
        if (mlp->cout[l] <= 0 || mlp->cout[l] > 1024 || (mlp->cout[l] % 32) != 0) return 0;
    }
    return 1;
}

extern "C" size_t pcst_sa_mlp_max_workspace_bytes(int B, int N, int S, int K, int D, const pcst_mlp3_t* mlp,
                                                  int precision) {
    if (B <= 0 || N <= 0 || S <= 0 || K <= 0 || D < 0 || !check_mlp(mlp)) return 0;
    // stages whose layers do not fit the tensor-core kernel (Cout > 256: the tiny group_all stage) run
    // on the fp32 path, which is the more precise of the two
    if (precision == 1 && sa_mlp_max_tc_supported(D, mlp)) return sa_mlp_max_tc_workspace(B, N, S, K, D, mlp);
    const size_t rows = (size_t)B * S * K;
    return align_up(rows * mlp->cout[0] * sizeof(float), 256) + align_up(rows * mlp->cout[1] * sizeof(float), 256);
}

extern "C" int pcst_sa_mlp_max_f32(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx,
                                   int B, int N, int S, int K, int D, const pcst_mlp3_t* mlp, int precision,
                                   float* out, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, "bad sizes");
    PCST_CHECK_ARG(D == 0 || feats, "feats is NULL but D > 0");
    PCST_CHECK_ARG(check_mlp(mlp), "mlp: null pointers, or Cout not a multiple of 32 in [32, 1024]");
    PCST_CHECK_ARG(idx || (S == 1 && K == N && !new_xyz), "idx == NULL means group_all: S = 1, K = N, new_xyz = NULL");
    PCST_CHECK_ARG(!idx || new_xyz, "new_xyz is required with idx");
    PCST_CHECK_ARG(precision == 0 || precision == 1, "precision must be 0 (fp32) or 1 (bf16 tensor cores)");
    const size_t need = pcst_sa_mlp_max_workspace_bytes(B, N, S, K, D, mlp, precision);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
        set_error("pcst_sa_mlp_max_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
        return PCST_ERR_WORKSPACE;
    }
    if (precision == 1 && sa_mlp_max_tc_supported(D, mlp))
        return sa_mlp_max_tc(xyz, feats, new_xyz, idx, B, N, S, K, D, mlp, out, ws, ws_bytes, stream);

    const size_t rows_sz = (size_t)B * S * K;
    PCST_CHECK_ARG(rows_sz < (1u << 30), "B*S*K too large");
    const int rows = (int)rows_sz;
    float* act0 = (float*)ws;
    float* act1 = (float*)((char*)ws + align_up(rows_sz * mlp->cout[0] * sizeof(float), 256));
    PCST_CUDA(cudaMemsetAsync(out, 0, (size_t)B * mlp->cout[2] * S * sizeof(float), stream));

    LayerArgs a = {};
    a.xyz = xyz; a.feats = feats; a.new_xyz = new_xyz; a.idx = idx;
    a.N = N; a.S = S; a.K = K; a.D = D;
    a.rows = rows;
    int cin = 3 + D;
    const float* x_in = nullptr;
    for (int l = 0; l < 3; ++l) {
        a.gather = (l == 0);
        a.X = x_in;
        a.W = mlp->w[l]; a.scale = mlp->scale[l]; a.shift = mlp->shift[l];
        a.Cin = cin; a.Cout = mlp->cout[l];
        a.pool = (l == 2);
        a.Y = l == 0 ? act0 : act1;
        a.out_bits = (unsigned int*)out;
        dim3 grid((rows + kMT - 1) / kMT, (a.Cout + kNT - 1) / kNT);
        mlp_layer_kernel<<<grid, kMlpThreads, 0, stream>>>(a);
        PCST_CUDA(cudaGetLastError());
        x_in = a.Y;
        cin = a.Cout;
    }
    return PCST_OK;
}
