// gather.cu -- index_points (models/pointnet2_encoder.py:17-28) and the grouping step of
// SetAbstraction.forward (:94-101) as stand-alone HBM-bound gathers (sm_100a).
//
// Bound: HBM bandwidth (random row gather).  Bytes per output row of C floats: C*4 read + C*4
// written + 8 B of index.  One warp (or a slice of one) copies one row with 128-bit accesses when
// the row is 16-byte aligned, so both sides are coalesced within a row.  (The fused
// SetAbstraction kernel in group_mlp.cu performs the same gather in its prologue and never
// writes the grouped tensor; these entry points exist for the reference's public functions.)
#include "common.cuh"

namespace pcst {

__device__ __forceinline__ int clamp_idx(int64_t j, int N) {
    return j < 0 ? 0 : (j >= N ? N - 1 : (int)j);  // torch.clamp(idx, 0, N-1), pointnet2_encoder.py:26
}

// rows: total = B*S output rows; each row copied by `tpr` threads (power of two <= 32)
__global__ void index_points_kernel(const float* __restrict__ points, const int64_t* __restrict__ idx, int N, int C,
                                    int S, long total, int tpr, float* __restrict__ out) {
    const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long row = gt / tpr;
    const int sub = (int)(gt % tpr);
    if (row >= total) return;
    const int b = (int)(row / S);
    const int j = clamp_idx(idx[row], N);
    const float* src = points + ((size_t)b * N + j) * C;
    float* dst = out + (size_t)row * C;
    if ((C & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int c = sub; c < C / 4; c += tpr) d4[c] = __ldg(s4 + c);
    } else {
        for (int c = sub; c < C; c += tpr) dst[c] = __ldg(src + c);
    }
}

__global__ void index_points_bwd_kernel(const float* __restrict__ grad_out, const int64_t* __restrict__ idx, int N,
                                        int C, int S, long total, int tpr, float* __restrict__ grad_points) {
    const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long row = gt / tpr;
    const int sub = (int)(gt % tpr);
    if (row >= total) return;
    const int b = (int)(row / S);
    const int j = clamp_idx(idx[row], N);
    float* dst = grad_points + ((size_t)b * N + j) * C;
    const float* src = grad_out + (size_t)row * C;
    for (int c = sub; c < C; c += tpr) atomicAdd(dst + c, src[c]);
}

// out[b,s,k,:] = cat(xyz[b,idx] - new_xyz[b,s], feats[b,idx])
__global__ void group_kernel(const float* __restrict__ xyz, const float* __restrict__ feats,
                             const float* __restrict__ new_xyz, const int64_t* __restrict__ idx, int N, int S, int K,
                             int D, long total, int tpr, float* __restrict__ out) {
    const long gt = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long row = gt / tpr;  // (b, s, k) flattened
    const int sub = (int)(gt % tpr);
    if (row >= total) return;
    const long bs = row / K;
    const int b = (int)(bs / S);
    const int j = clamp_idx(idx[row], N);
    float* dst = out + (size_t)row * (3 + D);
    if (sub == 0) {
        const float* p = xyz + ((size_t)b * N + j) * 3;
        const float* c = new_xyz + (size_t)bs * 3;
        dst[0] = __fsub_rn(p[0], c[0]);
        dst[1] = __fsub_rn(p[1], c[1]);
        dst[2] = __fsub_rn(p[2], c[2]);
    }
    if (D > 0) {
        const float* f = feats + ((size_t)b * N + j) * D;
        for (int c = sub; c < D; c += tpr) dst[3 + c] = __ldg(f + c);
    }
}

static int pick_tpr(int C) {
    int t = 1;
    while (t < 32 && t * 4 < C) t <<= 1;
    return t;
}

}  // namespace pcst

using namespace pcst;

extern "C" int pcst_index_points_f32(const float* points, const int64_t* idx, int B, int N, int C, int S, float* out,
                                     pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(points && idx && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && C > 0 && S > 0, "B, N, C, S must be positive");
    const long total = (long)B * S;
    const int tpr = pick_tpr(C);
    const long threads = total * tpr;
    index_points_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(points, idx, N, C, S, total, tpr, out);
    return check_cuda(cudaGetLastError(), "index_points_kernel");
}

extern "C" int pcst_index_points_bwd_f32(const float* grad_out, const int64_t* idx, int B, int N, int C, int S,
                                         float* grad_points, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(grad_out && idx && grad_points, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && C > 0 && S > 0, "B, N, C, S must be positive");
    const long total = (long)B * S;
    const int tpr = pick_tpr(C);
    const long threads = total * tpr;
    index_points_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(grad_out, idx, N, C, S, total, tpr,
                                                                                   grad_points);
    return check_cuda(cudaGetLastError(), "index_points_bwd_kernel");
}

extern "C" int pcst_group_f32(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B,
                              int N, int S, int K, int D, float* out, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && new_xyz && idx && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && S > 0 && K > 0 && D >= 0, "bad sizes");
    PCST_CHECK_ARG(D == 0 || feats, "feats is NULL but D > 0");
    const long total = (long)B * S * K;
    const int tpr = pick_tpr(3 + D);
    const long threads = total * tpr;
    group_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(xyz, feats, new_xyz, idx, N, S, K, D, total,
                                                                        tpr, out);
    return check_cuda(cudaGetLastError(), "group_kernel");
}
