// ball_query.cu -- query_ball_point (models/pointnet2_encoder.py:47-59) and square_distance (:8-15), sm_100a.
//
// The reference materialises an int64 [B,S,N] index tensor and a fp32 [B,S,N] distance matrix and
// sorts the former along N.  Here the same full S x N sweep is done in two small kernels:
//   1. bq_mask_kernel: CTA = (candidate tile of 1024 points) x (128 queries).  The tile is staged in
//      shared memory as packed float4 by one 1-D TMA bulk copy and broadcast to all threads; each
//      thread owns one query and evaluates the predicate NOT(D > r^2) in the reference's exact fp32
//      rounding (D = ((-2*dot) + |q|^2) + |p|^2, dot as an FMA chain), packing 32 candidates per
//      32-bit word: a [B,S,N/32] bit matrix (1 bit per pair instead of the reference's 12 bytes).
//   2. bq_emit_kernel: one warp per query walks its bit row in index order (popc + warp scan),
//      emits the first nsample set bits, stops early, pads short rows with the first hit / N.
// Rows are identical to sort-and-slice because set bits are visited in ascending index order.
// Bound: FP32 CUDA cores (0.49 GFLOP per 512 x 120k sweep, ~9 issue slots per pair); HBM traffic:
// N*16 B packed points + S*N/8 B of bits, all L2-resident.
#include "common.cuh"

namespace pcst {

constexpr int kBQThreads = 128;  // queries per CTA in the mask kernel

__global__ void __launch_bounds__(kBQThreads)
bq_mask_kernel(const float4* __restrict__ P, const float* __restrict__ new_xyz, int Npad, int S, float radius_sq,
               unsigned int* __restrict__ mask) {
    __shared__ __align__(128) float4 tile[kTilePoints];
    __shared__ unsigned int words[kBQThreads][33];
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t = blockIdx.x;  // candidate tile
    const int b = blockIdx.z;
    const int s = blockIdx.y * kBQThreads + tid;
    const int words_per_row = Npad / 32;

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&full_bar, kTileBytes);
        tma_load_1d(tile, P + (size_t)b * Npad + (size_t)t * kTilePoints, kTileBytes, &full_bar);
    }
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (s < S) {
        const float* q = new_xyz + ((size_t)b * S + s) * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
    }
    const float qn = norm3_sq(qx, qy, qz);
    mbar_wait(&full_bar, 0);

#pragma unroll 1
    for (int c = 0; c < kTilePoints / 32; ++c) {
        unsigned int w = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float4 p = tile[c * 32 + i];  // same address for every thread: shared-memory broadcast
            float d = -2.0f * dot3_chain(qx, qy, qz, p.x, p.y, p.z);  // exact scaling
            d = __fadd_rn(d, qn);
            d = __fadd_rn(d, p.w);  // sentinel rows have |p|^2 = +inf -> never inside
            w |= (d > radius_sq ? 0u : 1u) << i;
        }
        words[tid][c] = w;
    }
    __syncthreads();
    // coalesced write-out: each warp stores 32 query rows of 32 words (128 B) each
    for (int r = warp * 32; r < warp * 32 + 32; ++r) {
        const int sq = blockIdx.y * kBQThreads + r;
        if (sq < S) mask[((size_t)b * S + sq) * words_per_row + (size_t)t * 32 + lane] = words[r][lane];
    }
}

__global__ void __launch_bounds__(128)
bq_emit_kernel(const unsigned int* __restrict__ mask, int N, int Npad, int S, int nsample, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int b = blockIdx.y;
    if (s >= S) return;  // whole warp
    const int words_per_row = Npad / 32;
    const unsigned int* row_bits = mask + ((size_t)b * S + s) * words_per_row;
    int64_t* row = out + ((size_t)b * S + s) * nsample;
    int cnt = 0;
    int first = N;
    for (int base = 0; base < words_per_row && cnt < nsample; base += 32) {
        unsigned int w = base + lane < words_per_row ? row_bits[base + lane] : 0u;
        const unsigned int any = __ballot_sync(0xffffffffu, w != 0u);
        if (any == 0u) continue;
        if (cnt == 0) {
            const int fl = __ffs(any) - 1;
            const unsigned int fw = __shfl_sync(0xffffffffu, w, fl);
            first = (base + fl) * 32 + (__ffs(fw) - 1);
        }
        const int c = __popc(w);
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        int pos = cnt + incl - c;
        while (w != 0u && pos < nsample) {
            const int bit = __ffs(w) - 1;
            row[pos++] = (int64_t)(base + lane) * 32 + bit;
            w &= w - 1u;
        }
        cnt += __shfl_sync(0xffffffffu, incl, 31);
    }
    // pad short rows with the first hit; empty rows with N (models/pointnet2_encoder.py:56-58)
    for (int k = cnt + lane; k < nsample; k += 32) row[k] = first;
}

// square_distance, materialised [B,N,M]: 4 B written per pair -> HBM-write bound.
__global__ void square_distance_kernel(const float* __restrict__ src, const float* __restrict__ dst, int N, int M,
                                       float* __restrict__ out) {
    const int b = blockIdx.z;
    const int i = blockIdx.y;
    const float* s = src + ((size_t)b * N + i) * 3;
    const float sx = s[0], sy = s[1], sz = s[2];
    const float sn = norm3_sq(sx, sy, sz);
    const float* D = dst + (size_t)b * M * 3;
    float* o = out + ((size_t)b * N + i) * M;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < M; j += gridDim.x * blockDim.x) {
        const float dx = D[3 * j], dy = D[3 * j + 1], dz = D[3 * j + 2];
        float d = -2.0f * dot3_chain(sx, sy, sz, dx, dy, dz);
        d = __fadd_rn(d, sn);
        d = __fadd_rn(d, norm3_sq(dx, dy, dz));
        o[j] = d;
    }
}

}  // namespace pcst

using namespace pcst;

extern "C" size_t pcst_ball_query_workspace_bytes(int B, int N, int S) {
    if (B <= 0 || N <= 0 || S <= 0) return 0;
    const size_t npad = padded_points(N);
    return align_up((size_t)B * npad * sizeof(float4), 256) + align_up((size_t)B * S * (npad / 32) * sizeof(unsigned int), 256);
}

extern "C" int pcst_ball_query_f32(const float* xyz, const float* new_xyz, int B, int N, int S, float radius_sq,
                                   int nsample, int64_t* out, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && new_xyz && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && S > 0, "B, N, S must be positive");
    PCST_CHECK_ARG(B <= 65535, "B must be <= 65535");
    PCST_CHECK_ARG(nsample >= 1 && nsample <= N, "nsample must be in [1, N] (the reference raises for nsample > N)");
    const size_t need = pcst_ball_query_workspace_bytes(B, N, S);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
        set_error("pcst_ball_query_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
        return PCST_ERR_WORKSPACE;
    }
    const int Npad = padded_points(N);
    float4* P = (float4*)ws;
    unsigned int* mask = (unsigned int*)((char*)ws + align_up((size_t)B * Npad * sizeof(float4), 256));
    int st = launch_pack(xyz, B, N, Npad, P, stream);
    if (st != PCST_OK) return st;
    const int qblocks = (S + kBQThreads - 1) / kBQThreads;
    PCST_CHECK_ARG(qblocks <= 65535, "S too large");
    bq_mask_kernel<<<dim3(Npad / kTilePoints, qblocks, B), kBQThreads, 0, stream>>>(P, new_xyz, Npad, S, radius_sq, mask);
    PCST_CUDA(cudaGetLastError());
    bq_emit_kernel<<<dim3((S + 3) / 4, B), 128, 0, stream>>>(mask, N, Npad, S, nsample, out);
    return check_cuda(cudaGetLastError(), "bq_emit_kernel");
}

extern "C" int pcst_square_distance_f32(const float* src, const float* dst, int B, int N, int M, float* out,
                                        pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(src && dst && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && M > 0, "B, N, M must be positive");
    PCST_CHECK_ARG(N <= 65535 && B <= 65535, "N and B must be <= 65535 for the materialised matrix");
    int bx = (M + 255) / 256;
    if (bx > 64) bx = 64;
    square_distance_kernel<<<dim3(bx, N, B), 256, 0, stream>>>(src, dst, N, M, out);
    return check_cuda(cudaGetLastError(), "square_distance_kernel");
}
