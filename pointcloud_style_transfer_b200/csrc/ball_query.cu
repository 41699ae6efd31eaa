// ball_query.cu -- query_ball_point (models/pointnet2_encoder.py:47-59) and square_distance (:8-15), sm_100a.
//
// The reference materialises an int64 [B,S,N] index tensor and a fp32 [B,S,N] distance matrix and
// sorts the former along N.  Here the same full S x N sweep is done in two small kernels:
//   1. bq_mask_kernel: CTA = (candidate tile of 1024 points) x (256 queries, two per thread as fp32x2 pairs).  The tile's raw xyz rows
//      (12 KiB) are staged in shared memory by one 1-D TMA bulk copy, repacked there to float4
//      (x, y, z, |p|^2) and broadcast to all threads; each thread owns one query and evaluates the
//      predicate NOT(D > r^2) in the reference's exact fp32 rounding (D = ((-2*dot) + |q|^2) + |p|^2,
//      dot as an FMA chain), packing 32 candidates per 32-bit word: a [B,S,N/32] bit matrix (1 bit
//      per pair instead of the reference's 12 bytes).
//   2. bq_emit_kernel: one CTA per query reads its bit row with wide independent loads, scans the hit
//      counts (popc + warp scan + 8-entry block scan) and emits the first nsample set bits in index
//      order, pads short rows with the first hit / N.
// Rows are identical to sort-and-slice because set bits are visited in ascending index order.
// Clouds of at most 2048 points (the second set-abstraction stage) take a single fused launch instead
// (bq_small_kernel).  Bound: FP32 CUDA cores (0.49 GFLOP per 512 x 120k sweep, ~9 issue slots per
// pair); HBM traffic: N*12 B of points + S*N/8 B of bits, all L2-resident.
#include "common.cuh"

namespace pcst {

constexpr int kBQThreads = 128;  // queries per CTA in the mask kernel

// Stage the raw [n,3] fp32 rows of one candidate tile in shared memory and repack them to float4
// (x, y, z, |p|^2) with the reference's norm rounding; slots n..kTilePoints-1 become +inf sentinels.
// Full, 16-byte aligned tiles arrive by one 1-D TMA bulk copy; a ragged or misaligned tile is read
// with coalesced loads.  Called by all threads of the CTA; ends with a __syncthreads.
__device__ __forceinline__ void bq_stage_tile(const float* __restrict__ src, int n, float* raw, float4* tile,
                                              uint64_t* bar, uint32_t parity) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const bool bulk = (n == kTilePoints) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    if (bulk) {
        if (tid == 0) {
            mbar_arrive_expect_tx(bar, kTilePoints * 12);
            tma_load_1d(raw, src, kTilePoints * 12, bar);
        }
        mbar_wait(bar, parity);
    } else {
        for (int e = tid; e < n * 3; e += nthreads) raw[e] = src[e];
        __syncthreads();
    }
    for (int i = tid; i < kTilePoints; i += nthreads) {
        float4 v = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
        if (i < n) {
            v.x = raw[3 * i]; v.y = raw[3 * i + 1]; v.z = raw[3 * i + 2];
            v.w = norm3_sq(v.x, v.y, v.z);
        }
        tile[i] = v;
    }
    __syncthreads();
}

// NOT(D > r^2) with D = ((-2 * dot) + |q|^2) + |p|^2 in the reference's rounding (pointnet2_encoder.py:12-14,54)
__device__ __forceinline__ bool bq_inside(float qx, float qy, float qz, float qn, const float4& p, float radius_sq) {
    float d = -2.0f * dot3_chain(qx, qy, qz, p.x, p.y, p.z);  // exact scaling
    d = __fadd_rn(d, qn);
    d = __fadd_rn(d, p.w);  // sentinel rows have |p|^2 = +inf -> never inside
    return !(d > radius_sq);
}

// Each thread owns TWO queries (s and s + 128) held as fp32x2 register pairs, so one packed FMUL2 / FFMA2 /
// FADD2 evaluates the candidate against both (the candidate's components are the broadcast scalar operand);
// 32 candidates fill one mask word per query, four words go out as one 128-bit store.
constexpr int kBQQueriesPerCta = 2 * kBQThreads;

__global__ void __launch_bounds__(kBQThreads)
bq_mask_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int Npad, int S,
               float radius_sq, unsigned int* __restrict__ mask) {
    __shared__ __align__(128) float4 tile[kTilePoints];
    __shared__ __align__(128) float raw[kTilePoints * 3];
    __shared__ __align__(8) uint64_t full_bar;

    const int tid = threadIdx.x;
    const int t = blockIdx.x;  // candidate tile
    const int b = blockIdx.z;
    const int s0 = blockIdx.y * kBQQueriesPerCta + tid, s1 = s0 + kBQThreads;
    const int words_per_row = Npad / 32;

    if (tid == 0) {
        mbar_init(&full_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    float2 qx = make_float2(0.f, 0.f), qy = qx, qz = qx;
    if (s0 < S) {
        const float* q = new_xyz + ((size_t)b * S + s0) * 3;
        qx.x = q[0]; qy.x = q[1]; qz.x = q[2];
    }
    if (s1 < S) {
        const float* q = new_xyz + ((size_t)b * S + s1) * 3;
        qx.y = q[0]; qy.y = q[1]; qz.y = q[2];
    }
    const float2 qn = make_float2(norm3_sq(qx.x, qy.x, qz.x), norm3_sq(qx.y, qy.y, qz.y));
    int n = N - t * kTilePoints;
    if (n > kTilePoints) n = kTilePoints;
    bq_stage_tile(xyz + ((size_t)b * N + (size_t)t * kTilePoints) * 3, n, raw, tile, &full_bar, 0);

    unsigned int* row0 = mask + ((size_t)b * S + s0) * words_per_row + (size_t)t * 32;
    unsigned int* row1 = mask + ((size_t)b * S + s1) * words_per_row + (size_t)t * 32;
#pragma unroll 1
    for (int c4 = 0; c4 < kTilePoints / 128; ++c4) {
        unsigned int w0[4], w1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            unsigned int a0 = 0, a1 = 0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float4 p = tile[c4 * 128 + k * 32 + i];  // same address for every thread: broadcast
                // D = ((-2 * dot) + |q|^2) + |p|^2 with dot = fma(qz,pz, fma(qy,py, qx*px)) (pointnet2_encoder.py:12-14)
                const float2 px = make_float2(p.x, p.x), py = make_float2(p.y, p.y), pz = make_float2(p.z, p.z);
                float2 d = __ffma2_rn(qz, pz, __ffma2_rn(qy, py, __fmul2_rn(qx, px)));
                d = __fmul2_rn(d, make_float2(-2.0f, -2.0f));  // exact scaling
                d = __fadd2_rn(d, qn);
                d = __fadd2_rn(d, make_float2(p.w, p.w));      // sentinel rows have |p|^2 = +inf -> never inside
                a0 |= (d.x > radius_sq ? 0u : 1u) << i;
                a1 |= (d.y > radius_sq ? 0u : 1u) << i;
            }
            w0[k] = a0;
            w1[k] = a1;
        }
        if (s0 < S) *reinterpret_cast<uint4*>(row0 + c4 * 4) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        if (s1 < S) *reinterpret_cast<uint4*>(row1 + c4 * 4) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    }
}

// Small clouds (N <= kBQSmallMax: the second set-abstraction stage, 512 candidates): one launch, no
// workspace.  The CTA stages the whole cloud in shared memory; each warp owns one query and walks the
// candidates in index order, 32 per step (predicate -> ballot -> ordered emit), stopping after nsample hits.
constexpr int kBQSmallMax = 2 * kTilePoints;

__global__ void __launch_bounds__(kBQThreads)
bq_small_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int N, int S, float radius_sq,
                int nsample, int64_t* __restrict__ out) {
    __shared__ __align__(128) float4 tile[kBQSmallMax];
    __shared__ __align__(128) float raw[kTilePoints * 3];
    __shared__ __align__(8) uint64_t full_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    if (tid == 0) {
        mbar_init(&full_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int ntiles = (N + kTilePoints - 1) / kTilePoints;
    for (int t = 0; t < ntiles; ++t) {
        int n = N - t * kTilePoints;
        if (n > kTilePoints) n = kTilePoints;
        bq_stage_tile(xyz + ((size_t)b * N + (size_t)t * kTilePoints) * 3, n, raw, tile + t * kTilePoints, &full_bar,
                      (uint32_t)(t & 1));
    }
    const int s = blockIdx.x * (kBQThreads / 32) + warp;
    if (s >= S) return;  // whole warp
    const float* q = new_xyz + ((size_t)b * S + s) * 3;
    const float qx = q[0], qy = q[1], qz = q[2];
    const float qn = norm3_sq(qx, qy, qz);
    int64_t* row = out + ((size_t)b * S + s) * nsample;
    int cnt = 0, first = N;
    for (int base = 0; base < N && cnt < nsample; base += 32) {
        const bool in = bq_inside(qx, qy, qz, qn, tile[base + lane], radius_sq);  // slots >= N are sentinels
        const unsigned int hits = __ballot_sync(0xffffffffu, in);
        if (hits == 0u) continue;
        if (cnt == 0) first = base + __ffs(hits) - 1;
        const int pos = cnt + __popc(hits & ((1u << lane) - 1u));
        if (in && pos < nsample) row[pos] = base + lane;
        cnt += __popc(hits);
    }
    for (int k = cnt + lane; k < nsample; k += 32) row[k] = first;  // pad: first hit, or N for an empty ball (:56-58)
}

// One CTA of 8 warps per query.  The bit row is cut into rounds of 8 x 512 words; in a round every lane
// loads its 16 consecutive words with four independent 128-bit loads (one memory round trip per round, the
// whole 120k-candidate row is a single round), counts its hits, and a warp scan + an 8-entry shared-memory
// scan give each lane the output position of its first hit, so hits are emitted in ascending index order.
constexpr int kEmitWarps = 8;
constexpr int kEmitWordsPerLane = 16;
constexpr int kEmitRound = kEmitWarps * 32 * kEmitWordsPerLane;  // words per round

__global__ void __launch_bounds__(kEmitWarps * 32)
bq_emit_kernel(const unsigned int* __restrict__ mask, int N, int Npad, int S, int nsample, int64_t* __restrict__ out) {
    __shared__ int wcount[kEmitWarps];
    __shared__ int wfirst[kEmitWarps];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x, b = blockIdx.y;
    const int words_per_row = Npad / 32;  // a multiple of 32: rows are 128-byte aligned
    const unsigned int* row_bits = mask + ((size_t)b * S + s) * words_per_row;
    int64_t* row = out + ((size_t)b * S + s) * nsample;
    constexpr int kNone = 0x7fffffff;
    int cnt = 0;     // hits emitted or skipped so far (CTA-uniform)
    int first = N;   // index of the row's first hit; N = none yet (pointnet2_encoder.py:56-58)
    for (int base = 0; base < words_per_row && cnt < nsample; base += kEmitRound) {
        const int w0 = base + warp * (32 * kEmitWordsPerLane) + lane * kEmitWordsPerLane;
        unsigned int w[kEmitWordsPerLane];
#pragma unroll
        for (int q = 0; q < kEmitWordsPerLane / 4; ++q) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (w0 + 4 * q < words_per_row) v = __ldg(reinterpret_cast<const uint4*>(row_bits + w0 + 4 * q));
            w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
        int c = 0, myfirst = kNone;
#pragma unroll
        for (int k = kEmitWordsPerLane - 1; k >= 0; --k) {
            c += __popc(w[k]);
            if (w[k] != 0u) myfirst = (w0 + k) * 32 + (__ffs(w[k]) - 1);
        }
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const unsigned int any = __ballot_sync(0xffffffffu, c > 0);
        const int wf = __shfl_sync(0xffffffffu, myfirst, any ? __ffs(any) - 1 : 0);
        if (lane == 31) wcount[warp] = incl;
        if (lane == 0) wfirst[warp] = any ? wf : kNone;
        __syncthreads();
        int before = cnt, total = cnt;
#pragma unroll
        for (int x = 0; x < kEmitWarps; ++x) {
            const int wc = wcount[x];
            if (x < warp) before += wc;
            total += wc;
            if (first == N && wfirst[x] != kNone) first = wfirst[x];
        }
        int pos = before + incl - c;
#pragma unroll
        for (int k = 0; k < kEmitWordsPerLane; ++k) {
            unsigned int bitsk = w[k];
            while (bitsk != 0u && pos < nsample) {
                row[pos++] = (int64_t)(w0 + k) * 32 + (__ffs(bitsk) - 1);
                bitsk &= bitsk - 1u;
            }
        }
        cnt = total;
        __syncthreads();  // wcount / wfirst are rewritten in the next round
    }
    // pad short rows with the first hit; empty rows with N (models/pointnet2_encoder.py:56-58)
    for (int k = cnt + tid; k < nsample; k += kEmitWarps * 32) row[k] = first;
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
    if (N <= kBQSmallMax) return 0;  // single fused launch, no scratch
    const size_t npad = padded_points(N);
    return align_up((size_t)B * S * (npad / 32) * sizeof(unsigned int), 256);
}

extern "C" int pcst_ball_query_f32(const float* xyz, const float* new_xyz, int B, int N, int S, float radius_sq,
                                   int nsample, int64_t* out, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && new_xyz && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && S > 0, "B, N, S must be positive");
    PCST_CHECK_ARG(B <= 65535, "B must be <= 65535");
    PCST_CHECK_ARG(nsample >= 1 && nsample <= N, "nsample must be in [1, N] (the reference raises for nsample > N)");
    if (N <= kBQSmallMax) {
        PCST_CHECK_ARG((S + 3) / 4 <= 0x7fffffff, "S too large");
        bq_small_kernel<<<dim3((S + 3) / 4, B), kBQThreads, 0, stream>>>(xyz, new_xyz, N, S, radius_sq, nsample, out);
        return check_cuda(cudaGetLastError(), "bq_small_kernel");
    }
    const size_t need = pcst_ball_query_workspace_bytes(B, N, S);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
        set_error("pcst_ball_query_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
        return PCST_ERR_WORKSPACE;
    }
    const int Npad = padded_points(N);
    unsigned int* mask = (unsigned int*)ws;
    const int qblocks = (S + kBQQueriesPerCta - 1) / kBQQueriesPerCta;
    PCST_CHECK_ARG(qblocks <= 65535, "S too large");
    bq_mask_kernel<<<dim3(Npad / kTilePoints, qblocks, B), kBQThreads, 0, stream>>>(xyz, new_xyz, N, Npad, S, radius_sq,
                                                                                    mask);
    PCST_CUDA(cudaGetLastError());
    PCST_CHECK_ARG(B <= 65535, "B must be <= 65535");
    bq_emit_kernel<<<dim3(S, B), kEmitWarps * 32, 0, stream>>>(mask, N, Npad, S, nsample, out);
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
