// voxel.cu -- the deterministic part of HierarchicalProcessor._voxel_grid_downsample_torch
// (models/diffusion_model.py:69-122), the step that turns a 120k-point scan into the 30k-point cloud the
// encoder and the denoiser see, sm_100a.
//
// Reference, per batch element (:78-97): bounding box -> voxel_size (scalar fp32 arithmetic, done by the caller
// exactly as torch does it) -> voxel index floor((p - min) / voxel_size).int() -> int32 hash
// (ix*73856093) ^ (iy*19349663) ^ (iz*83492791) -> torch.unique(sorted) -> per unique voxel the MEAN of the member
// point indices, formed as float32(sum) / float32(count) and truncated (:90-93).  The random top-up / thinning
// (:95-112) draws from torch's CPU generator and stays in the Python wrapper.
//
// Here: one launch for the bounding boxes, one for the hashes, a stable 8-bit LSD radix sort of the (hash, index) pairs
// spread over ceil(N / 4096) CTAs per cloud (4 passes x {count, scatter}; coalesced 32-key rounds per warp, ballot-built
// digit groups for the per-(digit, warp) counters and the in-round ranks) and two launches that reduce each run of
// equal hashes.  Everything is integer work except the index division, which reproduces torch's int64 / int64 ->
// float32 true division.  Bound: launch latency of the 11 small dependent launches (N*8 B of L2-resident keys per
// pass; HBM traffic N*12 B in, U*8 B out).
#include "common.cuh"

namespace pcst {

constexpr int kVoxThreads = 1024;

// ---- per-cloud bounding box: out[b] = (min x, min y, min z, max x, max y, max z) ------------------------------
__global__ void __launch_bounds__(kVoxThreads)
minmax_kernel(const float* __restrict__ xyz, int N, float* __restrict__ out) {
    __shared__ float red[6][kVoxThreads / 32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* p = xyz + (size_t)b * N * 3;
    const float inf = __int_as_float(0x7f800000);
    float v[6] = {inf, inf, inf, -inf, -inf, -inf};
    for (int i = tid; i < N; i += kVoxThreads) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float x = p[3 * i + c];
            v[c] = fminf(v[c], x);
            v[3 + c] = fmaxf(v[3 + c], x);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float u = __shfl_xor_sync(0xffffffffu, v[c], o);
            v[c] = c < 3 ? fminf(v[c], u) : fmaxf(v[c], u);
        }
        if (lane == 0) red[c][warp] = v[c];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            float u = red[c][lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float w = __shfl_xor_sync(0xffffffffu, u, o);
                u = c < 3 ? fminf(u, w) : fmaxf(u, w);
            }
            if (lane == 0) out[(size_t)b * 6 + c] = u;
        }
    }
}

// ---- voxel hash per point: key = hash with the sign bit flipped (unsigned order == torch's signed int32 order) ----
__global__ void vox_hash_kernel(const float* __restrict__ xyz, int N, const float* __restrict__ xyz_min,
                                const float* __restrict__ voxel_size, unsigned int* __restrict__ keys,
                                unsigned int* __restrict__ vals) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float* p = xyz + ((size_t)b * N + i) * 3;
    const float vs = voxel_size[b];
    unsigned int h = 0;
    const unsigned int mult[3] = {73856093u, 19349663u, 83492791u};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // floor((p - min) / voxel_size).int(): fp32 subtract, IEEE fp32 divide, floor, float -> int32
        const float q = floorf(__fdiv_rn(__fsub_rn(p[c], xyz_min[b * 3 + c]), vs));
        const int iq = (int)q;
        h ^= (unsigned int)iq * mult[c];  // int32 multiply wraps in torch; unsigned arithmetic has the same bits
    }
    keys[(size_t)b * N + i] = h ^ 0x80000000u;
    vals[(size_t)b * N + i] = (unsigned int)i;
}

// ---- stable LSD radix sort of (key, val), 8-bit digits, a cloud spread over ceil(N / 4096) CTAs -----------------
// Per pass: vox_count_kernel writes every tile's digit histogram to ghist[b][digit][tile]; vox_scatter_kernel
// turns the table into its own global bases (digit-major exclusive scan: all smaller digits of all tiles, then the
// same digit of the earlier tiles -- 256 threads, a few dozen loads each), ranks the tile's keys stably and
// scatters them.  Inside a tile warp w owns a contiguous 512-key segment and walks it 32 keys at a time; lanes holding
// the same digit find each other with eight ballots (one per digit bit): the group's size feeds the per-(digit,
// warp) counter, a lane's rank inside its group gives its slot, and scanning the counters in (digit, warp) order
// keeps the order of equal digits across warps.
constexpr int kSortTile = 4096;
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = kSortTile / kSortThreads;  // 32-key rounds per warp: 16

__device__ __forceinline__ unsigned int digit_peers8(unsigned int d, bool in) {
    unsigned int peers = __ballot_sync(0xffffffffu, in);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool set = (d >> bit) & 1u;
        const unsigned int bal = __ballot_sync(0xffffffffu, set);
        peers &= set ? bal : ~bal;
    }
    return peers;
}

__global__ void __launch_bounds__(kSortThreads)
vox_count_kernel(const unsigned int* __restrict__ keys, int N, int shift, unsigned int* __restrict__ ghist, int ntiles) {
    __shared__ unsigned int hist[256];
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    hist[tid] = 0;
    __syncthreads();
    const unsigned int* k = keys + (size_t)b * N;
    const int lo = tile * kSortTile, hi = min(lo + kSortTile, N);
    for (int i = lo + tid; i < hi; i += kSortThreads) atomicAdd(&hist[(k[i] >> shift) & 255u], 1u);
    __syncthreads();
    ghist[((size_t)b * 256 + tid) * ntiles + tile] = hist[tid];
}

__global__ void __launch_bounds__(kSortThreads)
vox_scatter_kernel(const unsigned int* __restrict__ kin, const unsigned int* __restrict__ vin,
                   unsigned int* __restrict__ kout, unsigned int* __restrict__ vout, int N, int shift,
                   const unsigned int* __restrict__ ghist, int ntiles) {
    __shared__ unsigned int wcnt[256 * kSortWarps];  // [digit][warp]
    __shared__ unsigned int warp_sums[kSortWarps];
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const size_t off = (size_t)b * N;

    // global base of (digit = tid, this tile)
    unsigned int total = 0, before = 0;
    {
        const unsigned int* g = ghist + ((size_t)b * 256 + tid) * ntiles;
#pragma unroll 8
        for (int c = 0; c < ntiles; ++c) {  // independent loads: keep several in flight
            const unsigned int h = g[c];
            before += c < tile ? h : 0u;
            total += h;
        }
    }
    unsigned int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) warp_sums[warp] = incl;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) wcnt[tid * kSortWarps + w] = 0;
    __syncthreads();
    unsigned int base = incl - total + before;
    for (int w = 0; w < warp; ++w) base += warp_sums[w];

    // the tile's keys stay in registers between the counting and the scattering sweep
    const int wlo = tile * kSortTile + warp * (kSortTile / kSortWarps);
    unsigned int kk[kSortRounds], vv[kSortRounds];
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int i = wlo + r * 32 + lane;
        kk[r] = vv[r] = 0u;
        if (i < N) {
            kk[r] = kin[off + i];
            vv[r] = vin[off + i];
        }
    }
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const bool in = wlo + r * 32 + lane < N;
        const unsigned int d = (kk[r] >> shift) & 255u;
        const unsigned int peers = digit_peers8(d, in);
        if (in && (peers & lt_mask) == 0) wcnt[d * kSortWarps + warp] += __popc(peers);  // leader; warp-private column
        __syncwarp();
    }
    __syncthreads();
    {   // digit = tid: exclusive scan over the warps, starting at the global base
        unsigned int run = base;
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const unsigned int t = wcnt[tid * kSortWarps + w];
            wcnt[tid * kSortWarps + w] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const bool in = wlo + r * 32 + lane < N;
        const unsigned int d = (kk[r] >> shift) & 255u;
        const unsigned int peers = digit_peers8(d, in);
        unsigned int pos = 0;
        if (in) pos = wcnt[d * kSortWarps + warp] + __popc(peers & lt_mask);
        __syncwarp();
        if (in && (peers & lt_mask) == 0) wcnt[d * kSortWarps + warp] += __popc(peers);
        __syncwarp();
        if (in) {
            kout[off + pos] = kk[r];
            vout[off + pos] = vv[r];
        }
    }
}

// ---- runs of equal keys -> one truncated float32 mean of the member indices per run ---------------------------
__global__ void __launch_bounds__(kSortThreads)
vox_heads_kernel(const unsigned int* __restrict__ keys, int N, unsigned int* __restrict__ gcount, int ntiles,
                 unsigned long long* __restrict__ sum, unsigned int* __restrict__ cnt) {
    __shared__ unsigned int wsum[kSortWarps];
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int* k = keys + (size_t)b * N;
    const int lo = tile * kSortTile, hi = min(lo + kSortTile, N);
    unsigned int heads = 0;
    for (int i = lo + tid; i < hi; i += kSortThreads) {
        heads += (i == 0 || k[i] != k[i - 1]) ? 1u : 0u;
        sum[(size_t)b * N + i] = 0ull;   // the per-run accumulators of vox_runs_kernel (run ranks are < N)
        cnt[(size_t)b * N + i] = 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) heads += __shfl_xor_sync(0xffffffffu, heads, o);
    if (lane == 0) wsum[warp] = heads;
    __syncthreads();
    if (tid == 0) {
        unsigned int t = 0;
        for (int w = 0; w < kSortWarps; ++w) t += wsum[w];
        gcount[(size_t)b * ntiles + tile] = t;
    }
}

// Every element adds its index to the accumulator of its run (rank = number of run heads up to and including it, minus
// one): a warp first reduces each stretch of 32 consecutive elements that share a run (prefix sum over the lanes, one
// difference per segment) and the segment's last lane issues ONE atomic for it.  A run of any length therefore costs
// length / 32 atomics spread over its warps -- the first version let the head's thread walk its run alone, and a LiDAR
// scan's near-range voxels hold thousands of points: 307 us of a 120k-point call were that tail.
__global__ void __launch_bounds__(kSortThreads)
vox_runs_kernel(const unsigned int* __restrict__ keys, const unsigned int* __restrict__ vals, int N,
                const unsigned int* __restrict__ gcount, int ntiles, unsigned long long* __restrict__ sum,
                unsigned int* __restrict__ cnt, int* __restrict__ count) {
    __shared__ unsigned int wsum[kSortWarps];
    const int tile = blockIdx.x, b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int le_mask = lane == 31 ? 0xffffffffu : (2u << lane) - 1u;
    const unsigned int* k = keys + (size_t)b * N;
    const unsigned int* v = vals + (size_t)b * N;
    unsigned int before = 0, total = 0;
#pragma unroll 8
    for (int c = 0; c < ntiles; ++c) {
        const unsigned int h = gcount[(size_t)b * ntiles + c];
        before += c < tile ? h : 0u;
        total += h;
    }
    if (tile == 0 && tid == 0) count[b] = (int)total;
    const int wlo = tile * kSortTile + warp * (kSortTile / kSortWarps);
    unsigned int hb[kSortRounds], mine = 0;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int i = wlo + r * 32 + lane;
        const bool head = i < N && (i == 0 || k[i] != k[i - 1]);
        hb[r] = __ballot_sync(0xffffffffu, head);
        mine += __popc(hb[r]);
    }
    if (lane == 0) wsum[warp] = mine;
    __syncthreads();
    unsigned int run = before;
    for (int w = 0; w < warp; ++w) run += wsum[w];
    unsigned long long* osum = sum + (size_t)b * N;
    unsigned int* ocnt = cnt + (size_t)b * N;
#pragma unroll
    for (int r = 0; r < kSortRounds; ++r) {
        const int i = wlo + r * 32 + lane;
        const bool in = i < N;
        unsigned int pre = in ? v[i] : 0u;   // inclusive prefix sum over the lanes (32 indices < 2^27: no overflow)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int t = __shfl_up_sync(0xffffffffu, pre, o);
            if (lane >= o) pre += t;
        }
        const unsigned int mine_heads = hb[r] & le_mask;
        const int seg = mine_heads ? 31 - __clz(mine_heads) : 0;       // first lane of this lane's segment
        const unsigned int below = __shfl_sync(0xffffffffu, pre, seg > 0 ? seg - 1 : 0);
        const bool last = in && (lane == 31 || ((hb[r] >> (lane + 1)) & 1u) || i + 1 >= N);
        if (last) {
            // lanes before the round's first head continue the previous run: rank run - 1 (element 0 is a head, so run >= 1)
            const unsigned int rank = run + __popc(mine_heads) - 1u;
            atomicAdd(osum + rank, (unsigned long long)(pre - (seg > 0 ? below : 0u)));
            atomicAdd(ocnt + rank, (unsigned int)(lane - seg + 1));
        }
        run += __popc(hb[r]);
    }
}

// (sum / bincount).long() with int64 operands: torch true-divides in float32 (:93)
__global__ void vox_finish_kernel(const unsigned long long* __restrict__ sum, const unsigned int* __restrict__ cnt, int N,
                                  const int* __restrict__ count, int64_t* __restrict__ rep) {
    const int b = blockIdx.y;
    const int total = count[b];
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < total; r += gridDim.x * blockDim.x)
        rep[(size_t)b * N + r] =
            (int64_t)__fdiv_rn((float)(long long)sum[(size_t)b * N + r], (float)(long long)cnt[(size_t)b * N + r]);
}

}  // namespace pcst

using namespace pcst;

extern "C" int pcst_minmax_f32(const float* xyz, int B, int N, float* out, pcst_stream_t stream_) {
    PCST_CHECK_ARG(xyz && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0, "B, N must be positive");
    minmax_kernel<<<B, kVoxThreads, 0, (cudaStream_t)stream_>>>(xyz, N, out);
    return check_cuda(cudaGetLastError(), "minmax_kernel");
}

extern "C" size_t pcst_voxel_representatives_workspace_bytes(int B, int N) {
    if (B <= 0 || N <= 0) return 0;
    const size_t ntiles = ((size_t)N + kSortTile - 1) / kSortTile;
    return 5 * align_up((size_t)B * N * sizeof(unsigned int), 256) + align_up((size_t)B * 256 * ntiles * 4, 256) +
           align_up((size_t)B * ntiles * 4, 256);
}

extern "C" int pcst_voxel_representatives_f32(const float* xyz, int B, int N, const float* xyz_min,
                                              const float* voxel_size, int64_t* rep, int* count, void* ws,
                                              size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && xyz_min && voxel_size && rep && count, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && B <= 65535, "B in [1, 65535], N positive");
    const size_t need = pcst_voxel_representatives_workspace_bytes(B, N);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
        set_error("pcst_voxel_representatives_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
        return PCST_ERR_WORKSPACE;
    }
    const int ntiles = (N + kSortTile - 1) / kSortTile;
    const size_t stride = align_up((size_t)B * N * sizeof(unsigned int), 256);
    unsigned int* k0 = (unsigned int*)ws;
    unsigned int* v0 = (unsigned int*)((char*)ws + stride);
    unsigned int* k1 = (unsigned int*)((char*)ws + 2 * stride);
    unsigned int* v1 = (unsigned int*)((char*)ws + 3 * stride);
    unsigned int* cnt = (unsigned int*)((char*)ws + 4 * stride);          // per-run member counts
    unsigned int* ghist = (unsigned int*)((char*)ws + 5 * stride);
    unsigned int* gcount = (unsigned int*)((char*)ghist + align_up((size_t)B * 256 * ntiles * 4, 256));
    vox_hash_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(xyz, N, xyz_min, voxel_size, k0, v0);
    PCST_CUDA(cudaGetLastError());
    const dim3 grid(ntiles, B);
    for (int pass = 0; pass < 4; ++pass) {  // an even number of passes: the sorted data ends in k0 / v0
        vox_count_kernel<<<grid, kSortThreads, 0, stream>>>(k0, N, 8 * pass, ghist, ntiles);
        PCST_CUDA(cudaGetLastError());
        vox_scatter_kernel<<<grid, kSortThreads, 0, stream>>>(k0, v0, k1, v1, N, 8 * pass, ghist, ntiles);
        PCST_CUDA(cudaGetLastError());
        unsigned int* t = k0; k0 = k1; k1 = t;
        t = v0; v0 = v1; v1 = t;
    }
    // the sorted pairs are back in the first two buffers (even number of passes): the other two, contiguous, hold the
    // per-run 64-bit index sums
    unsigned long long* sum = (unsigned long long*)k1;
    vox_heads_kernel<<<grid, kSortThreads, 0, stream>>>(k0, N, gcount, ntiles, sum, cnt);
    PCST_CUDA(cudaGetLastError());
    vox_runs_kernel<<<grid, kSortThreads, 0, stream>>>(k0, v0, N, gcount, ntiles, sum, cnt, count);
    PCST_CUDA(cudaGetLastError());
    vox_finish_kernel<<<dim3((N + 1023) / 1024, B), 256, 0, stream>>>(sum, cnt, N, count, rep);
    return check_cuda(cudaGetLastError(), "vox_finish_kernel");
}
