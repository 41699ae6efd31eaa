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
// Here: one launch for the bounding boxes, one for the hashes, and one CTA per cloud that sorts (hash, index)
// pairs with a stable 4-bit LSD radix sort (8 passes; coalesced 32-key rounds per warp, ballot-built digit groups for the
// per-(digit, warp) histogram and the in-round ranks, block-wide scan) and then reduces each run of equal hashes.
// Everything is integer work except the index division, which reproduces torch's int64 / int64 -> float32 true division.  Bound: latency of the 8 dependent passes over
// N * 8 B of L2-resident keys (HBM traffic N*12 B in, U*8 B out).
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

// lanes of the warp that are in range and hold the same 4-bit digit as this lane: four ballots, one per digit bit
// (MATCH.ANY computes the same mask but is issued at a small fraction of the VOTE rate)
__device__ __forceinline__ unsigned int digit_peers(unsigned int d, bool in) {
    unsigned int peers = __ballot_sync(0xffffffffu, in);
#pragma unroll
    for (int bit = 0; bit < 4; ++bit) {
        const bool set = (d >> bit) & 1u;
        const unsigned int bal = __ballot_sync(0xffffffffu, set);
        peers &= set ? bal : ~bal;
    }
    return peers;
}

// ---- one CTA per cloud: stable LSD radix sort of (key, val), then the per-run index mean -----------------------
// Warp w owns the contiguous segment [w * seg, (w + 1) * seg) and walks it 32 keys at a time (coalesced).  Inside
// a round, lanes holding the same 4-bit digit find each other with four ballots: the group's size feeds the
// per-(digit, warp) histogram, a lane's rank inside its group gives its stable output slot.  The 16 x 32 counters
// are scanned in (digit, warp) order, which makes the scatter stable across warps as well.
__global__ void __launch_bounds__(kVoxThreads)
vox_sort_reduce_kernel(unsigned int* __restrict__ keys0, unsigned int* __restrict__ vals0,
                       unsigned int* __restrict__ keys1, unsigned int* __restrict__ vals1, int N,
                       int64_t* __restrict__ rep, int* __restrict__ count) {
    constexpr int kWarps = kVoxThreads / 32;
    constexpr int kU = 8;  // rounds whose loads are issued together
    __shared__ unsigned int hist[16 * kWarps];  // [digit][warp]
    __shared__ unsigned int warp_sums[kWarps];
    __shared__ unsigned int total_sh;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned int* kin = keys0 + (size_t)b * N;
    unsigned int* vin = vals0 + (size_t)b * N;
    unsigned int* kout = keys1 + (size_t)b * N;
    unsigned int* vout = vals1 + (size_t)b * N;
    const int seg = ((N + kWarps - 1) / kWarps + 31) / 32 * 32;
    const int wlo = min(warp * seg, N), whi = min(wlo + seg, N);

    // block-wide exclusive scan of one value per thread (returns the exclusive prefix; total in total_sh)
    auto block_exscan = [&](unsigned int v) -> unsigned int {
        unsigned int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            unsigned int w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int u = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += u;
            }
            warp_sums[lane] = wi - w;
            if (lane == 31) total_sh = wi;
        }
        __syncthreads();
        const unsigned int r = warp_sums[warp] + incl - v;
        __syncthreads();  // warp_sums / total_sh are reused by the next call
        return r;
    };

    for (int pass = 0; pass < 8; ++pass) {
        const int shift = 4 * pass;
        if (tid < 16 * kWarps) hist[tid] = 0;
        __syncthreads();
        for (int i0 = wlo; i0 < whi; i0 += 32 * kU) {
            unsigned int kk[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {  // kU independent coalesced loads in flight before the dependent part
                const int i = i0 + u * 32 + lane;
                kk[u] = i < whi ? kin[i] : 0u;
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * 32 + lane;
                const bool in = i < whi;
                const unsigned int d = (kk[u] >> shift) & 15u;
                const unsigned int peers = digit_peers(d, in);
                if (in && (peers & lt_mask) == 0) hist[d * kWarps + warp] += __popc(peers);  // group leader; warp-private column
                __syncwarp();  // the next round's leader of the same digit may be another lane
            }
        }
        __syncthreads();
        // exclusive scan of the 512 counters in (digit, warp) order
        const unsigned int mine = tid < 16 * kWarps ? hist[tid] : 0u;
        const unsigned int ex = block_exscan(mine);
        if (tid < 16 * kWarps) hist[tid] = ex;
        __syncthreads();
        for (int i0 = wlo; i0 < whi; i0 += 32 * kU) {
            unsigned int kk[kU], vv[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * 32 + lane;
                kk[u] = vv[u] = 0u;
                if (i < whi) {
                    kk[u] = kin[i];
                    vv[u] = vin[i];
                }
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * 32 + lane;
                const bool in = i < whi;
                const unsigned int d = (kk[u] >> shift) & 15u;
                const unsigned int peers = digit_peers(d, in);
                unsigned int pos = 0;
                if (in) pos = hist[d * kWarps + warp] + __popc(peers & lt_mask);
                __syncwarp();
                if (in && (peers & lt_mask) == 0) hist[d * kWarps + warp] += __popc(peers);
                __syncwarp();
                if (in) {
                    kout[pos] = kk[u];
                    vout[pos] = vv[u];
                }
            }
        }
        __syncthreads();
        unsigned int* t0 = kin; kin = kout; kout = t0;
        unsigned int* t1 = vin; vin = vout; vout = t1;
    }
    // after 8 passes (an even number) the sorted data is back in keys0 / vals0 (= kin / vin)

    // run heads -> run ids (ordered: warp segments, then rounds, then lanes) -> one mean per run
    unsigned int heads = 0;
    for (int i0 = wlo; i0 < whi; i0 += 32) {
        const int i = i0 + lane;
        const bool head = i < whi && (i == 0 || kin[i] != kin[i - 1]);
        heads += __popc(__ballot_sync(0xffffffffu, head));  // every lane counts the whole round
    }
    unsigned int run = block_exscan(lane == 0 ? heads : 0u);  // lane 0 of each warp contributes the warp's count
    run = __shfl_sync(0xffffffffu, run, 0);
    if (tid == 0) count[b] = (int)total_sh;
    int64_t* out = rep + (size_t)b * N;
    for (int i0 = wlo; i0 < whi; i0 += 32) {
        const int i = i0 + lane;
        const bool head = i < whi && (i == 0 || kin[i] != kin[i - 1]);
        const unsigned int hb = __ballot_sync(0xffffffffu, head);
        if (head) {
            long long s = 0, c = 0;
            const unsigned int k = kin[i];
            for (int j = i; j < N && kin[j] == k; ++j) {  // runs are short (a few points per voxel); may cross segments
                s += vin[j];
                ++c;
            }
            // (sum / bincount).long() with int64 operands: torch true-divides in float32 (:93)
            out[run + __popc(hb & lt_mask)] = (int64_t)__fdiv_rn((float)s, (float)c);
        }
        run += __popc(hb);
    }
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
    return 4 * align_up((size_t)B * N * sizeof(unsigned int), 256);
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
    const size_t stride = align_up((size_t)B * N * sizeof(unsigned int), 256);
    unsigned int* k0 = (unsigned int*)ws;
    unsigned int* v0 = (unsigned int*)((char*)ws + stride);
    unsigned int* k1 = (unsigned int*)((char*)ws + 2 * stride);
    unsigned int* v1 = (unsigned int*)((char*)ws + 3 * stride);
    vox_hash_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(xyz, N, xyz_min, voxel_size, k0, v0);
    PCST_CUDA(cudaGetLastError());
    vox_sort_reduce_kernel<<<B, kVoxThreads, 0, stream>>>(k0, v0, k1, v1, N, rep, count);
    return check_cuda(cudaGetLastError(), "vox_sort_reduce_kernel");
}
