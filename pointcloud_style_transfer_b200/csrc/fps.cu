// fps.cu -- farthest point sampling, one persistent thread-block cluster per cloud (sm_100a).
//
// Replaces the npoint x (gather, sub/pow/sum, <, masked index_put, max) launch train of
// models/pointnet2_encoder.py:38-44 (about 7 launches and one host sync per iteration) with a
// single launch per batch.  The cloud is split over the CTAs of a cluster (up to 16, the
// non-portable maximum); every thread keeps P points (x, y, z) and their running minimum
// distance in registers for the whole kernel, so HBM is touched once (N*12 B in, npoint*8 B out).
//
// Work layout: the cloud is cut into chunks of 32*P consecutive points; chunk c belongs to warp
// (c / C) of CTA (c % C), so neighbouring chunks sit on different SMs.  Per iteration and warp:
//   1. exact skip test: if the chunk's bounding box is at least as far from the new centroid as the
//      chunk's current maximum running distance (lower bound evaluated with the same fp32 operation
//      sequence as the distances, which is monotone), no distance in the chunk can change and the
//      warp re-uses its cached (max, argmax);  clouds whose index order is spatially coherent (LiDAR
//      scans) skip most chunks after the first few dozen samples;
//   2. otherwise packed fp32x2 distance update + per-thread argmax + REDUX warp argmax;
//   3. shared-memory block argmax (one __syncthreads), then a DSMEM all-to-all of
//      (value, index, xyz) by st.async with completion counted on a per-CTA mbarrier
//      (no barrier.cluster / MEMBAR.GPU in the loop).
//
// Arithmetic (bit-exact with the reference's fp32 CPU path):
//   d = ((dx*dx) + (dy*dy)) + (dz*dz), no FMA;  dist = min(dist, d), dist0 = 1e10;
//   next = argmax(dist), lowest index among equal maxima.
// Running distances are >= +0, so their IEEE bit patterns order like signed integers; padding
// slots hold -1.0f and never win.  NaN / Inf coordinates are outside the contract.
#include "common.cuh"

namespace pcst {

constexpr int kFpsMaxThreads = 512;
constexpr int kFpsMaxWarps = kFpsMaxThreads / 32;
constexpr int kFpsMaxCluster = 16;
constexpr unsigned kNoIdx = 0xffffffffu;

struct FpsShared {
    int2 wslot[2][kFpsMaxWarps];       // per-warp (value bits, index), double-buffered by iteration parity
    float4 cslot[2][kFpsMaxCluster];   // per-CTA (x, y, z, value) of the CTA's best point
    unsigned cidx[2][kFpsMaxCluster];  // per-CTA index of the CTA's best point
    uint64_t cbar[2];                  // transaction barriers: C x 20 bytes land per use
};

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// one-sided gap between a coordinate interval [lo, hi] and c, rounded exactly like (x - c) is
__device__ __forceinline__ float box_gap(float lo, float hi, float c) {
    return fmaxf(fmaxf(__fsub_rn(lo, c), __fsub_rn(c, hi)), 0.0f);
}

// P > 0: register-resident, P points per thread (P even), blockDim.x in {32, 128, 512}.
// P == 0: streaming fallback for clouds that do not fit the register file of one cluster: distances
// live in a global workspace and the points are re-read (from L2) every iteration.
// PROBE = true compiles the latency probes selected by `prune` >= 3 (fps.prune tuning knob, invalid results); the
// production instantiation does not carry their predicates on its per-iteration path.
template <int P, bool PROBE>
__global__ void __launch_bounds__(kFpsMaxThreads, 1)
fps_kernel(const float* __restrict__ xyz, int N, int npoint, const int64_t* __restrict__ start,
           int64_t* __restrict__ out, float* __restrict__ new_xyz, int pts_per_cta, float* __restrict__ dist_ws,
           int prune, int log2c) {
    extern __shared__ __align__(16) float4 smem_pts[];  // P > 0: (x, y, z, -) copy of this CTA's points, one LDS.128 per look-up
    __shared__ FpsShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const unsigned C = cluster_nctarank();
    const unsigned rank = cluster_ctarank();
    const int b = blockIdx.x >> log2c;  // the cluster size is a power of two
    const float* pts = xyz + (size_t)b * N * 3;

    // P > 0: chunk of this warp;  P == 0: contiguous range of this CTA
    constexpr int CH = P > 0 ? 32 * P : 1;
    const int chunk = warp * (int)C + (int)rank;
    const int cbase = P > 0 ? chunk * CH : (int)rank * pts_per_cta;
    const int sbase = warp * CH;  // slot of the chunk inside this CTA's shared-memory copy
    int count = N - (P > 0 ? 0 : cbase);
    if (P == 0) {
        if (count > pts_per_cta) count = pts_per_cta;
        if (count < 0) count = 0;
    }
    float* gdist = (P == 0) ? dist_ws + (size_t)b * N + cbase : nullptr;

    constexpr int PP = P > 0 ? P / 2 : 1;
    float2 px[PP], py[PP], pz[PP], pd[PP];
    float lox = 0.f, loy = 0.f, loz = 0.f, hix = 0.f, hiy = 0.f, hiz = 0.f;
    int cmax = (int)0xbf800000;  // cached warp max (bits of -1.0f: "no valid point")
    unsigned cidx = kNoIdx;     // cached warp argmax
    if (P > 0) {
        const float inf = __int_as_float(0x7f800000);
        float mnx = inf, mny = inf, mnz = inf, mxx = -inf, mxy = -inf, mxz = -inf;
#pragma unroll
        for (int k = 0; k < PP; ++k) {
            float v[2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int l = (2 * k + h) * 32 + lane;
                if (cbase + l < N) {
                    const float* q = pts + (size_t)(cbase + l) * 3;
                    v[h][0] = q[0]; v[h][1] = q[1]; v[h][2] = q[2]; v[h][3] = 1e10f;
                    smem_pts[sbase + l] = make_float4(v[h][0], v[h][1], v[h][2], 0.f);
                    mnx = fminf(mnx, v[h][0]); mxx = fmaxf(mxx, v[h][0]);
                    mny = fminf(mny, v[h][1]); mxy = fmaxf(mxy, v[h][1]);
                    mnz = fminf(mnz, v[h][2]); mxz = fmaxf(mxz, v[h][2]);
                } else {
                    v[h][0] = v[h][1] = v[h][2] = 0.f; v[h][3] = -1.0f;
                }
            }
            px[k] = make_float2(v[0][0], v[1][0]);
            py[k] = make_float2(v[0][1], v[1][1]);
            pz[k] = make_float2(v[0][2], v[1][2]);
            pd[k] = make_float2(v[0][3], v[1][3]);
        }
        lox = warp_min_f(mnx); loy = warp_min_f(mny); loz = warp_min_f(mnz);
        hix = warp_max_f(mxx); hiy = warp_max_f(mxy); hiz = warp_max_f(mxz);
        if (cbase < N) {  // non-empty chunk: every running distance starts at 1e10, lowest index first
            cmax = __float_as_int(1e10f);
            cidx = (unsigned)cbase;
        }
    } else {
        for (int l = tid; l < count; l += nthreads) gdist[l] = 1e10f;
    }

    long long far = start[b];
    if (far < 0) far = 0;
    if (far >= N) far = N - 1;
    float cx = pts[3 * far], cy = pts[3 * far + 1], cz = pts[3 * far + 2];
    if (tid == 0) {
        mbar_init(&sh.cbar[0], 1);
        mbar_init(&sh.cbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (C > 1) cluster_sync_all();  // every CTA is resident and its barriers are initialised before DSMEM traffic

    for (int it = 0; it < npoint; ++it) {
        if (rank == 0 && tid == 0) {
            out[(size_t)b * npoint + it] = far;
            if (new_xyz) {
                float* o = new_xyz + ((size_t)b * npoint + it) * 3;
                o[0] = cx; o[1] = cy; o[2] = cz;
            }
        }
        if (it == npoint - 1) break;
        const int par = it & 1;

        int wmax;
        unsigned widx;
        if (P > 0) {
            // ---- exact skip test (warp-uniform) ----
            // (x and y together as one packed fp32x2 sequence; every lane operation is the same RN add / multiply)
            const float2 c2 = make_float2(cx, cy);
            const float2 ga = __fadd2_rn(make_float2(lox, loy), make_float2(-c2.x, -c2.y));
            const float2 gb = __fadd2_rn(c2, make_float2(-hix, -hiy));
            const float2 g2 = make_float2(fmaxf(fmaxf(ga.x, gb.x), 0.0f), fmaxf(fmaxf(ga.y, gb.y), 0.0f));
            const float gz = box_gap(loz, hiz, cz);
            const float2 sq = __fmul2_rn(g2, g2);
            const float lb = __fadd_rn(__fadd_rn(sq.x, sq.y), __fmul_rn(gz, gz));
            bool skip = prune && lb >= __int_as_float(cmax);
            if (PROBE) skip = skip || (prune >= 3 && it > 0);
            if (skip) {
                wmax = cmax;
                widx = cidx;
            } else {
                // ---- distance update + per-thread argmax (ascending index, strict > keeps the lowest) ----
                float best = -1.0f;
                unsigned bi = kNoIdx;
                const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
                for (int k = 0; k < PP; ++k) {
                    const float2 dx = __fadd2_rn(px[k], ncx), dy = __fadd2_rn(py[k], ncy), dz = __fadd2_rn(pz[k], ncz);
                    const float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
                    pd[k].x = fminf(pd[k].x, d.x);
                    pd[k].y = fminf(pd[k].y, d.y);
                    if (pd[k].x > best) { best = pd[k].x; bi = cbase + (2 * k) * 32 + lane; }
                    if (pd[k].y > best) { best = pd[k].y; bi = cbase + (2 * k + 1) * 32 + lane; }
                }
                // ---- warp argmax: two REDUX instead of a 5-step shuffle tree ----
                const int vb = __float_as_int(best);
                wmax = __reduce_max_sync(0xffffffffu, vb);
                widx = __reduce_min_sync(0xffffffffu, vb == wmax ? bi : kNoIdx);
                cmax = wmax;
                cidx = widx;
            }
        } else {
            float best = -1.0f;
            unsigned bi = kNoIdx;
            for (int l = tid; l < count; l += nthreads) {
                const float* q = pts + (size_t)(cbase + l) * 3;
                const float dx = __fsub_rn(q[0], cx), dy = __fsub_rn(q[1], cy), dz = __fsub_rn(q[2], cz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                const float cur = fminf(gdist[l], d);
                gdist[l] = cur;
                if (cur > best) { best = cur; bi = cbase + l; }
            }
            const int vb = __float_as_int(best);
            wmax = __reduce_max_sync(0xffffffffu, vb);
            widx = __reduce_min_sync(0xffffffffu, vb == wmax ? bi : kNoIdx);
        }

        // ---- block argmax, computed redundantly by every warp ----
        int bmax = wmax;
        unsigned bidx = widx;
        if (nwarps > 1 && !(PROBE && prune == 4)) {
            if (lane == 0) sh.wslot[par][warp] = make_int2(wmax, (int)widx);
            __syncthreads();
            int2 e = lane < nwarps ? sh.wslot[par][lane] : make_int2((int)0x80000000, (int)kNoIdx);
            bmax = __reduce_max_sync(0xffffffffu, e.x);
            bidx = __reduce_min_sync(0xffffffffu, e.x == bmax ? (unsigned)e.y : kNoIdx);
        }
        // shared-memory slot of the block's best point (P > 0)
        const int bslot = P > 0 ? (int)((bidx / CH) >> log2c) * CH + (int)(bidx % CH) : 0;

        if (C == 1 || (PROBE && prune == 5)) {
            far = bidx;
            if (P > 0) {
                const float4 q4 = smem_pts[bslot];
                cx = q4.x; cy = q4.y; cz = q4.z;
            } else {
                const float* q = pts + (size_t)bidx * 3;
                cx = q[0]; cy = q[1]; cz = q[2];
            }
        } else {
            // ---- cluster argmax: each CTA pushes its best (xyz, value, index) into every CTA's slot with
            // st.async; the 20 bytes per sender complete a transaction count on the receiver's mbarrier.
            // Buffer `par` is reused every second iteration; its k-th use is phase k of cbar[par].
            if (tid == 0) mbar_arrive_expect_tx(&sh.cbar[par], 20u * C);
            if (warp == 0 && lane < (int)C) {
                float x = 0.f, y = 0.f, z = 0.f;
                if (bidx != kNoIdx) {
                    if (P > 0) {
                        const float4 q4 = smem_pts[bslot];
                        x = q4.x; y = q4.y; z = q4.z;
                    } else {
                        const float* q = pts + (size_t)bidx * 3;
                        x = q[0]; y = q[1]; z = q[2];
                    }
                }
                const uint32_t rbar = mapa_shared(smem_u32(&sh.cbar[par]), lane);
                asm volatile(
                    "st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                        mapa_shared(smem_u32(&sh.cslot[par][rank]), lane)),
                    "f"(x), "f"(y), "f"(z), "f"(__int_as_float(bmax)), "r"(rbar)
                    : "memory");
                asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(
                                 mapa_shared(smem_u32(&sh.cidx[par][rank]), lane)),
                             "r"(bidx), "r"(rbar)
                             : "memory");
            }
            mbar_wait(&sh.cbar[par], (uint32_t)((it >> 1) & 1));
            float4 a = make_float4(0.f, 0.f, 0.f, -1.0f);
            unsigned ii = kNoIdx;
            if (lane < (int)C) {
                a = sh.cslot[par][lane];
                ii = sh.cidx[par][lane];
            }
            const int gb = __float_as_int(a.w);
            const int gmax = __reduce_max_sync(0xffffffffu, gb);
            const unsigned gidx = __reduce_min_sync(0xffffffffu, gb == gmax ? ii : kNoIdx);
            if (P > 0) {
                // the winner's coordinates sit in the slot of the CTA that owns its chunk (chunk mod C): one
                // broadcast LDS instead of a ballot and three shuffles
                const float4 wp = sh.cslot[par][(gidx / CH) & (C - 1u)];
                cx = wp.x; cy = wp.y; cz = wp.z;
            } else {
                const unsigned hit = __ballot_sync(0xffffffffu, ii == gidx);
                const int wl = __ffs(hit) - 1;
                cx = __shfl_sync(0xffffffffu, a.x, wl);
                cy = __shfl_sync(0xffffffffu, a.y, wl);
                cz = __shfl_sync(0xffffffffu, a.z, wl);
            }
            far = gidx;
        }
    }
    // Every CTA has received all C messages of the last exchange before it gets here, so no peer
    // writes into an exited CTA's shared memory.
}

struct FpsPlan {
    int C, P, threads, pts_per_cta;
    size_t smem, ws;
};

static FpsPlan fps_plan(int B, int N) {
    FpsPlan p;
    int C = tuning("fps.cluster", 0);
    if (C == 0) C = N <= 8192 ? 1 : (N <= 32768 ? 8 : 16);
    if (C > kFpsMaxCluster) C = kFpsMaxCluster;
    while (C & (C - 1)) C &= C - 1;  // power of two (the kernel shifts by log2 C)
    if (C < 1) C = 1;
    // a batch that would need more than one wave of clusters: an iteration costs about the same for any cluster
    // size (it is the exchange chain), so prefer smaller clusters as long as the cloud still fits their registers
    if (tuning("fps.cluster", 0) == 0)
        while (C > 1 && (long)B * C > num_sms() && (long)N <= (long)(C / 2) * kFpsMaxThreads * 16) C /= 2;
    p.C = C;
    int threads = tuning("fps.threads", 0);
    // measured (B200, 512 -> 128): 128 threads 63 us, 32 threads 72 us, 512 threads 65 us per launch
    if (threads == 0) threads = (C == 1 && N <= 64) ? 32 : ((C == 1 && N <= 2048) ? 128 : kFpsMaxThreads);
    if (threads != 32 && threads != 64 && threads != 128 && threads != 256) threads = kFpsMaxThreads;
    p.threads = threads;
    const long lanes = (long)threads * C;                 // threads that share one cloud
    const long per_thread = (N + lanes - 1) / lanes;
    p.P = per_thread <= 2 ? 2 : per_thread <= 4 ? 4 : per_thread <= 8 ? 8 : per_thread <= 16 ? 16 : 0;
    p.pts_per_cta = p.P > 0 ? threads * p.P : (int)align_up((size_t)((N + C - 1) / C), 32);
    p.smem = p.P > 0 ? (size_t)p.pts_per_cta * sizeof(float4) : 0;
    p.ws = p.P > 0 ? 0 : align_up((size_t)B * N * sizeof(float), 256);
    return p;
}

template <int P>
static int fps_launch(const FpsPlan& p, const float* xyz, int B, int N, int npoint, const int64_t* start,
                      int64_t* out, float* new_xyz, float* dist_ws, cudaStream_t stream) {
    int prune = tuning("fps.prune", 1);
    prune = prune == 2 ? 0 : prune;
    auto kern = prune >= 3 ? fps_kernel<P, true> : fps_kernel<P, false>;
    if (p.smem > 32 * 1024)  // dynamic + static (FpsShared) must stay under the opt-in limit, not the 48 KiB default
        PCST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    if (p.C > 8) PCST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * p.C);
    cfg.blockDim = dim3(p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int pts_per_cta = p.pts_per_cta;
    // 2 = off (for A/B measurements); 3 = skip every chunk after the first iteration (WRONG results:
    // measures the latency floor of the exchange chain alone); 4 = 3 without the block-level reduction;
    // 5 = 3 without the cluster exchange (latency probes, results invalid)
    int log2c = 0;
    while ((1 << log2c) < p.C) ++log2c;
    PCST_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, npoint, start, out, new_xyz, pts_per_cta, dist_ws, prune, log2c));
    return PCST_OK;
}

}  // namespace pcst

using namespace pcst;

extern "C" size_t pcst_fps_workspace_bytes(int B, int N, int npoint) {
    (void)npoint;
    if (B <= 0 || N <= 0) return 0;
    return fps_plan(B, N).ws;
}

// How many clouds of N points the device runs CONCURRENTLY (one cluster each): a batch larger than this
// takes a second wave of clusters.  cudaOccupancyMaxActiveClusters accounts for the GPC-local placement of the
// 16-CTA clusters (148 SMs do not hold nine, and not every GPC holds one).
template <int P>
static int fps_max_clusters(const FpsPlan& p) {
    auto kern = fps_kernel<P, false>;
    if (p.smem > 32 * 1024 &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem) != cudaSuccess)
        return 0;
    if (p.C > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.C);
    cfg.blockDim = dim3(p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) return 0;
    return n;
}

extern "C" int pcst_fps_max_concurrent_clouds(int N) {
    if (N <= 0) return 0;
    const FpsPlan p = fps_plan(1, N);
    switch (p.P) {
        case 2: return fps_max_clusters<2>(p);
        case 4: return fps_max_clusters<4>(p);
        case 8: return fps_max_clusters<8>(p);
        case 16: return fps_max_clusters<16>(p);
        default: return fps_max_clusters<0>(p);
    }
}

extern "C" int pcst_fps_f32(const float* xyz, int B, int N, int npoint, const int64_t* start, int64_t* out,
                            float* new_xyz, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(xyz && start && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && npoint > 0, "B, N, npoint must be positive");
    const FpsPlan p = fps_plan(B, N);
    if (p.ws > 0 && (!ws || ws_bytes < p.ws || ((uintptr_t)ws & 255))) {
        set_error("pcst_fps_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, p.ws);
        return PCST_ERR_WORKSPACE;
    }
    float* dist_ws = (float*)ws;
    switch (p.P) {
        case 2: return fps_launch<2>(p, xyz, B, N, npoint, start, out, new_xyz, dist_ws, stream);
        case 4: return fps_launch<4>(p, xyz, B, N, npoint, start, out, new_xyz, dist_ws, stream);
        case 8: return fps_launch<8>(p, xyz, B, N, npoint, start, out, new_xyz, dist_ws, stream);
        case 16: return fps_launch<16>(p, xyz, B, N, npoint, start, out, new_xyz, dist_ws, stream);
        default: return fps_launch<0>(p, xyz, B, N, npoint, start, out, new_xyz, dist_ws, stream);
    }
}
