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

// 16-byte shared-memory load by 32-bit shared-window address (the address is formed once, outside the sampling loop: a
// generic pointer into the dynamic shared array costs an S2R + address arithmetic at every use on the critical path)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ int2 lds64(uint32_t addr) {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds32(uint32_t addr) {
    unsigned v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, int a, int b, int c, int d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {  // a value the compiler will not re-derive at every use
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}
__device__ __forceinline__ void mbar_arrive_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}

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
    uint32_t pts_sm;  // opaque to the compiler, which would otherwise re-derive the window address (S2R) at every use
    asm volatile("mov.u32 %0, %1;" : "=r"(pts_sm) : "r"(smem_u32(smem_pts)));
    // The sample list is written by ONE thread of the cloud's first CTA.  Not by warp 0 (which drives the block argmax and
    // the cluster exchange, the critical path of an iteration) and not at the top of the iteration: the last warp's lane 0
    // stores sample `it` after the block-level barrier, in the shadow of the exchange it would otherwise only wait for.
    const bool writer = rank == 0 && warp == nwarps - 1 && lane == 0;
    auto store_sample = [&](int it) {
        out[(size_t)b * npoint + it] = far;
        if (new_xyz) {
            float* o = new_xyz + ((size_t)b * npoint + it) * 3;
            o[0] = cx; o[1] = cy; o[2] = cz;
        }
    };
    if (tid == 0) {
        mbar_init(&sh.cbar[0], 1);
        mbar_init(&sh.cbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (C > 1) cluster_sync_all();  // every CTA is resident and its barriers are initialised before DSMEM traffic

    for (int it = 0; it < npoint; ++it) {
        if (it == npoint - 1) {
            if (writer) store_sample(it);
            break;
        }
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
        if (writer) store_sample(it);  // far / cx / cy / cz still hold sample `it`; nobody waits for this warp now

        if (C == 1 || (PROBE && prune == 5)) {
            far = bidx;
            if (P > 0) {
                const float4 q4 = lds128(pts_sm + (uint32_t)bslot * 16u);
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
                        const float4 q4 = lds128(pts_sm + (uint32_t)bslot * 16u);
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

// ---- exact two-sample look-ahead ------------------------------------------------------------------------------------------
// One iteration of the kernel above = one cluster exchange = ONE sample.  The exchange chain is what bounds the kernel, so
// this variant retires up to TWO samples per exchange, exactly:
//   let c1 = argmax dist (the next sample) and c2 = the runner-up under the same order (value descending, index ascending).
//   Updating with c1 can only lower distances.  If it leaves c2's own distance unchanged, i.e. NOT(|c2 - c1|^2 < dist[c2])
//   in the update's own fp32 arithmetic, then after the update every other point is still <= dist[c2] under that order (a
//   point that ties c2 has a higher index), and c1 itself drops to 0 -- so c2 IS the following sample (dist[c2] > 0 keeps
//   the all-zero degenerate cloud, where the next argmax is the lowest index of all, on the one-sample path).
// Every level of the argmax therefore carries the two best candidates (thread -> warp -> CTA -> cluster), both candidates'
// coordinates travel in the exchange, every CTA evaluates the acceptance test redundantly, and the next round applies one or
// two centroids.  Samples come out in the reference's order; indices are bit-identical (test_fps_lookahead_kernel_*).
// Measured on the synthetic LiDAR scan (120k -> 512): 313 exchanges instead of 511 (tools/sim_fps_lookahead.py), but a
// round costs 1.9x (see fps_launch): an opt-in experiment (fps.lookahead = 1), not the default.
struct Fps2Shared {
    int4 wslot[2][kFpsMaxWarps];             // per-warp (best value bits, best index, second value bits, second index)
    float4 cslot[2][2 * kFpsMaxCluster];     // entry rank * 2 + k: (x, y, z, value) of CTA `rank`'s k-th candidate
    unsigned cidx[2][2 * kFpsMaxCluster];
    uint64_t cbar[2];
};

// top two of 32 (value bits, index) pairs, one per lane, under (value descending, index ascending); values are IEEE bits of
// floats >= +0 or the sentinel -1.0f, so signed integer order is the float order
__device__ __forceinline__ void warp_top2(int v, unsigned i, int& m1, unsigned& i1, int& m2, unsigned& i2) {
    m1 = __reduce_max_sync(0xffffffffu, v);
    i1 = __reduce_min_sync(0xffffffffu, v == m1 ? i : kNoIdx);
    const bool first = (v == m1) && (i == i1);
    const int v2 = first ? (int)0x80000000 : v;
    m2 = __reduce_max_sync(0xffffffffu, v2);
    i2 = __reduce_min_sync(0xffffffffu, (v2 == m2 && !first) ? i : kNoIdx);
}

template <int P>
__global__ void __launch_bounds__(kFpsMaxThreads, 1)
fps2_kernel(const float* __restrict__ xyz, int N, int npoint, const int64_t* __restrict__ start, int64_t* __restrict__ out,
            float* __restrict__ new_xyz, int log2c, int mode) {
    extern __shared__ __align__(16) float4 smem_pts[];
    __shared__ Fps2Shared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = blockDim.x >> 5;
    const unsigned C = cluster_nctarank();
    const unsigned rank = cluster_ctarank();
    const int b = blockIdx.x >> log2c;
    const float* pts = xyz + (size_t)b * N * 3;
    constexpr int CH = 32 * P;
    constexpr int PP = P / 2;
    const int chunk = warp * (int)C + (int)rank;
    const int cbase = chunk * CH;
    const int sbase = warp * CH;

    float2 px[PP], py[PP], pz[PP], pd[PP];
    float lox, loy, loz, hix, hiy, hiz;
    {
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
    }
    // cached top two of this warp's chunk: all running distances start at 1e10, lowest indices first
    const int none = (int)0xbf800000;  // bits of -1.0f: "no candidate"
    int cm1 = cbase < N ? __float_as_int(1e10f) : none, cm2 = cbase + 1 < N ? __float_as_int(1e10f) : none;
    unsigned ci1 = cbase < N ? (unsigned)cbase : kNoIdx, ci2 = cbase + 1 < N ? (unsigned)cbase + 1u : kNoIdx;

    long long far0 = start[b];
    if (far0 < 0) far0 = 0;
    if (far0 >= N) far0 = N - 1;
    // pending centroids of the round: one or two samples that are already decided but not yet applied to the distances
    int nc = 1;
    unsigned f0 = (unsigned)far0, f1 = kNoIdx;
    float c0x = pts[3 * far0], c0y = pts[3 * far0 + 1], c0z = pts[3 * far0 + 2];
    float c1x = 0.f, c1y = 0.f, c1z = 0.f;
    // shared-window addresses formed ONCE: a generic pointer into shared memory costs an S2R (SR_CgaCtaId) plus address
    // arithmetic at every use in a cluster launch, on the critical path right behind the barriers
    const uint32_t pts_sm = opaque_u32(smem_u32(smem_pts));
    const uint32_t w_sm = opaque_u32(smem_u32(&sh.wslot[0][0]));   // + par * 256
    const uint32_t c_sm = opaque_u32(smem_u32(&sh.cslot[0][0]));   // + par * 512
    const uint32_t i_sm = opaque_u32(smem_u32(&sh.cidx[0][0]));    // + par * 128
    const uint32_t b_sm = opaque_u32(smem_u32(&sh.cbar[0]));       // + par * 8
    if (tid == 0) {
        mbar_init(&sh.cbar[0], 1);
        mbar_init(&sh.cbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (C > 1) cluster_sync_all();
    const bool writer = rank == 0 && warp == nwarps - 1 && lane == 0;
    auto store_samples = [&](int at) {
        out[(size_t)b * npoint + at] = (long long)f0;
        if (nc == 2) out[(size_t)b * npoint + at + 1] = (long long)f1;
        if (new_xyz) {
            float* o = new_xyz + ((size_t)b * npoint + at) * 3;
            o[0] = c0x; o[1] = c0y; o[2] = c0z;
            if (nc == 2) { o[3] = c1x; o[4] = c1y; o[5] = c1z; }
        }
    };
    auto slot_of = [&](unsigned idx) { return (int)((idx / CH) >> log2c) * CH + (int)(idx % CH); };

    int done = 0;  // samples written before this round's pending ones
    for (int round = 0;; ++round) {
        if (done + nc >= npoint) {
            if (writer) store_samples(done);
            break;
        }
        const int par = round & 1;
        // ---- exact skip test against every pending centroid (warp-uniform) ----
        bool touched = false;
        {
            const float g0x = box_gap(lox, hix, c0x), g0y = box_gap(loy, hiy, c0y), g0z = box_gap(loz, hiz, c0z);
            const float lb0 = __fadd_rn(__fadd_rn(__fmul_rn(g0x, g0x), __fmul_rn(g0y, g0y)), __fmul_rn(g0z, g0z));
            touched = !(lb0 >= __int_as_float(cm1));
            if (nc == 2) {
                const float g1x = box_gap(lox, hix, c1x), g1y = box_gap(loy, hiy, c1y), g1z = box_gap(loz, hiz, c1z);
                const float lb1 = __fadd_rn(__fadd_rn(__fmul_rn(g1x, g1x), __fmul_rn(g1y, g1y)), __fmul_rn(g1z, g1z));
                touched = touched || !(lb1 >= __int_as_float(cm1));
            }
        }
        if (mode == 4 && round > 0) touched = false;  // latency probe (invalid results): every chunk skipped
        if (touched) {
            // ---- distance update (one or two centroids) + per-thread top two, ascending index ----
            float b1 = -1.0f, b2 = -1.0f;
            unsigned j1 = kNoIdx, j2 = kNoIdx;
            const float2 n0x = make_float2(-c0x, -c0x), n0y = make_float2(-c0y, -c0y), n0z = make_float2(-c0z, -c0z);
            const float2 n1x = make_float2(-c1x, -c1x), n1y = make_float2(-c1y, -c1y), n1z = make_float2(-c1z, -c1z);
#pragma unroll
            for (int k = 0; k < PP; ++k) {
                {
                    const float2 dx = __fadd2_rn(px[k], n0x), dy = __fadd2_rn(py[k], n0y), dz = __fadd2_rn(pz[k], n0z);
                    const float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
                    pd[k].x = fminf(pd[k].x, d.x);
                    pd[k].y = fminf(pd[k].y, d.y);
                }
                if (nc == 2) {
                    const float2 dx = __fadd2_rn(px[k], n1x), dy = __fadd2_rn(py[k], n1y), dz = __fadd2_rn(pz[k], n1z);
                    const float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
                    pd[k].x = fminf(pd[k].x, d.x);
                    pd[k].y = fminf(pd[k].y, d.y);
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float v = h == 0 ? pd[k].x : pd[k].y;
                    const unsigned idx = (unsigned)(cbase + (2 * k + h) * 32 + lane);
                    if (v > b1) { b2 = b1; j2 = j1; b1 = v; j1 = idx; }
                    else if (v > b2) { b2 = v; j2 = idx; }
                }
            }
            // ---- warp top two out of the 64 per-thread candidates ----
            const int v1 = __float_as_int(b1);
            cm1 = __reduce_max_sync(0xffffffffu, v1);
            ci1 = __reduce_min_sync(0xffffffffu, v1 == cm1 ? j1 : kNoIdx);
            const bool mine = j1 == ci1 && j1 != kNoIdx;       // this lane's best is the warp's best: offer its second
            const int vr = __float_as_int(mine ? b2 : b1);
            const unsigned jr = mine ? j2 : j1;
            cm2 = __reduce_max_sync(0xffffffffu, vr);
            ci2 = __reduce_min_sync(0xffffffffu, vr == cm2 ? jr : kNoIdx);
        }
        // ---- CTA top two (every warp, redundantly) ----
        int bm1 = cm1, bm2 = cm2;
        unsigned bi1 = ci1, bi2 = ci2;
        if (nwarps > 1) {
            const uint32_t wp = w_sm + (uint32_t)par * (uint32_t)sizeof(sh.wslot[0]);
            if (lane == 0) sts128(wp + (uint32_t)warp * 16u, cm1, (int)ci1, cm2, (int)ci2);
            __syncthreads();
            int2 e = make_int2((int)0x80000000, (int)kNoIdx);
            if (lane < 2 * nwarps) e = lds64(wp + (uint32_t)lane * 8u);
            warp_top2(e.x, (unsigned)e.y, bm1, bi1, bm2, bi2);
        }
        if (writer) store_samples(done);  // in the shadow of the exchange (see fps_kernel)
        done += nc;

        int g1v, g2v;
        unsigned g1i, g2i;
        float4 w1, w2;
        if (C == 1) {
            g1v = bm1; g1i = bi1; g2v = bm2; g2i = bi2;
            w1 = lds128(pts_sm + (uint32_t)slot_of(g1i) * 16u);
            w2 = g2i != kNoIdx ? lds128(pts_sm + (uint32_t)slot_of(g2i) * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            // ---- cluster exchange: both candidates of every CTA to every CTA (40 bytes per sender) ----
            const uint32_t bar = b_sm + (uint32_t)par * 8u;
            const uint32_t cp = c_sm + (uint32_t)par * (uint32_t)sizeof(sh.cslot[0]);
            const uint32_t ip = i_sm + (uint32_t)par * (uint32_t)sizeof(sh.cidx[0]);
            if (tid == 0) mbar_arrive_expect_tx_a(bar, 40u * C);
            if (warp == 0 && lane < 2 * (int)C) {
                const unsigned peer = (unsigned)lane & (C - 1u);
                const int k = lane >> log2c;  // 0 = best, 1 = second
                const unsigned idx = k == 0 ? bi1 : bi2;
                const int val = k == 0 ? bm1 : bm2;
                float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx != kNoIdx) q4 = lds128(pts_sm + (uint32_t)slot_of(idx) * 16u);
                const uint32_t e = rank * 2u + (uint32_t)k;
                const uint32_t rbar = mapa_shared(bar, peer);
                asm volatile(
                    "st.async.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                        mapa_shared(cp + e * 16u, peer)),
                    "f"(q4.x), "f"(q4.y), "f"(q4.z), "f"(__int_as_float(val)), "r"(rbar)
                    : "memory");
                asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.u32 [%0], %1, [%2];" ::"r"(
                                 mapa_shared(ip + e * 4u, peer)),
                             "r"(idx), "r"(rbar)
                             : "memory");
            }
            mbar_wait_a(bar, (uint32_t)((round >> 1) & 1));
            float4 a = make_float4(0.f, 0.f, 0.f, -1.0f);
            unsigned ii = kNoIdx;
            if (lane < 2 * (int)C) {
                a = lds128(cp + (uint32_t)lane * 16u);
                ii = lds32(ip + (uint32_t)lane * 4u);
            }
            warp_top2(__float_as_int(a.w), ii, g1v, g1i, g2v, g2i);
            // coordinates of the two winners: the lanes that hold them
            const unsigned h1 = __ballot_sync(0xffffffffu, ii == g1i);
            const unsigned h2 = __ballot_sync(0xffffffffu, ii == g2i && g2i != kNoIdx);
            w1 = lds128(cp + (uint32_t)(__ffs(h1) - 1) * 16u);
            w2 = h2 ? lds128(cp + (uint32_t)(__ffs(h2) - 1) * 16u) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // ---- next round's pending samples: the argmax, and the runner-up if the argmax leaves it untouched ----
        f0 = g1i; c0x = w1.x; c0y = w1.y; c0z = w1.z;
        nc = 1;
        if (g2i != kNoIdx && done + 2 <= npoint && mode != 3) {  // mode 3 (A/B knob): never accept the runner-up
            const float dx = __fadd_rn(w2.x, -c0x), dy = __fadd_rn(w2.y, -c0y), dz = __fadd_rn(w2.z, -c0z);
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            const float v2 = __int_as_float(g2v);
            if (v2 > 0.f && !(d < v2)) {
                nc = 2;
                f1 = g2i; c1x = w2.x; c1y = w2.y; c1z = w2.z;
            }
        }
    }
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
static int fps2_launch(const FpsPlan& p, const float* xyz, int B, int N, int npoint, const int64_t* start, int64_t* out,
                       float* new_xyz, cudaStream_t stream) {
    auto kern = fps2_kernel<(P > 0 ? P : 2)>;
    if (p.smem > 32 * 1024)
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
    int log2c = 0;
    while ((1 << log2c) < p.C) ++log2c;
    int mode = tuning("fps.lookahead", 2);
    PCST_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, npoint, start, out, new_xyz, log2c, mode));
    return PCST_OK;
}

template <int P>
static int fps_launch(const FpsPlan& p, const float* xyz, int B, int N, int npoint, const int64_t* start,
                      int64_t* out, float* new_xyz, float* dist_ws, cudaStream_t stream) {
    int prune = tuning("fps.prune", 1);
    prune = prune == 2 ? 0 : prune;
    // fps.lookahead = 1 / 3 / 4: the exact two-sample look-ahead kernel (1), its A/B probes (3, 4).  NOT the default: it
    // needs 313 instead of 511 exchanges on the 120k LiDAR scan, but carrying two candidates through every argmax level
    // (8 instead of 4 dependent warp reductions, twice the exchange messages, the acceptance test) makes a round 1.9x
    // as long: measured 381 us against 322 us for the one-sample kernel (profiles/r02/fps_lookahead.md).
    const int la = tuning("fps.lookahead", 2);
    if (P > 0 && prune == 1 && la != 2) return fps2_launch<P>(p, xyz, B, N, npoint, start, out, new_xyz, stream);
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
