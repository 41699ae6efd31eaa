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
//
// Kernels in this file:
//   nn_min_kernel<FORM,R,ARG,VEC2>   one direction (row minima), optionally with the argmin;
//   nn_min_pair_kernel<FORM>          BOTH directions from one sweep: every pair evaluated once, row minima in
//                                     registers, column minima by FMNMX3 tree + REDUX + per-warp strips;
//   nn_min_pair_arg_kernel            the same sweep plus both argmins (block tracking + exact fix-up kernels);
//   chamfer_bwd_kernel                gradient scatter of the Chamfer loss.
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

// 3-input minimum (sm_100: one FMNMX3 on the half-rate ALU pipe instead of two FMNMX)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
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
            // packed fp32x2 path: two rows per FFMA2 / FADD2 / FMUL2 issue slot, two candidates per
            // FMNMX3 (the minimum runs on the half-rate ALU pipe)
            auto pair2 = [&](const float4& c, int r) -> float2 {
                const float2 cx = make_float2(c.x, c.x), cy = make_float2(c.y, c.y), cz = make_float2(c.z, c.z),
                             cw = make_float2(c.w, c.w);
                const float2 vx = make_float2(ax[r], ax[r + 1]), vy = make_float2(ay[r], ay[r + 1]),
                             vz = make_float2(az[r], az[r + 1]), vn = make_float2(an[r], an[r + 1]);
                const float2 dot = __ffma2_rn(vz, cz, __ffma2_rn(vy, cy, __fmul2_rn(vx, cx)));
                if (FORM == 0) return __ffma2_rn(make_float2(-2.0f, -2.0f), dot, __fadd2_rn(vn, cw));
                if (FORM == 1) return __fadd2_rn(__fadd2_rn(dot, vn), cw);
                return __fadd2_rn(__fadd2_rn(dot, cw), vn);
            };
#pragma unroll 4
            for (int j = 0; j < kTilePoints; j += 2) {
                const float4 c0 = tile[j], c1 = tile[j + 1];
#pragma unroll
                for (int r = 0; r < R; r += 2) {
                    const float2 d0 = pair2(c0, r), d1 = pair2(c1, r);
                    best[r] = fmin3(best[r], d0.x, d1.x);
                    best[r + 1] = fmin3(best[r + 1], d0.y, d1.y);
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



// ---- both directions in one sweep ------------------------------------------------------------
// The second direction's pair matrix of the Chamfer loss is bit-for-bit the transpose of the first
// (the K=3 FMA chain and the norm sum are symmetric in their arguments), and
// the same holds for the rows / columns of torch.cdist.  nn_min_pair_kernel therefore evaluates every
// pair ONCE and reduces it into a row minimum (registers, as above) and a column minimum:
//   per candidate: minimum over the thread's R rows (FMNMX3 tree) -> clamp -> warp minimum with one
//   REDUX on the IEEE bits (non-negative floats order like unsigned integers) -> lane (j mod 32) keeps
//   it -> one STS per 32 candidates into the warp's private column strip -> after the tile, the eight
//   strips are merged and sent to HBM with atomicMin (4 B per candidate per CTA and tile).
// Work per pair: 5 FP32-pipe lane operations + ~1.1 ALU-pipe operations, against 2 x (5 + 1) for two
// one-directional sweeps.
constexpr int kPairR = 8;       // rows per thread
constexpr int kPairStages = 2;  // a 1024-candidate tile is ~50 us of work per CTA: two stages hide the 16 KiB copy

template <int FORM>
__global__ void __launch_bounds__(kNNThreads, 2)
nn_min_pair_kernel(const float4* __restrict__ A, const float4* __restrict__ Bp, int N, int M, int Npad, int Mpad,
                   unsigned int* __restrict__ rowmin_bits, unsigned int* __restrict__ colmin_bits) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tiles = reinterpret_cast<float4*>(smem_raw);                                         // [stages][1024]
    unsigned int* colw = reinterpret_cast<unsigned int*>(smem_raw + kPairStages * kTileBytes);   // [2][8 warps][1024]
    __shared__ __align__(8) uint64_t full_bar[kPairStages];
    constexpr int R = kPairR;
    constexpr int kWarps = kNNThreads / 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    // Candidate split blockIdx.y of gridDim.y owns a contiguous range of 32-candidate UNITS, balanced to one unit: the
    // splits need not be whole 1024-candidate tiles, so the host can pick the split count that fills the machine exactly
    // (a query shard of 15 000 rows is 8 row tiles: 37 splits = 296 CTAs of 101-102 units; whole-tile splits left 20 % of
    // the tile slots empty).  Tiles still arrive whole by TMA; the first and last tile of a range are swept partially.
    const int units_total = Mpad / 32;
    const int u_lo = (int)(((long long)units_total * blockIdx.y) / gridDim.y);
    const int u_hi = (int)(((long long)units_total * (blockIdx.y + 1)) / gridDim.y);
    if (u_hi <= u_lo) return;  // uniform per CTA
    constexpr int kUnitsPerTile = kTilePoints / 32;
    const int tile0 = u_lo / kUnitsPerTile;
    const int ntiles = (u_hi + kUnitsPerTile - 1) / kUnitsPerTile - tile0;
    const float4* cand = Bp + (size_t)b * Mpad + (size_t)tile0 * kTilePoints;

    if (tid == 0) {
        for (int s = 0; s < kPairStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int pre = ntiles < kPairStages ? ntiles : kPairStages;
        for (int s = 0; s < pre; ++s) {
            mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
            tma_load_1d(tiles + (size_t)s * kTilePoints, cand + (size_t)s * kTilePoints, kTileBytes, &full_bar[s]);
        }
    }

    const int row0 = blockIdx.x * (kNNThreads * R) + tid;
    const float4* arow = A + (size_t)b * Npad;
    float2 vx[R / 2], vy[R / 2], vz[R / 2], vn[R / 2];
    float best[R];
#pragma unroll
    for (int r = 0; r < R; r += 2) {
        const float4 p = arow[row0 + r * kNNThreads], q = arow[row0 + (r + 1) * kNNThreads];
        const float sc = FORM == 0 ? 1.0f : -2.0f;  // cdist: x1.mul(-2), exact
        vx[r / 2] = make_float2(sc * p.x, sc * q.x);
        vy[r / 2] = make_float2(sc * p.y, sc * q.y);
        vz[r / 2] = make_float2(sc * p.z, sc * q.z);
        vn[r / 2] = make_float2(p.w, q.w);
        best[r] = best[r + 1] = __int_as_float(0x7f800000);
    }

    auto pair8 = [&](const float4& c, float (&d)[R]) {
        const float2 cx = make_float2(c.x, c.x), cy = make_float2(c.y, c.y), cz = make_float2(c.z, c.z),
                     cw = make_float2(c.w, c.w);
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {
            float2 o;
            const float2 dot = __ffma2_rn(vz[h], cz, __ffma2_rn(vy[h], cy, __fmul2_rn(vx[h], cx)));
            if (FORM == 0) {
                o = __ffma2_rn(make_float2(-2.0f, -2.0f), dot, __fadd2_rn(vn[h], cw));  // (|a|^2 + |b|^2) + (-2 dot)
            } else {
                o = __fadd2_rn(__fadd2_rn(dot, vn[h]), cw);  // ATen _euclidean_dist: x1's norm first
            }
            d[2 * h] = o.x;
            d[2 * h + 1] = o.y;
        }
    };
    auto colmin8 = [&](const float (&d)[R]) -> unsigned int {
        float m = fmin3(fmin3(d[0], d[1], d[2]), fmin3(d[3], d[4], d[5]), fminf(d[6], d[7]));
        m = fmaxf(m, 0.0f);  // clamp(min=0) commutes with the minimum; makes the bits order like unsigned
        return __reduce_min_sync(0xffffffffu, __float_as_uint(m));
    };

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kPairStages;
        mbar_wait(&full_bar[s], (uint32_t)((t / kPairStages) & 1));
        const float4* tile = tiles + (size_t)s * kTilePoints;
        unsigned int* strip = colw + ((size_t)(t & 1) * kWarps + warp) * kTilePoints;
        const int ubase = (tile0 + t) * kUnitsPerTile;   // this CTA's part of the tile, in candidates
        const int jb_lo = (u_lo > ubase ? u_lo - ubase : 0) * 32;
        const int jb_hi = (u_hi - ubase < kUnitsPerTile ? u_hi - ubase : kUnitsPerTile) * 32;
#pragma unroll 1
        for (int jb = jb_lo; jb < jb_hi; jb += 32) {
            unsigned int mine = 0x7f800000u;
#pragma unroll 4
            for (int jj = 0; jj < 32; jj += 2) {
                float d0[R], d1[R];
                pair8(tile[jb + jj], d0);
                pair8(tile[jb + jj + 1], d1);
#pragma unroll
                for (int r = 0; r < R; ++r) best[r] = fmin3(best[r], d0[r], d1[r]);
                const unsigned int u0 = colmin8(d0), u1 = colmin8(d1);
                if (lane == jj) mine = u0;
                if (lane == jj + 1) mine = u1;
            }
            strip[jb + lane] = mine;
        }
        __syncthreads();  // all warps are done with stage s and have written their strips of parity (t & 1)
        if (tid == 0 && t + kPairStages < ntiles) {
            mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
            tma_load_1d(tiles + (size_t)s * kTilePoints, cand + (size_t)(t + kPairStages) * kTilePoints, kTileBytes,
                        &full_bar[s]);
        }
        // merge the strips; strips of this parity are rewritten in tile t + 2, after the next __syncthreads
        const unsigned int* base = colw + (size_t)(t & 1) * kWarps * kTilePoints;
        const int jbase = (tile0 + t) * kTilePoints;
        for (int q = jb_lo + tid; q < jb_hi; q += kNNThreads) {
            unsigned int v = base[q];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) v = min(v, base[w * kTilePoints + q]);
            if (jbase + q < M) atomicMin(colmin_bits + (size_t)b * M + jbase + q, v);
        }
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r * kNNThreads;
        if (i < N) atomicMin(rowmin_bits + (size_t)b * N + i, __float_as_uint(fmaxf(best[r], 0.0f)));
    }
}


// ---- both directions in one sweep, WITH the argmins the Chamfer backward needs -------------------------------
// Tracking an index per pair costs two extra ALU-pipe instructions per pair (the one-directional argmin kernel is
// ALU-bound at half the speed of the value-only kernel).  Here the sweep only remembers, per row, WHICH 32-candidate
// sub-block produced the minimum (one compare + two selects per row per 32 pairs) and, per candidate, which
// 256-row block of the grid did (the warp that wins the REDUX / strip merge); both are merged across CTAs as 64-bit
// (value bits << 32 | block) keys with atomicMin, so ties go to the lowest block.  Two small fix-up kernels then
// re-evaluate the 32 (resp. 256) pairs of the winning block with the same arithmetic and return the FIRST index
// whose clamped distance equals the minimum -- the reference's tie-break (torch.min returns the first minimum).
// Rows are laid out so that a warp owns 256 CONSECUTIVE rows (thread t holds rows 8t .. 8t+7 of the CTA's 2048).
__global__ void __launch_bounds__(kNNThreads, 2)
nn_min_pair_arg_kernel(const float4* __restrict__ A, const float4* __restrict__ Bp, int N, int M, int Npad, int Mpad,
                       int tiles_per_split, unsigned long long* __restrict__ rowkey,
                       unsigned long long* __restrict__ colkey) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float4* tiles = reinterpret_cast<float4*>(smem_raw);                                         // [stages][1024]
    unsigned int* colw = reinterpret_cast<unsigned int*>(smem_raw + kPairStages * kTileBytes);   // [2][8 warps][1024]
    __shared__ __align__(8) uint64_t full_bar[kPairStages];
    constexpr int R = kPairR;
    constexpr int kWarps = kNNThreads / 32;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.z;
    const int total_tiles = Mpad / kTilePoints;
    const int tile0 = blockIdx.y * tiles_per_split;
    int ntiles = total_tiles - tile0;
    if (ntiles > tiles_per_split) ntiles = tiles_per_split;
    if (ntiles <= 0) return;  // uniform per CTA
    const float4* cand = Bp + (size_t)b * Mpad + (size_t)tile0 * kTilePoints;

    if (tid == 0) {
        for (int s = 0; s < kPairStages; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        const int pre = ntiles < kPairStages ? ntiles : kPairStages;
        for (int s = 0; s < pre; ++s) {
            mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
            tma_load_1d(tiles + (size_t)s * kTilePoints, cand + (size_t)s * kTilePoints, kTileBytes, &full_bar[s]);
        }
    }

    const int row0 = blockIdx.x * (kNNThreads * R) + tid * R;  // 8 consecutive rows per thread
    const float4* arow = A + (size_t)b * Npad + row0;
    float2 vx[R / 2], vy[R / 2], vz[R / 2], vn[R / 2];
    float best[R];
    int bblk[R];
#pragma unroll
    for (int r = 0; r < R; r += 2) {
        const float4 p = arow[r], q = arow[r + 1];
        vx[r / 2] = make_float2(p.x, q.x);
        vy[r / 2] = make_float2(p.y, q.y);
        vz[r / 2] = make_float2(p.z, q.z);
        vn[r / 2] = make_float2(p.w, q.w);
        best[r] = best[r + 1] = __int_as_float(0x7f800000);
        bblk[r] = bblk[r + 1] = 0;
    }

    auto pair8 = [&](const float4& c, float (&d)[R]) {
        const float2 cx = make_float2(c.x, c.x), cy = make_float2(c.y, c.y), cz = make_float2(c.z, c.z),
                     cw = make_float2(c.w, c.w);
#pragma unroll
        for (int h = 0; h < R / 2; ++h) {
            const float2 dot = __ffma2_rn(vz[h], cz, __ffma2_rn(vy[h], cy, __fmul2_rn(vx[h], cx)));
            const float2 o = __ffma2_rn(make_float2(-2.0f, -2.0f), dot, __fadd2_rn(vn[h], cw));  // losses.py:38
            d[2 * h] = o.x;
            d[2 * h + 1] = o.y;
        }
    };
    auto colmin8 = [&](const float (&d)[R]) -> unsigned int {
        float m = fmin3(fmin3(d[0], d[1], d[2]), fmin3(d[3], d[4], d[5]), fminf(d[6], d[7]));
        m = fmaxf(m, 0.0f);
        return __reduce_min_sync(0xffffffffu, __float_as_uint(m));
    };

    const unsigned int my_block = blockIdx.x * kWarps + warp;  // this warp's 256-row block of the grid
    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kPairStages;
        mbar_wait(&full_bar[s], (uint32_t)((t / kPairStages) & 1));
        const float4* tile = tiles + (size_t)s * kTilePoints;
        unsigned int* strip = colw + ((size_t)(t & 1) * kWarps + warp) * kTilePoints;
#pragma unroll 1
        for (int jb = 0; jb < kTilePoints; jb += 32) {
            unsigned int mine = 0x7f800000u;
            float sub[R];  // minimum of this 32-candidate sub-block per row
#pragma unroll
            for (int r = 0; r < R; ++r) sub[r] = __int_as_float(0x7f800000);
#pragma unroll 4
            for (int jj = 0; jj < 32; jj += 2) {
                float d0[R], d1[R];
                pair8(tile[jb + jj], d0);
                pair8(tile[jb + jj + 1], d1);
#pragma unroll
                for (int r = 0; r < R; ++r) sub[r] = fmin3(sub[r], d0[r], d1[r]);
                const unsigned int u0 = colmin8(d0), u1 = colmin8(d1);
                if (lane == jj) mine = u0;
                if (lane == jj + 1) mine = u1;
            }
            const int blk = (tile0 + t) * (kTilePoints / 32) + (jb >> 5);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float m = fmaxf(sub[r], 0.0f);  // clamp, then strict <: ties keep the earlier sub-block
                if (m < best[r]) {
                    best[r] = m;
                    bblk[r] = blk;
                }
            }
            strip[jb + lane] = mine;
        }
        __syncthreads();
        if (tid == 0 && t + kPairStages < ntiles) {
            mbar_arrive_expect_tx(&full_bar[s], kTileBytes);
            tma_load_1d(tiles + (size_t)s * kTilePoints, cand + (size_t)(t + kPairStages) * kTilePoints, kTileBytes,
                        &full_bar[s]);
        }
        const unsigned int* base = colw + (size_t)(t & 1) * kWarps * kTilePoints;
        const int jbase = (tile0 + t) * kTilePoints;
        for (int q = tid; q < kTilePoints; q += kNNThreads) {
            unsigned int v = base[q], w = 0;
#pragma unroll
            for (int x = 1; x < kWarps; ++x) {
                const unsigned int u = base[x * kTilePoints + q];
                if (u < v) {  // strict: the lowest warp (= lowest rows) wins a tie
                    v = u;
                    w = x;
                }
            }
            if (jbase + q < M)
                atomicMin(colkey + (size_t)b * M + jbase + q,
                          ((unsigned long long)v << 32) | (unsigned long long)(blockIdx.x * kWarps + w));
        }
    }
    (void)my_block;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int i = row0 + r;
        if (i < N)
            atomicMin(rowkey + (size_t)b * N + i,
                      ((unsigned long long)__float_as_uint(best[r]) << 32) | (unsigned long long)(unsigned int)bblk[r]);
    }
}

__global__ void nn_min_key_init_kernel(unsigned long long* a, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = ~0ull;
}

__device__ __forceinline__ float pair_loss_clamped(const float4& a, const float4& c) {
    return fmaxf(__fmaf_rn(-2.0f, dot3_chain(a.x, a.y, a.z, c.x, c.y, c.z), __fadd_rn(a.w, c.w)), 0.0f);
}

// rows: one thread per row re-evaluates the 32 candidates of the winning sub-block
__global__ void nn_min_pair_fix_rows_kernel(const float4* __restrict__ A, const float4* __restrict__ Bp, int N, int M,
                                            int Npad, int Mpad, const unsigned long long* __restrict__ rowkey,
                                            float* __restrict__ rowmin, int64_t* __restrict__ rowarg) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned long long key = rowkey[(size_t)b * N + i];
    const float best = __uint_as_float((unsigned int)(key >> 32));
    const int j0 = (int)(unsigned int)(key & 0xffffffffull) * 32;
    const float4 a = A[(size_t)b * Npad + i];
    const float4* c = Bp + (size_t)b * Mpad + j0;
    int arg = j0;
#pragma unroll 4
    for (int j = 31; j >= 0; --j)
        if (pair_loss_clamped(a, c[j]) == best) arg = j0 + j;  // descending: the lowest matching index survives
    rowmin[(size_t)b * N + i] = best;
    rowarg[(size_t)b * N + i] = arg;
}

// columns: one warp per candidate re-evaluates the 256 rows of the winning block
__global__ void nn_min_pair_fix_cols_kernel(const float4* __restrict__ A, const float4* __restrict__ Bp, int N, int M,
                                            int Npad, int Mpad, const unsigned long long* __restrict__ colkey,
                                            float* __restrict__ colmin, int64_t* __restrict__ colarg) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= M) return;  // whole warp
    const unsigned long long key = colkey[(size_t)b * M + j];
    const float best = __uint_as_float((unsigned int)(key >> 32));
    const int i0 = (int)(unsigned int)(key & 0xffffffffull) * 256;
    const float4 c = Bp[(size_t)b * Mpad + j];
    const float4* a = A + (size_t)b * Npad + i0;
    unsigned int arg = 0xffffffffu;
#pragma unroll
    for (int k = 7; k >= 0; --k) {
        const int i = k * 32 + lane;
        if (pair_loss_clamped(a[i], c) == best) arg = (unsigned int)(i0 + i);
    }
    arg = __reduce_min_sync(0xffffffffu, arg);
    if (lane == 0) {
        colmin[(size_t)b * M + j] = best;
        colarg[(size_t)b * M + j] = (int64_t)arg;
    }
}

__global__ void nn_min_pair_init_kernel(unsigned int* a, size_t na, unsigned int* b, size_t nb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) a[i] = 0x7f800000u;
    else if (i - na < nb) b[i - na] = 0x7f800000u;
}
__global__ void nn_min_pair_sqrt_kernel(float* a, size_t na, float* b, size_t nb) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) a[i] = __fsqrt_rn(a[i]);
    else if (i - na < nb) b[i - na] = __fsqrt_rn(b[i - na]);
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
    int pair_splits;  // candidate splits of nn_min_pair_kernel (unit-balanced)
    size_t off_a, off_b, off_key, total;
};

static NNPlan make_plan(int B, int N, int M, int R = 4) {
    NNPlan p;
    p.R = R;
    const int rows_per_cta = kNNThreads * p.R;
    p.Npad = (int)align_up(align_up((size_t)N, rows_per_cta), kTilePoints);
    p.Mpad = padded_points(M);
    p.row_tiles = (N + rows_per_cta - 1) / rows_per_cta;
    const int cand_tiles = p.Mpad / kTilePoints;
    // choose the number of candidate splits that best fills whole waves of 2 CTAs/SM x 148 SMs
    const long wave = 2L * num_sms();
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
    {   // nn_min_pair_kernel balances its splits to one 32-candidate unit: any split count is as good as its wave fill
        const int units = p.Mpad / 32;
        int bs = 1;
        double be = -1.0;
        for (int s = 1; s <= units && s <= 1024; ++s) {
            if (forced) s = forced < units ? forced : units;
            const long ctas = (long)p.row_tiles * s * B;
            const long waves = (ctas + wave - 1) / wave;
            // per-CTA cost = its units + ~11 units of fixed overhead (row load, atomics, launch tail) + the partial tiles' fetch
            const double eff = ((double)p.row_tiles * units * B) / ((double)waves * wave * ((double)units / s + 11.0));
            if (eff > be + 1e-9) {
                be = eff;
                bs = s;
            }
            if (forced) break;
        }
        p.pair_splits = bs;
    }
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


extern "C" size_t pcst_nn_min_pair_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    return make_plan(B, N, M, kPairR).total;
}

extern "C" int pcst_nn_min_pair_f32(const float* a, const float* b, int B, int N, int M, int form, float* rowmin,
                                    float* colmin, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(a && b && rowmin && colmin, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && M > 0, "B, N, M must be positive");
    PCST_CHECK_ARG(B <= 65535, "B must be <= 65535");
    PCST_CHECK_ARG(form == 0 || form == 1, "form must be 0 (loss) or 1 (cdist)");
    const NNPlan p = make_plan(B, N, M, kPairR);
    if (!ws || ws_bytes < p.total || ((uintptr_t)ws & 255)) {
        set_error("pcst_nn_min_pair_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, p.total);
        return PCST_ERR_WORKSPACE;
    }
    char* w = (char*)ws;
    float4* A = (float4*)(w + p.off_a);
    float4* Bp = (float4*)(w + p.off_b);
    int st;
    if ((st = launch_pack(a, B, N, p.Npad, A, stream)) != PCST_OK) return st;
    if ((st = launch_pack(b, B, M, p.Mpad, Bp, stream)) != PCST_OK) return st;
    const size_t na = (size_t)B * N, nb = (size_t)B * M;
    nn_min_pair_init_kernel<<<(unsigned)((na + nb + 255) / 256), 256, 0, stream>>>((unsigned int*)rowmin, na,
                                                                                    (unsigned int*)colmin, nb);
    PCST_CUDA(cudaGetLastError());
    const int smem = kPairStages * kTileBytes + 2 * (kNNThreads / 32) * kTilePoints * (int)sizeof(unsigned int);
    dim3 grid(p.row_tiles, p.pair_splits, B);
    if (form == 0) {
        auto kern = nn_min_pair_kernel<0>;
        PCST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, kNNThreads, smem, stream>>>(A, Bp, N, M, p.Npad, p.Mpad, (unsigned int*)rowmin, (unsigned int*)colmin);
    } else {
        auto kern = nn_min_pair_kernel<1>;
        PCST_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, kNNThreads, smem, stream>>>(A, Bp, N, M, p.Npad, p.Mpad, (unsigned int*)rowmin, (unsigned int*)colmin);
    }
    PCST_CUDA(cudaGetLastError());
    if (form != 0) {
        nn_min_pair_sqrt_kernel<<<(unsigned)((na + nb + 255) / 256), 256, 0, stream>>>(rowmin, na, colmin, nb);
        PCST_CUDA(cudaGetLastError());
    }
    return PCST_OK;
}


extern "C" size_t pcst_nn_min_pair_arg_workspace_bytes(int B, int N, int M) {
    if (B <= 0 || N <= 0 || M <= 0) return 0;
    const NNPlan p = make_plan(B, N, M, kPairR);
    return p.off_key + align_up((size_t)B * N * 8, 256) + align_up((size_t)B * M * 8, 256);
}

extern "C" int pcst_nn_min_pair_arg_f32(const float* a, const float* b, int B, int N, int M, float* rowmin,
                                        int64_t* rowarg, float* colmin, int64_t* colarg, void* ws, size_t ws_bytes,
                                        pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(a && b && rowmin && rowarg && colmin && colarg, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0 && M > 0, "B, N, M must be positive");
    PCST_CHECK_ARG(B <= 65535, "B must be <= 65535");
    const NNPlan p = make_plan(B, N, M, kPairR);
    const size_t need = pcst_nn_min_pair_arg_workspace_bytes(B, N, M);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
        set_error("pcst_nn_min_pair_arg_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
        return PCST_ERR_WORKSPACE;
    }
    char* w = (char*)ws;
    float4* A = (float4*)(w + p.off_a);
    float4* Bp = (float4*)(w + p.off_b);
    unsigned long long* rowkey = (unsigned long long*)(w + p.off_key);
    unsigned long long* colkey = (unsigned long long*)(w + p.off_key + align_up((size_t)B * N * 8, 256));
    int st;
    if ((st = launch_pack(a, B, N, p.Npad, A, stream)) != PCST_OK) return st;
    if ((st = launch_pack(b, B, M, p.Mpad, Bp, stream)) != PCST_OK) return st;
    const size_t nkeys = ((char*)colkey - (char*)rowkey) / 8 + (size_t)B * M;  // both key arrays and the padding between
    nn_min_key_init_kernel<<<(unsigned)((nkeys + 255) / 256), 256, 0, stream>>>(rowkey, nkeys);
    PCST_CUDA(cudaGetLastError());
    const int smem = kPairStages * kTileBytes + 2 * (kNNThreads / 32) * kTilePoints * (int)sizeof(unsigned int);
    PCST_CUDA(cudaFuncSetAttribute(nn_min_pair_arg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dim3 grid(p.row_tiles, p.splits, B);
    nn_min_pair_arg_kernel<<<grid, kNNThreads, smem, stream>>>(A, Bp, N, M, p.Npad, p.Mpad, p.tiles_per_split, rowkey, colkey);
    PCST_CUDA(cudaGetLastError());
    nn_min_pair_fix_rows_kernel<<<dim3((N + 255) / 256, B), 256, 0, stream>>>(A, Bp, N, M, p.Npad, p.Mpad, rowkey, rowmin, rowarg);
    PCST_CUDA(cudaGetLastError());
    nn_min_pair_fix_cols_kernel<<<dim3((M + 7) / 8, B), 256, 0, stream>>>(A, Bp, N, M, p.Npad, p.Mpad, colkey, colmin, colarg);
    return check_cuda(cudaGetLastError(), "nn_min_pair_fix_cols_kernel");
}

// ---- backward of the Chamfer loss (models/losses.py:24-61 under autograd) ---------------------
// chamfer[b] = mean_i D(p_i, t_{a(i)}) + mean_j D(t_j, p_{c(j)}),  dD/dp = 2 (p - t), dD/dt = -2 (p - t).
// D = clamp(raw, min=0) (losses.py:39,56): where the expanded-form distance rounds BELOW zero the clamp is active and
// autograd passes no gradient; the raw value is re-evaluated here in the forward's own rounding order to decide that.
__device__ __forceinline__ bool chamfer_pair_clamped(const float* p, const float* t) {
    const float raw = __fmaf_rn(-2.0f, dot3_chain(p[0], p[1], p[2], t[0], t[1], t[2]),
                                __fadd_rn(norm3_sq(p[0], p[1], p[2]), norm3_sq(t[0], t[1], t[2])));
    return raw < 0.0f;
}
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
            if (chamfer_pair_clamped(P + 3 * i, T + 3 * j)) continue;
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
            if (chamfer_pair_clamped(P + 3 * i, T + 3 * j)) continue;   // the second direction's matrix is the exact transpose
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
    if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
    chamfer_bwd_kernel<<<dim3(blocks, B), 256, 0, stream>>>(pred, target, arg_pt, arg_tp, grad_out, N, M, grad_pred,
                                                             grad_target);
    return check_cuda(cudaGetLastError(), "chamfer_bwd_kernel");
}

// ---- query-sharded Chamfer: the small kernels either side of its single collective --------------------------------------
// (SURVEY.md 8(e) variant 2.)  Rank r has swept its [n_r x M] tile: complete row minima of its n_r queries, PARTIAL column
// minima of all M targets.  Instead of all_reduce(MIN) of the columns + all_reduce(SUM) of the row sums + host-side glue:
//   pack   : payload[b] = colmin[b, 0..M) | kShardParts fp64 partial sums of rowmin[b, :]      -> ONE all-gather of payloads
//   finish : out[b] = sum_r rowsum_r / N  +  sum_m min_r colmin_r[m] / M       (fp64 sums, / 2 for the metric form)
// Both are spread over many CTAs (a 120k-column row is 480 KB) and deterministic: partial sums are written to fixed slots
// and added in a fixed order, so every rank gets the same bits.
namespace pcst {

constexpr int kShardParts = 128;  // CTAs per batch element in pack / finish = fp64 partial sums per rank and element
constexpr int kShardThreads = 256;

__device__ __forceinline__ double block_sum_f64(double s, double* part) {
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
    return tot;  // valid in thread 0
}

// payload row: M floats (column minima), then kShardParts doubles stored as float pairs (row stride M + 2 * kShardParts)
__global__ void __launch_bounds__(kShardThreads)
chamfer_shard_pack_kernel(const float* __restrict__ rowmin, const float* __restrict__ colmin, int n, int M,
                          float* __restrict__ payload) {
    __shared__ double part[kShardThreads / 32];
    const int b = blockIdx.y, c = blockIdx.x, t = threadIdx.x;
    const int stride = M + 2 * kShardParts;
    float* out = payload + (size_t)b * stride;
    const float* cm = colmin + (size_t)b * M;
    for (int m = c * kShardThreads + t; m < M; m += kShardParts * kShardThreads) out[m] = cm[m];
    double s = 0.0;
    const float* rm = rowmin + (size_t)b * n;
    for (int i = c * kShardThreads + t; i < n; i += kShardParts * kShardThreads) s += (double)rm[i];
    const double tot = block_sum_f64(s, part);
    if (t == 0) {
        const unsigned long long bits = (unsigned long long)__double_as_longlong(tot);
        out[M + 2 * c] = __uint_as_float((unsigned)(bits & 0xffffffffull));
        out[M + 2 * c + 1] = __uint_as_float((unsigned)(bits >> 32));
    }
}

// stage 1: CTA c of element b: min over the G ranks of its column slice, fp64 sum -> colpart[b][c]
__global__ void __launch_bounds__(kShardThreads)
chamfer_shard_colmin_kernel(const float* __restrict__ gathered, int G, int B, int M, double* __restrict__ colpart) {
    __shared__ double part[kShardThreads / 32];
    const int b = blockIdx.y, c = blockIdx.x, t = threadIdx.x;
    const size_t stride = (size_t)B * (M + 2 * kShardParts);
    const float* base = gathered + (size_t)b * (M + 2 * kShardParts);
    double s = 0.0;
    for (int m = c * kShardThreads + t; m < M; m += kShardParts * kShardThreads) {
        float v = base[m];
        for (int g = 1; g < G; ++g) v = fminf(v, base[(size_t)g * stride + m]);
        s += (double)v;
    }
    const double tot = block_sum_f64(s, part);
    if (t == 0) colpart[(size_t)b * kShardParts + c] = tot;
}

// stage 2: one CTA of kShardParts threads per element: thread c adds column partial c and the G row partials of slot c,
// then a fixed-shape tree over the slots (deterministic: every rank performs the same additions in the same order)
__global__ void __launch_bounds__(kShardParts)
chamfer_shard_finish_kernel(const float* __restrict__ gathered, const double* __restrict__ colpart, int G, int B, int M,
                            double n_total, int form, float* __restrict__ out) {
    __shared__ double cs[kShardParts], rs[kShardParts];
    const int b = blockIdx.x, c = threadIdx.x;
    const size_t stride = (size_t)B * (M + 2 * kShardParts);
    double rows = 0.0;
    for (int g = 0; g < G; ++g) {
        const float* p = gathered + (size_t)g * stride + (size_t)b * (M + 2 * kShardParts) + M + 2 * c;
        const unsigned long long bits = (unsigned long long)__float_as_uint(p[0]) | ((unsigned long long)__float_as_uint(p[1]) << 32);
        rows += __longlong_as_double((long long)bits);
    }
    cs[c] = colpart[(size_t)b * kShardParts + c];
    rs[c] = rows;
    __syncthreads();
    for (int o = kShardParts / 2; o > 0; o >>= 1) {
        if (c < o) {
            cs[c] += cs[c + o];
            rs[c] += rs[c + o];
        }
        __syncthreads();
    }
    if (c == 0) {
        double v = rs[0] / n_total + cs[0] / (double)M;
        if (form != 0) v *= 0.5;
        out[b] = (float)v;
    }
}

}  // namespace pcst

extern "C" int pcst_chamfer_shard_payload_floats(int M) { return M > 0 ? M + 2 * pcst::kShardParts : 0; }

extern "C" int pcst_chamfer_shard_pack_f32(const float* rowmin, const float* colmin, int B, int n, int M, float* payload,
                                           pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(colmin && payload && (rowmin || n == 0), "null pointer");
    PCST_CHECK_ARG(B > 0 && B <= 65535 && n >= 0 && M > 0, "bad sizes");
    chamfer_shard_pack_kernel<<<dim3(kShardParts, B), kShardThreads, 0, stream>>>(rowmin, colmin, n, M, payload);
    return check_cuda(cudaGetLastError(), "chamfer_shard_pack_kernel");
}

extern "C" int pcst_chamfer_shard_finish_f32(const float* gathered, int G, int B, int M, long long n_total, int form, float* out,
                                             void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(gathered && out && ws, "null pointer");
    PCST_CHECK_ARG(G > 0 && B > 0 && B <= 65535 && M > 0 && n_total > 0, "bad sizes");
    if (ws_bytes < (size_t)B * kShardParts * sizeof(double) || ((uintptr_t)ws & 255)) {
        set_error("pcst_chamfer_shard_finish_f32: workspace too small or misaligned (needs B * 128 doubles)");
        return PCST_ERR_WORKSPACE;
    }
    chamfer_shard_colmin_kernel<<<dim3(kShardParts, B), kShardThreads, 0, stream>>>(gathered, G, B, M, (double*)ws);
    PCST_CUDA(cudaGetLastError());
    chamfer_shard_finish_kernel<<<B, kShardParts, 0, stream>>>(gathered, (const double*)ws, G, B, M, (double)n_total, form, out);
    return check_cuda(cudaGetLastError(), "chamfer_shard_finish_kernel");
}
