// knn.cu -- k nearest neighbours + inverse-distance interpolation (sm_100a).
//
// Replaces the CPU round trip of HierarchicalProcessor.upsample_knn (models/diffusion_model.py:
// 133-152: .cpu().numpy(), sklearn kd-tree build + query in fp64, numpy weights, .to(device)) and
// the sklearn kNN of coverage_score / uniformity_score (evaluation/metrics.py:126-127,152-153).
//
// Parity target is sklearn's fp64 result, so the k best (r, j) per query are ranked on distances evaluated in
// fp64 in sklearn's summation order, r = ((dx*dx) + (dy*dy)) + (dz*dz), with explicit non-fused double intrinsics;
// the list stays sorted in registers (strict <, ascending j => ties to the lower index).  Almost every pair is
// rejected earlier by an fp32 PREFILTER: d32 = fma(dz,dz, fma(dy,dy, dx*dx)) carries a relative error below
// 5 * 2^-24, so a candidate whose exact distance beats the current k-th best always satisfies
// d32 <= roundup(kth_best) * (1 + 2e-6); only those (a few dozen per query: the list converges like a
// running minimum) pay the fp64 evaluation and the insertion.  Results are therefore identical to the all-fp64
// sweep, at FP32-pipe speed.  One thread per query; reference points are staged through shared memory as
// per-component pairs of consecutive candidates shared by the whole CTA (three broadcast LDS.64 per two pairs, and
// one packed fp32x2 instruction sequence for both).
// Bound: FP32 CUDA cores (6 FP32-pipe operations per pair); HBM traffic is negligible.
#include "common.cuh"

namespace pcst {

constexpr int kKnnThreads = 128;
constexpr int kKnnTile = 1024;  // reference points per shared-memory tile (16 KiB as float4)
constexpr int kKnnMaxK = 16;

template <int KMAX>
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(const float* __restrict__ query, const float* __restrict__ ref, int Q, int R, int k,
           int64_t* __restrict__ idx, double* __restrict__ dist, int splits, int tiles_per_split,
           double* __restrict__ part_d, int* __restrict__ part_i, const int* __restrict__ qlist = nullptr,
           const int* __restrict__ qcount = nullptr, const int* __restrict__ only_if = nullptr) {
    // reference points of the tile as pairs of consecutive candidates per component, so that one packed
    // FADD2 / FMUL2 / FFMA2 sequence prefilters two candidates against the thread's query
    // (KMAX <= 4) or as float4 rows, one broadcast LDS.128 per candidate (longer lists)
    __shared__ __align__(16) float4 tile[kKnnTile];
    float2* tx = reinterpret_cast<float2*>(tile);
    float2* ty = tx + kKnnTile / 2;
    float2* tz = ty + kKnnTile / 2;
    const int b = blockIdx.y;
    if (only_if && !only_if[b]) return;   // (grid search) only the clouds flagged for the plain sweep
    int q = blockIdx.x * kKnnThreads + threadIdx.x;
    const int slot = q;   // listed mode: position in the list (partial results are stored per slot)
    bool active = q < Q;
    if (qlist) {
        // the queries the grid search handed back (knn_grid.cu): slot -> query index; CTAs past the list leave at once
        const int n = qcount[b];
        if ((int)blockIdx.x * kKnnThreads >= n) return;
        active = q < n;
        q = active ? qlist[(size_t)b * Q + q] : 0;
    }
    const float* qp = query + ((size_t)b * Q + (active ? q : 0)) * 3;
    const float fx = qp[0], fy = qp[1], fz = qp[2];
    const double qx = fx, qy = fy, qz = fz;
    const float2 nqx = make_float2(-fx, -fx), nqy = make_float2(-fy, -fy), nqz = make_float2(-fz, -fz);
    const float* rp = ref + (size_t)b * R * 3;

    double bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
        bd[t] = __longlong_as_double(0x7ff0000000000000ll);  // +inf
        bi[t] = 0;
    }
    float thr = __int_as_float(0x7f800000);  // fp32 upper bound of bd[KMAX - 1]

    auto consider = [&](float cx, float cy, float cz, int j) {
        const double dx = __dsub_rn(qx, (double)cx), dy = __dsub_rn(qy, (double)cy), dz = __dsub_rn(qz, (double)cz);
        const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        if (d < bd[KMAX - 1]) {
            // sorted insertion, fully unrolled so the list stays in registers
            double cd = d;
            int ci = j;
#pragma unroll
            for (int u = 0; u < KMAX; ++u) {
                if (cd < bd[u]) {
                    const double td = bd[u];
                    const int ti = bi[u];
                    bd[u] = cd;
                    bi[u] = ci;
                    cd = td;
                    ci = ti;
                }
            }
            thr = __fmul_ru(__double2float_ru(bd[KMAX - 1]), 1.000002f);  // +inf stays +inf
        }
    };

    // blockIdx.z = slice of the reference range (few queries: the scan is split so that the launch fills the GPU)
    const int jbegin = blockIdx.z * tiles_per_split * kKnnTile;
    const int jend = min(R, jbegin + tiles_per_split * kKnnTile);
    for (int j0 = jbegin; j0 < jend; j0 += kKnnTile) {
        const int n = jend - j0 < kKnnTile ? jend - j0 : kKnnTile;
        __syncthreads();
        for (int t = threadIdx.x; t < kKnnTile; t += kKnnThreads) {
            // slots beyond the cloud hold +inf coordinates: their distance is +inf (or NaN) and never passes `<=`
            float x = __int_as_float(0x7f800000), y = x, z = x;
            if (t < n) {
                const float* p = rp + (size_t)(j0 + t) * 3;
                x = p[0]; y = p[1]; z = p[2];
            }
            if constexpr (KMAX <= 4) {
                reinterpret_cast<float*>(tx)[t] = x;
                reinterpret_cast<float*>(ty)[t] = y;
                reinterpret_cast<float*>(tz)[t] = z;
            } else {
                tile[t] = make_float4(x, y, z, 0.f);
            }
        }
        __syncthreads();
        if (!active) continue;
        if constexpr (KMAX <= 4) {
            // short lists (the 3-NN of upsample_knn): two candidates per packed prefilter step (measured 1.35 ms
            // against 1.75 ms for 90k x 30k; with longer lists the two inlined insertion paths cost more than that)
            const int npairs = (n + 1) / 2;
#pragma unroll 4
            for (int t = 0; t < npairs; ++t) {
                const float2 cx = tx[t], cy = ty[t], cz = tz[t];
                // e = c - q (the sign does not matter for the square), d32 = fma(ez,ez, fma(ey,ey, ex*ex))
                const float2 ex = __fadd2_rn(cx, nqx), ey = __fadd2_rn(cy, nqy), ez = __fadd2_rn(cz, nqz);
                const float2 d32 = __ffma2_rn(ez, ez, __ffma2_rn(ey, ey, __fmul2_rn(ex, ex)));
                if (d32.x <= thr) consider(cx.x, cy.x, cz.x, j0 + 2 * t);
                if (d32.y <= thr) consider(cx.y, cy.y, cz.y, j0 + 2 * t + 1);  // (thr may just have tightened: still a superset)
            }
        } else {
#pragma unroll 4
            for (int t = 0; t < n; ++t) {
                const float4 c = tile[t];
                const float ex = fx - c.x, ey = fy - c.y, ez = fz - c.z;
                const float d32 = fmaf(ez, ez, fmaf(ey, ey, ex * ex));
                if (d32 <= thr) consider(c.x, c.y, c.z, j0 + t);
            }
        }
    }
    if (active) {
        for (int t = 0; t < k; ++t) {
            // KMAX >= k: the first k entries of the KMAX-list are the k best
            double v = 0.0;
            int j = 0;
#pragma unroll
            for (int u = 0; u < KMAX; ++u)
                if (u == t) { v = bd[u]; j = bi[u]; }
            if (splits == 1) {
                idx[((size_t)b * Q + q) * k + t] = j;
                dist[((size_t)b * Q + q) * k + t] = __dsqrt_rn(v);
            } else {  // squared distances of this slice's k best; merged by knn_merge_kernel
                const size_t o = (((size_t)b * Q + (qlist ? slot : q)) * splits + blockIdx.z) * k + t;
                part_d[o] = v;
                part_i[o] = j;
            }
        }
    }
}

// k best of `splits` sorted partial lists per query.  Slices cover ascending reference ranges and each list is sorted
// with ties in index order, so inserting them slice by slice with a strict `<` keeps the lowest index among equals.
__global__ void knn_merge_kernel(const double* __restrict__ part_d, const int* __restrict__ part_i, long total, int splits,
                                 int k, int64_t* __restrict__ idx, double* __restrict__ dist) {
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= total) return;
    double bd[kKnnMaxK];
    int bi[kKnnMaxK];
    for (int t = 0; t < k; ++t) {
        bd[t] = __longlong_as_double(0x7ff0000000000000ll);
        bi[t] = 0;
    }
    const double* pd = part_d + (size_t)q * splits * k;
    const int* pi = part_i + (size_t)q * splits * k;
    for (int e = 0; e < splits * k; ++e) {
        double cd = pd[e];
        int ci = pi[e];
        if (!(cd < bd[k - 1])) continue;
        for (int u = 0; u < k; ++u) {
            if (cd < bd[u]) {
                const double td = bd[u];
                const int ti = bi[u];
                bd[u] = cd;
                bi[u] = ci;
                cd = td;
                ci = ti;
            }
        }
    }
    for (int t = 0; t < k; ++t) {
        idx[(size_t)q * k + t] = bi[t];
        dist[(size_t)q * k + t] = __dsqrt_rn(bd[t]);
    }
}

// the same merge for LISTED queries: slot s of cloud b (s < qcount[b]) is query qlist[b, s]
__global__ void knn_merge_listed_kernel(const double* __restrict__ part_d, const int* __restrict__ part_i, int Q, int splits, int k,
                                        const int* __restrict__ qlist, const int* __restrict__ qcount,
                                        int64_t* __restrict__ idx, double* __restrict__ dist) {
    const int b = blockIdx.y;
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= qcount[b]) return;
    const int q = qlist[(size_t)b * Q + s];
    double bd[kKnnMaxK];
    int bi[kKnnMaxK];
    for (int t = 0; t < k; ++t) {
        bd[t] = __longlong_as_double(0x7ff0000000000000ll);
        bi[t] = 0;
    }
    const double* pd = part_d + ((size_t)b * Q + s) * splits * k;
    const int* pi = part_i + ((size_t)b * Q + s) * splits * k;
    for (int e = 0; e < splits * k; ++e) {
        double cd = pd[e];
        int ci = pi[e];
        if (!(cd < bd[k - 1])) continue;
        for (int u = 0; u < k; ++u) {
            if (cd < bd[u]) {
                const double td = bd[u];
                const int ti = bi[u];
                bd[u] = cd;
                bi[u] = ci;
                cd = td;
                ci = ti;
            }
        }
    }
    for (int t = 0; t < k; ++t) {
        idx[((size_t)b * Q + q) * k + t] = bi[t];
        dist[((size_t)b * Q + q) * k + t] = __dsqrt_rn(bd[t]);
    }
}

// out[b,q,:] = sum_k w_k feat[b, idx_k, :],  w = 1/(dist + 1e-8), normalised; fp64, rounded once.
__global__ void knn_interpolate_kernel(const float* __restrict__ feat, const int64_t* __restrict__ idx,
                                       const double* __restrict__ dist, int R, int Q, int k, int C,
                                       float* __restrict__ out) {
    const int b = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int64_t* id = idx + ((size_t)b * Q + q) * k;
    const double* di = dist + ((size_t)b * Q + q) * k;
    double w[kKnnMaxK];
    double sum = 0.0;
    for (int t = 0; t < k; ++t) {
        w[t] = __ddiv_rn(1.0, __dadd_rn(di[t], 1e-8));
        sum = __dadd_rn(sum, w[t]);  // numpy sum over the short axis: sequential
    }
    for (int t = 0; t < k; ++t) w[t] = __ddiv_rn(w[t], sum);
    for (int c = 0; c < C; ++c) {
        double acc = 0.0;
        for (int t = 0; t < k; ++t) {
            const double f = (double)feat[((size_t)b * R + id[t]) * C + c];
            const double term = __dmul_rn(f, w[t]);
            acc = t == 0 ? term : __dadd_rn(acc, term);
        }
        out[((size_t)b * Q + q) * C + c] = (float)acc;
    }
}

size_t knn_grid_workspace_bytes(int B, int Q, int R, int k);
int knn_grid_run(const float* query, const float* ref, int B, int Q, int R, int k, int64_t* idx, double* dist, void* ws,
                 cudaStream_t stream);

// brute-force sweep restricted to the queries listed in qlist [B, Q] (first qcount[b] entries of row b): the grid search's
// fallback for queries it could not finish inside its budget.  The list is usually short (a CTA or two), so the reference
// range is cut into kListedSplits slices that run as separate CTAs (one CTA sweeping 120k references alone takes a
// millisecond) and a small kernel merges the per-slice lists.  part_d / part_i: [B, Q, kListedSplits, k] (by list slot).
// the plain sweep over ALL queries of the clouds whose flag is set (the grid search's verdict "too crowded for a grid")
int knn_sweep_flagged(const float* query, const float* ref, int B, int Q, int R, int k, int64_t* idx, double* dist,
                      const int* flag, cudaStream_t stream) {
    const int tiles = (R + kKnnTile - 1) / kKnnTile;
    dim3 grid((Q + kKnnThreads - 1) / kKnnThreads, B, 1);
    if (k <= 1) knn_kernel<1><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, 1, tiles, nullptr, nullptr, nullptr, nullptr, flag);
    else if (k <= 4) knn_kernel<4><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, 1, tiles, nullptr, nullptr, nullptr, nullptr, flag);
    else if (k <= 9) knn_kernel<9><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, 1, tiles, nullptr, nullptr, nullptr, nullptr, flag);
    else knn_kernel<16><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, 1, tiles, nullptr, nullptr, nullptr, nullptr, flag);
    return check_cuda(cudaGetLastError(), "knn_kernel (flagged clouds)");
}

constexpr int kListedSplits = 16;
size_t knn_listed_part_bytes(int B, int Q, int k) {
    return align_up((size_t)B * Q * kListedSplits * k * sizeof(double), 256) + align_up((size_t)B * Q * kListedSplits * k * sizeof(int), 256);
}
int knn_sweep_listed(const float* query, const float* ref, int B, int Q, int R, int k, int64_t* idx, double* dist,
                     const int* qlist, const int* qcount, void* part, cudaStream_t stream) {
    const int tiles = (R + kKnnTile - 1) / kKnnTile;
    int splits = tiles < kListedSplits ? tiles : kListedSplits;
    const int tps = (tiles + splits - 1) / splits;
    splits = (tiles + tps - 1) / tps;
    while (splits > 1 && (R - (splits - 1) * tps * kKnnTile) < k) splits = 1;   // every slice must hold at least k points
    double* part_d = (double*)part;
    int* part_i = (int*)((char*)part + align_up((size_t)B * Q * kListedSplits * k * sizeof(double), 256));
    const int tps_arg = splits > 1 ? tps : tiles;
    dim3 grid((Q + kKnnThreads - 1) / kKnnThreads, B, splits);
    if (k <= 1) knn_kernel<1><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i, qlist, qcount);
    else if (k <= 4) knn_kernel<4><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i, qlist, qcount);
    else if (k <= 9) knn_kernel<9><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i, qlist, qcount);
    else knn_kernel<16><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i, qlist, qcount);
    PCST_CUDA(cudaGetLastError());
    if (splits > 1) {
        knn_merge_listed_kernel<<<dim3((Q + 127) / 128, B), 128, 0, stream>>>(part_d, part_i, Q, splits, k, qlist, qcount, idx, dist);
        PCST_CUDA(cudaGetLastError());
    }
    return PCST_OK;
}

// The exact grid search (knn_grid.cu) takes over when both clouds are large: it beats the sweep on every cloud pair measured
// (profiles/r02/knn_grid.md): 3-NN 90k x 30k 1.36 -> 0.75 ms, 9-NN self query 120k 9.41 -> 1.04 ms, 120k x 120k between two
// scans 6.1 -> 3.3 ms, Gaussian noise clouds 1.51 -> 1.08 ms.  knn.grid = 1 / 2 forces it on / off.
static bool knn_use_grid(int Q, int R, bool self_query = false) {
    (void)self_query;
    const int t = tuning("knn.grid", 0);
    if (t == 1) return true;
    if (t == 2) return false;
    return R >= 16384 && Q >= 16384;
}

}  // namespace pcst

using namespace pcst;

// reference-range slices: 1 when the queries alone fill the GPU, more when they do not
static int knn_splits(int B, int Q, int R) {
    const long ctas = (long)B * ((Q + kKnnThreads - 1) / kKnnThreads);
    const int tiles = (R + kKnnTile - 1) / kKnnTile;
    long s = (2L * num_sms() + ctas - 1) / ctas;
    if (s > tiles) s = tiles;
    if (s > 64) s = 64;
    return s < 1 ? 1 : (int)s;
}

extern "C" size_t pcst_knn_workspace_bytes(int B, int Q, int R, int k) {
    if (B <= 0 || Q <= 0 || R <= 0 || k <= 0) return 0;
    // (a caller cannot say here whether query == ref: size for the grid whenever a self query of this shape would take it)
    if (knn_use_grid(Q, R, Q == R)) return knn_grid_workspace_bytes(B, Q, R, k);
    const int s = knn_splits(B, Q, R);
    if (s == 1) return 0;
    return align_up((size_t)B * Q * s * k * sizeof(double), 256) + align_up((size_t)B * Q * s * k * sizeof(int), 256);
}

extern "C" int pcst_knn_kernel_launches(int B, int Q, int R, int k, int self_query) {
    if (B <= 0 || Q <= 0 || R <= 0 || k <= 0) return 0;
    if (knn_use_grid(Q, R, self_query && Q == R)) return 16;  // bounding box, parameters, count, refine, count, scan, scatter, ring walk, listed sweep (+ memsets)
    return knn_splits(B, Q, R) > 1 ? 2 : 1;
}

extern "C" int pcst_knn_f32(const float* query, const float* ref, int B, int Q, int R, int k, int64_t* idx,
                            double* dist, void* ws, size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(query && ref && idx && dist, "null pointer");
    PCST_CHECK_ARG(B > 0 && Q > 0 && R > 0, "B, Q, R must be positive");
    PCST_CHECK_ARG(B <= 65535, "B must be <= 65535");
    PCST_CHECK_ARG(k >= 1 && k <= kKnnMaxK && k <= R, "k must be in [1, min(16, R)]");
    if (knn_use_grid(Q, R, query == ref && Q == R)) {
        const size_t need_g = knn_grid_workspace_bytes(B, Q, R, k);
        if (!ws || ws_bytes < need_g || ((uintptr_t)ws & 255)) {
            set_error("pcst_knn_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need_g);
            return PCST_ERR_WORKSPACE;
        }
        return knn_grid_run(query, ref, B, Q, R, k, idx, dist, ws, stream);
    }
    int splits = knn_splits(B, Q, R);
    const int tiles = (R + kKnnTile - 1) / kKnnTile;
    const int tps = (tiles + splits - 1) / splits;
    splits = (tiles + tps - 1) / tps;  // no empty slice
    // a slice must hold at least k points so that every partial list is full (the last slice may be shorter)
    while (splits > 1 && (R - (splits - 1) * tps * kKnnTile) < k) splits = 1;
    const size_t need = pcst_knn_workspace_bytes(B, Q, R, k);
    double* part_d = nullptr;
    int* part_i = nullptr;
    if (splits > 1) {
        if (!ws || ws_bytes < need || ((uintptr_t)ws & 255)) {
            set_error("pcst_knn_f32: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
            return PCST_ERR_WORKSPACE;
        }
        part_d = (double*)ws;
        part_i = (int*)((char*)ws + align_up((size_t)B * Q * knn_splits(B, Q, R) * k * sizeof(double), 256));
    }
    const int tps_arg = splits > 1 ? tps : tiles;
    dim3 grid((Q + kKnnThreads - 1) / kKnnThreads, B, splits);
    if (k <= 1) knn_kernel<1><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i);
    else if (k <= 4) knn_kernel<4><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i);
    else if (k <= 9) knn_kernel<9><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i);
    else knn_kernel<16><<<grid, kKnnThreads, 0, stream>>>(query, ref, Q, R, k, idx, dist, splits, tps_arg, part_d, part_i);
    PCST_CUDA(cudaGetLastError());
    if (splits > 1) {
        const long total = (long)B * Q;
        knn_merge_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(part_d, part_i, total, splits, k, idx, dist);
        PCST_CUDA(cudaGetLastError());
    }
    return PCST_OK;
}

extern "C" int pcst_knn_interpolate_f32(const float* feat, const int64_t* idx, const double* dist, int B, int R,
                                        int Q, int k, int C, float* out, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(feat && idx && dist && out, "null pointer");
    PCST_CHECK_ARG(B > 0 && R > 0 && Q > 0 && C > 0, "B, R, Q, C must be positive");
    PCST_CHECK_ARG(k >= 1 && k <= kKnnMaxK, "k must be in [1, 16]");
    dim3 grid((Q + 127) / 128, B);
    knn_interpolate_kernel<<<grid, 128, 0, stream>>>(feat, idx, dist, R, Q, k, C, out);
    return check_cuda(cudaGetLastError(), "knn_interpolate_kernel");
}
