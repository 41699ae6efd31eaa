/*
 * pcst_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
 *
 * A plain-C restatement of the arithmetic of the reference's point-set hot path
 * (wangxy0820/PointCloud_style_transfer).  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library, and only as the
 * checker or as the timed CPU baseline -- never on the product path.
 *
 * Parity pin: the reference ships no tests or golden vectors, so
 * this restatement is pinned against outputs of the reference itself, imported
 * and executed in the build container by oracle/gen_golden.py (small fixtures
 * committed under tests/golden/) and by oracle/pin_against_reference.py (full
 * 120k-point sizes, report in oracle/PINNING.md).
 *
 * Build: gcc -O3 -mavx2 -mfma -ffp-contract=off -fopenmp -shared -fPIC (see Makefile).
 * -ffp-contract=off is REQUIRED: the reference's roundings are reproduced with
 * explicit fmaf() where MKL fuses and plain mul/add where ATen does not.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- arithmetic atoms --------------------------------------------------- */

/* torch.matmul / torch.bmm with K=3 on CPU (MKL sgemm): a sequential FMA chain
 * over k, first term a plain product. */
static inline float dot3_chain(const float* a, const float* b) {
    return fmaf(a[2], b[2], fmaf(a[1], b[1], a[0] * b[0]));
}

/* torch.sum(x ** 2, -1) over the size-3 axis: (x*x + y*y) + z*z, no FMA
 * (models/pointnet2_encoder.py:13-14, models/losses.py:24-25). */
static inline float norm3_sq(const float* a) {
    float xx = a[0] * a[0], yy = a[1] * a[1], zz = a[2] * a[2];
    return (xx + yy) + zz;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- square_distance: models/pointnet2_encoder.py:8-15 -------------------
 * dist = -2 * matmul(src, dst^T); dist += |src|^2; dist += |dst|^2  (that order). */
void oracle_square_distance(const float* src, const float* dst, int B, int N, int M,
                            float* out) {
    for (int b = 0; b < B; ++b) {
        const float* s = src + (size_t)b * N * 3;
        const float* d = dst + (size_t)b * M * 3;
        float* o = out + (size_t)b * N * M;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i) {
            float sn = norm3_sq(s + 3 * i);
            for (int j = 0; j < M; ++j) {
                float dn = norm3_sq(d + 3 * j);
                float t = -2.0f * dot3_chain(s + 3 * i, d + 3 * j);
                t = t + sn;
                t = t + dn;
                o[(size_t)i * M + j] = t;
            }
        }
    }
}

/* ---- farthest_point_sample: models/pointnet2_encoder.py:30-45 ------------
 * distance0 = 1e10; slot i receives the index chosen BEFORE update i;
 * d = ((dx*dx)+(dy*dy))+(dz*dz); distance = d < distance ? d : distance;
 * farthest = argmax(distance), first occurrence among equal maxima.
 * The start index is an input (the reference draws it with torch.randint on the
 * CPU generator, :36 -- that draw stays in the Python caller). */
void oracle_fps(const float* xyz, int B, int N, int npoint, const int64_t* start,
                int64_t* out) {
    float* dist = (float*)malloc(sizeof(float) * (size_t)N);
    int nchunk = oracle_num_threads();
    if (nchunk > 64) nchunk = 64;
    if (N < 16384) nchunk = 1; /* not worth a parallel region per iteration */
    float bestv[64];
    int64_t besti[64];
    for (int b = 0; b < B; ++b) {
        const float* p = xyz + (size_t)b * N * 3;
        for (int j = 0; j < N; ++j) dist[j] = 1e10f;
        int64_t far = start[b];
        for (int i = 0; i < npoint; ++i) {
            out[(size_t)b * npoint + i] = far;
            const float cx = p[3 * far], cy = p[3 * far + 1], cz = p[3 * far + 2];
#pragma omp parallel for schedule(static) num_threads(nchunk)
            for (int c = 0; c < nchunk; ++c) {
                int lo = (int)((int64_t)N * c / nchunk), hi = (int)((int64_t)N * (c + 1) / nchunk);
                float bv = -1.0f;
                int64_t bi = lo;
                for (int j = lo; j < hi; ++j) {
                    float dx = p[3 * j] - cx, dy = p[3 * j + 1] - cy, dz = p[3 * j + 2] - cz;
                    float d = (dx * dx + dy * dy) + dz * dz;
                    float cur = dist[j];
                    if (d < cur) { cur = d; dist[j] = d; }
                    if (cur > bv) { bv = cur; bi = j; }
                }
                bestv[c] = bv;
                besti[c] = bi;
            }
            float bv = bestv[0];
            int64_t bi = besti[0];
            for (int c = 1; c < nchunk; ++c)
                if (bestv[c] > bv) { bv = bestv[c]; bi = besti[c]; }
            far = bi;
        }
    }
    free(dist);
}

/* ---- query_ball_point: models/pointnet2_encoder.py:47-59 -----------------
 * keep index j iff NOT (sqrdist > fp32(radius^2)), sqrdist as in square_distance
 * with src = new_xyz (queries), dst = xyz; ascending index order; first nsample;
 * short rows padded with the row's first index; empty rows = N everywhere.
 * (The reference sorts a masked arange; scanning in index order with an early
 * exit yields the same rows.) */
void oracle_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S,
                       float radius_sq, int nsample, int64_t* out) {
    for (int b = 0; b < B; ++b) {
        const float* p = xyz + (size_t)b * N * 3;
        const float* q = new_xyz + (size_t)b * S * 3;
#pragma omp parallel for schedule(dynamic, 4)
        for (int s = 0; s < S; ++s) {
            int64_t* row = out + ((size_t)b * S + s) * nsample;
            float qn = norm3_sq(q + 3 * s);
            int cnt = 0;
            for (int j = 0; j < N && cnt < nsample; ++j) {
                float t = -2.0f * dot3_chain(q + 3 * s, p + 3 * j);
                t = t + qn;
                t = t + norm3_sq(p + 3 * j);
                if (!(t > radius_sq)) row[cnt++] = j;
            }
            int64_t first = cnt > 0 ? row[0] : (int64_t)N;
            for (int k = cnt; k < nsample; ++k) row[k] = first;
        }
    }
}

/* ---- NN-min reduction ----------------------------------------------------
 * form 0: Chamfer loss, models/losses.py:36-41,53-58:
 *         D = clamp((|a_i|^2 + |b_j|^2) + (-2 * dot_chain(a_i, b_j)), min=0)   (squared)
 * form 1: metric, evaluation/metrics.py:32 (torch.cdist p=2, mm path of ATen
 *         _euclidean_dist): K=5 sgemm row [-2a, |a|^2, 1] . [b, 1, |b|^2] as an FMA
 *         chain, clamp_min(0), sqrt.  The min is taken on the clamped square and the
 *         sqrt applied once (sqrt is monotone and correctly rounded).
 *         torch's vectorised CPU sqrt is not correctly rounded (0.7% of values are 1 ulp
 *         off sqrtf), so form-1/2 values are pinned to the reference within 1 ulp only.
 * form 2: the same cdist matrix reduced along its other axis (metrics.py:40,
 *         min over dim=1): rows are cdist's x2 points, candidates its x1 points, i.e.
 *         D2(i,j) = form1(b_j, a_i) -- the K=5 chain is not symmetric in its last two terms.
 * rowmin[i] = min_j D(i,j); rowarg[i] = first j attaining it (may be NULL).
 * For form 0 the second direction is the same call with roles swapped: D(j,i) of the
 * swapped call is bit-identical to D(i,j) (Appendix A.3). */
static inline float pair_form0(const float* a, float an, const float* b, float bn) {
    float t = an + bn;
    float m = -2.0f * dot3_chain(a, b);
    return t + m;
}
static inline float pair_form1(const float* a, float an, const float* b, float bn) {
    float r = (-2.0f * a[0]) * b[0];
    r = fmaf(-2.0f * a[1], b[1], r);
    r = fmaf(-2.0f * a[2], b[2], r);
    r = fmaf(an, 1.0f, r);
    r = fmaf(1.0f, bn, r);
    return r;
}

void oracle_nn_min(const float* a, const float* b, int B, int N, int M, int form,
                   float* rowmin, int64_t* rowarg) {
    float* bn = (float*)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
    for (int bb = 0; bb < B; ++bb) {
        const float* pa = a + (size_t)bb * N * 3;
        const float* pb = b + (size_t)bb * M * 3;
        for (int j = 0; j < M; ++j) bn[j] = norm3_sq(pb + 3 * j);
#pragma omp parallel for schedule(static)
        for (int i = 0; i < N; ++i) {
            float an = norm3_sq(pa + 3 * i);
            float best = INFINITY;
            int64_t arg = 0;
            if (form == 0) {
                for (int j = 0; j < M; ++j) {
                    float d = pair_form0(pa + 3 * i, an, pb + 3 * j, bn[j]);
                    d = d < 0.0f ? 0.0f : d;
                    if (d < best) { best = d; arg = j; }
                }
            } else if (form == 1) {
                for (int j = 0; j < M; ++j) {
                    float d = pair_form1(pa + 3 * i, an, pb + 3 * j, bn[j]);
                    d = d < 0.0f ? 0.0f : d;
                    if (d < best) { best = d; arg = j; }
                }
                best = sqrtf(best);
            } else {
                for (int j = 0; j < M; ++j) {
                    float d = pair_form1(pb + 3 * j, bn[j], pa + 3 * i, an);
                    d = d < 0.0f ? 0.0f : d;
                    if (d < best) { best = d; arg = j; }
                }
                best = sqrtf(best);
            }
            rowmin[(size_t)bb * N + i] = best;
            if (rowarg) rowarg[(size_t)bb * N + i] = arg;
        }
    }
    free(bn);
}

/* ---- kNN: sklearn NearestNeighbors(n_neighbors=k).kneighbors --------------
 * models/diffusion_model.py:146-147, evaluation/metrics.py:126-127,152-153,
 * data/preprocessing.py:122-123.  Inputs are up-cast to fp64; reduced distance
 * r = ((dx*dx) + (dy*dy)) + (dz*dz) in fp64; the k smallest, ascending, ties to
 * the lower index; returned distance sqrt(r).
 * sklearn 1.9.0 (reference pins >=1.3.2) is a third-party dependency absent from
 * /root/reference; this is its published brute-force-equivalent definition. */
void oracle_knn(const float* query, const float* ref, int B, int Q, int R, int k,
                int64_t* idx, double* dist) {
    for (int b = 0; b < B; ++b) {
        const float* q = query + (size_t)b * Q * 3;
        const float* r = ref + (size_t)b * R * 3;
#pragma omp parallel for schedule(static)
        for (int i = 0; i < Q; ++i) {
            double bd[64];
            int64_t bi[64];
            int n = 0;
            double qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
            for (int j = 0; j < R; ++j) {
                double dx = qx - (double)r[3 * j], dy = qy - (double)r[3 * j + 1],
                       dz = qz - (double)r[3 * j + 2];
                double d = (dx * dx + dy * dy) + dz * dz;
                if (n == k && !(d < bd[k - 1])) continue;
                int pos = n < k ? n : k - 1;
                while (pos > 0 && d < bd[pos - 1]) {
                    bd[pos] = bd[pos - 1];
                    bi[pos] = bi[pos - 1];
                    --pos;
                }
                bd[pos] = d;
                bi[pos] = j;
                if (n < k) ++n;
            }
            for (int t = 0; t < k; ++t) {
                idx[((size_t)b * Q + i) * k + t] = bi[t];
                dist[((size_t)b * Q + i) * k + t] = sqrt(bd[t]);
            }
        }
    }
}

/* ---- SetAbstraction.apply_mlp: models/pointnet2_encoder.py:106-112 (eval-mode BatchNorm) ----------
 * x [G*K, C0] (row = one grouped point, channels last), three layers relu(bn(conv1x1(.))), then the
 * max over each group's K rows (torch.max(points, 3)[0], :112) -> out [G, C3].
 * wt[l] is the conv weight TRANSPOSED to [Cin_l, Cout_l] (so that the inner loop runs over output
 * channels and vectorises without re-association); bias/gamma/beta/mean/var [Cout_l].
 * Arithmetic: z = sum_ci x[ci] * w[co,ci] (sequential over ci, separate multiply and add) + bias;
 * y = (z - mean) / sqrt(var + eps) * gamma + beta; relu.  Differs from MKL's blocked summation only
 * in rounding (tests: rtol 1e-4).  Used by bench.py's CPU baseline so that the whole encoder runs in
 * ONE OpenMP runtime (no OpenMP x BLAS thread-pool oversubscription). */
void oracle_apply_mlp3(const float* x, long G, int K, int C0, const int* cout, const float* const* wt,
                       const float* const* bias, const float* const* gamma, const float* const* beta,
                       const float* const* mean, const float* const* var, float eps, float* out) {
    const int C1 = cout[0], C2 = cout[1], C3 = cout[2];
    int cmax = C1 > C2 ? C1 : C2;
    if (C3 > cmax) cmax = C3;
    if (C0 > cmax) cmax = C0;
    float* inv[3];
    for (int l = 0; l < 3; ++l) {
        inv[l] = (float*)malloc(sizeof(float) * (size_t)cout[l]);
        for (int c = 0; c < cout[l]; ++c) inv[l][c] = sqrtf(var[l][c] + eps);
    }
#pragma omp parallel
    {
        float* a = (float*)malloc(sizeof(float) * (size_t)cmax);
        float* z = (float*)malloc(sizeof(float) * (size_t)cmax);
#pragma omp for schedule(static)
        for (long g = 0; g < G; ++g) {
            float* o = out + (size_t)g * C3;
            for (int k = 0; k < K; ++k) {
                const float* row = x + ((size_t)g * K + k) * C0;
                int cin = C0;
                for (int c = 0; c < C0; ++c) a[c] = row[c];
                for (int l = 0; l < 3; ++l) {
                    const int co_n = cout[l];
                    const float* w = wt[l];
                    for (int co = 0; co < co_n; ++co) z[co] = 0.f;
                    for (int ci = 0; ci < cin; ++ci) {
                        const float xv = a[ci];
                        const float* wr = w + (size_t)ci * co_n;
                        for (int co = 0; co < co_n; ++co) z[co] += xv * wr[co];
                    }
                    for (int co = 0; co < co_n; ++co) {
                        float y = (z[co] + bias[l][co] - mean[l][co]) / inv[l][co] * gamma[l][co] + beta[l][co];
                        a[co] = y > 0.f ? y : 0.f;
                    }
                    cin = co_n;
                }
                if (k == 0) for (int c = 0; c < C3; ++c) o[c] = a[c];
                else for (int c = 0; c < C3; ++c) o[c] = a[c] > o[c] ? a[c] : o[c];
            }
        }
        free(a);
        free(z);
    }
    for (int l = 0; l < 3; ++l) free(inv[l]);
}
