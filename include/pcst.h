/*
 * pcst.h -- C ABI of libpcst.so: B200 (sm_100a) kernels for the point-set hot path of
 * wangxy0820/PointCloud_style_transfer.
 *
 * The reference has no FFI layer: its boundary is the Python symbol surface of
 *   models/pointnet2_encoder.py, models/losses.py, evaluation/metrics.py and
 *   models/diffusion_model.py::HierarchicalProcessor.upsample_knn      (SURVEY.md §8(b)).
 * Each entry point below names the reference function (file:line, relative to the reference
 * root) whose arithmetic it replaces.  The Python package pointcloud_style_transfer_b200 binds
 * these symbols with ctypes and re-exposes them as torch custom ops (pcst::*) behind the
 * reference's own function / class names; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - Every function returns an int status: 0 = PCST_OK, negative = error.  Nothing throws across
 *    the ABI.  pcst_last_error() returns a thread-local, human-readable message for the last
 *    non-zero status on the calling thread.
 *  - All data pointers are DEVICE pointers on the current CUDA device, contiguous in the stated
 *    row-major layout.  Floating data are fp32 unless stated, indices are int64 (the reference's
 *    dtypes).  The caller owns every buffer: the library never allocates, frees or retains
 *    device memory.  Scratch space is passed in as (ws, ws_bytes); query the size with the
 *    matching *_workspace_bytes() function.  ws must be 256-byte aligned.
 *  - Kernels are launched on the cudaStream_t passed as `stream` (a void* here so that the header
 *    needs no CUDA include); calls are stream-ordered, asynchronous and re-entrant.  There are no
 *    implicit device synchronisations and no host<->device copies.  The only process-wide mutable
 *    state is the tuning-knob table (pcst_set_tuning, mutex-guarded) and the profiling hook
 *    pcst_sa_mlp_set_probe, both of which exist for measurements and default to "off".
 *  - There is no CPU fallback.  On a device that is not compute capability 10.x every compute
 *    entry point returns PCST_ERR_UNSUPPORTED.
 */
#ifndef PCST_H_
#define PCST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCST_OK 0
#define PCST_ERR_INVALID (-1)     /* bad argument (null pointer, non-positive size, k out of range ...) */
#define PCST_ERR_UNSUPPORTED (-2) /* shape / device outside what the kernels support */
#define PCST_ERR_CUDA (-3)        /* a CUDA runtime call or launch failed; see pcst_last_error() */
#define PCST_ERR_WORKSPACE (-4)   /* ws_bytes smaller than *_workspace_bytes() or ws misaligned */

typedef void* pcst_stream_t; /* cudaStream_t */

const char* pcst_version(void);
const char* pcst_last_error(void);
/* 0 if the current device can run the kernels (compute capability 10.x), else PCST_ERR_UNSUPPORTED. */
int pcst_device_check(void);
/* Internal tuning knobs (kernel variant selection for benchmarking sweeps); unknown keys -> PCST_ERR_INVALID. */
int pcst_set_tuning(const char* key, int value);
int pcst_get_tuning(const char* key, int* value);

/* Warm the L2 cache with [ptr, ptr + bytes) (one prefetch per 128-byte line; returns immediately, stream-ordered).
 * The encoder uses it to pull the packed MLP weights of all three stages into the 126 MB L2 on a parallel
 * stream while the first FPS runs on 16 of the 148 SMs; no reference counterpart (a scheduling aid, not arithmetic). */
int pcst_l2_prefetch(const void* ptr, size_t bytes, pcst_stream_t stream);

/* FP32-pipe calibration for the roofline of the distance kernels: launches independent packed-FMA chains on every SM
 * and returns the number of flop the launch performs (0 on error); the caller times it with CUDA events.
 * scratch: one float of device memory (never written in practice). */
long long pcst_fp32_probe(int iters, float* scratch, pcst_stream_t stream);

/* ---- farthest_point_sample: models/pointnet2_encoder.py:30-45 --------------------------------
 * xyz [B,N,3]; start [B] = the start index per cloud (the reference draws it with torch.randint
 * on the CPU generator, :36 -- that draw stays in the caller); out [B,npoint] int64.
 * new_xyz (optional, may be NULL) [B,npoint,3] receives xyz[b, out[b,i], :], i.e. the
 * index_points(xyz, fps_idx) of SetAbstraction.forward (:92), at no extra cost.
 * One persistent thread-block cluster per cloud; points and running distances stay on chip. */
size_t pcst_fps_workspace_bytes(int B, int N, int npoint);
/* Number of N-point clouds the current device processes concurrently (one cluster each); larger batches run in
 * waves.  0 on error.  A scheduling hint for callers that batch scans; no reference counterpart. */
int pcst_fps_max_concurrent_clouds(int N);
int pcst_fps_f32(const float* xyz, int B, int N, int npoint, const int64_t* start, int64_t* out,
                 float* new_xyz, void* ws, size_t ws_bytes, pcst_stream_t stream);

/* ---- query_ball_point: models/pointnet2_encoder.py:47-59 --------------------------------------
 * xyz [B,N,3], new_xyz [B,S,3] -> out [B,S,nsample] int64: the first nsample indices j (ascending)
 * with NOT(sqrdist(new_xyz_s, xyz_j) > radius_sq), sqrdist in the reference's expanded fp32 form;
 * short rows padded with the row's first index, empty rows filled with N.
 * radius_sq must be fp32(radius ** 2) exactly as torch's type promotion produces it (:54).
 * Requires 1 <= nsample <= N (the reference raises IndexError for nsample > N, :58). */
size_t pcst_ball_query_workspace_bytes(int B, int N, int S);
int pcst_ball_query_f32(const float* xyz, const float* new_xyz, int B, int N, int S, float radius_sq,
                        int nsample, int64_t* out, void* ws, size_t ws_bytes, pcst_stream_t stream);

/* ---- square_distance: models/pointnet2_encoder.py:8-15 ----------------------------------------
 * src [B,N,3], dst [B,M,3] -> out [B,N,M] = ((-2 * dot) + |src|^2) + |dst|^2, bit-exact with the
 * reference's fp32 CPU result.  (The fused kernels never materialise this matrix; the entry point
 * exists because the function is part of the reference's public surface.) */
int pcst_square_distance_f32(const float* src, const float* dst, int B, int N, int M, float* out,
                             pcst_stream_t stream);

/* ---- index_points: models/pointnet2_encoder.py:17-28 -------------------------------------------
 * points [B,N,C], idx [B,S] (any trailing index shape flattened to S) -> out [B,S,C];
 * indices are clamped to [0, N-1] (:26).  _bwd accumulates grad_out [B,S,C] into grad_points
 * [B,N,C], which the caller has zero-initialised (autograd of the advanced indexing at :27). */
int pcst_index_points_f32(const float* points, const int64_t* idx, int B, int N, int C, int S,
                          float* out, pcst_stream_t stream);
int pcst_index_points_bwd_f32(const float* grad_out, const int64_t* idx, int B, int N, int C, int S,
                              float* grad_points, pcst_stream_t stream);

/* ---- grouping of SetAbstraction.forward: models/pointnet2_encoder.py:94-101 --------------------
 * out [B,S,K,3+D] = cat([xyz[idx] - new_xyz[:, :, None, :], feats[idx]], -1); feats may be NULL (D=0).
 * xyz [B,N,3], feats [B,N,D], new_xyz [B,S,3], idx [B,S,K] (clamped like index_points). */
int pcst_group_f32(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx,
                   int B, int N, int S, int K, int D, float* out, pcst_stream_t stream);

/* ---- SetAbstraction.apply_mlp: models/pointnet2_encoder.py:106-112 (eval-mode BatchNorm) --------
 * Fused grouping gather (:94-101) + 3 x relu(bn(conv1x1(.))) + max over the K samples of each group.
 *
 * Parameters are PACKED ONCE per parameter version and reused by every forward:
 *   layer l (l = 0..2): weight w[l] [Cout_l, Cin_l] fp32 row-major (the Conv2d weight [Cout,Cin,1,1]),
 *   scale[l], shift[l] [Cout_l]: y = relu(scale * (w . x) + shift), i.e. conv bias and eval-mode BN
 *   folded by the caller: scale = gamma / sqrt(var + eps), shift = (bias - mean) * scale + beta.
 *   Cin_0 = 3 + D, Cin_l = Cout_{l-1}.  Supported: Cout_l a multiple of 32, Cout_l <= 1024.
 * precision: 0 = fp32 CUDA-core path; 1 = bf16 tcgen05/TMEM tensor-core path (fp32 accumulate; needs
 *   Cout_0 <= 256 and Cout_1, Cout_2 <= 512, otherwise the fp32 path is used).  pcst_sa_mlp_pack_f32
 *   writes the layout the chosen path wants (tensor cores: bf16 K-major UMMA operand blocks + fp32
 *   scale/shift) into `packed` (caller-owned, 256-byte aligned, pcst_sa_mlp_packed_bytes() bytes).
 *
 * pcst_sa_mlp_max_f32:
 *   xyz [B,N,3], feats [B,N,D] or NULL, new_xyz [B,S,3] or NULL, idx [B,S,K] int64 (clamped) or NULL.
 *   idx == NULL means group_all (:81-89): S = 1, K = N, the "group" is the whole cloud in order and
 *   no centroid is subtracted.  cout/precision/D must be the ones `packed` was built with.
 * out [B, S, Cout_2] POINT-major; the reference's channel-first [B, Cout_2, S] is out.permute(0, 2, 1)
 * (the next stage consumes the point-major form, models/pointnet2_encoder.py:128-129). */
typedef struct {
    const float* w[3];
    const float* scale[3];
    const float* shift[3];
    int cout[3];
} pcst_mlp3_t;
/* cluster: CTAs per 128-row tile on the tensor-core path (1, 2, 4 or 8): stages with few rows (the group_all stage is
 * one tile per scan) split every layer's output channels over a thread-block cluster.  pcst_sa_mlp_pick_cluster returns
 * the value to use for a shape (1 for the fp32 path); the packed layout depends on it, so callers cache one blob per
 * (precision, cluster). */
int pcst_sa_mlp_pick_cluster(int B, int S, int K, int D, const int* cout /*[3]*/, int precision);
size_t pcst_sa_mlp_packed_bytes(int D, const int* cout /*[3]*/, int precision, int cluster);
int pcst_sa_mlp_pack_f32(const pcst_mlp3_t* mlp, int D, int precision, int cluster, void* packed, size_t packed_bytes,
                         pcst_stream_t stream);
size_t pcst_sa_mlp_max_workspace_bytes(int B, int N, int S, int K, int D, const int* cout /*[3]*/, int precision);
/* __global__ launches of one pcst_sa_mlp_max_f32 call with these sizes (bookkeeping for launch counts; 0 = bad sizes) */
int pcst_sa_mlp_max_kernel_launches(int B, int N, int S, int K, int D, const int* cout /*[3]*/, int precision);
int pcst_sa_mlp_max_f32(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx,
                        int B, int N, int S, int K, int D, const int* cout /*[3]*/, int precision, int cluster,
                        const void* packed, float* out, void* ws, size_t ws_bytes, pcst_stream_t stream);
/* Profiling aid (no reference counterpart): while `stamps` is non-NULL every tensor-core launch records, for each of its
 * first `tiles` 128-row tiles, 16 uint64 in device memory: SM-clock stamps [0] tile start, then per epilogue-bearing step
 * [2i+1] accumulator ready (gather / previous epilogue + the step's MMAs done) and [2i+2] epilogue done; [15] = the SM id.
 * NULL switches it off. */
void pcst_sa_mlp_set_probe(unsigned long long* stamps /*[tiles,16] device*/, int tiles);

/* ---- SetAbstraction.apply_mlp in TRAINING mode: models/pointnet2_encoder.py:74,106-112 under module.train() --------
 * (the step of training/trainer.py:78-117 that the reference runs through cuDNN / cuBLAS + autograd)
 * Same fused grouping gather as pcst_sa_mlp_max_f32, but nn.BatchNorm2d uses BATCH statistics over all B*S*K rows
 * (biased variance, eps inside the square root), updates running_mean / running_var (momentum, unbiased variance) and
 * num_batches_tracked in place, and the backward returns the gradients autograd would.
 * bf16 tcgen05 GEMMs (forward, dgrad, wgrad) with fp32 accumulation, fp64 batch statistics; the pre-BatchNorm
 * activations between the layer passes are fp32 [rows, C] in the caller-owned `saved` blob, which the backward reads
 * (it is what autograd keeps).
 * precision: 1 = bf16 operands (one MMA per K step; the autocast mode of BASELINE config 4);
 *            0 = split operands, x = bf16(x) + bf16(x - bf16(x)), three MMAs per K step: fp32-faithful GEMMs, results
 *                track the reference's fp32 autograd (rtol 2e-3 in the tests).  Forward and backward must use the same value.
 * Supported: Cout_l a multiple of 16 and <= 512, 3 + D <= 512.
 *   w[l] [Cout_l, Cin_l] fp32 (the Conv2d weight), bias[l], gamma[l], beta[l] [Cout_l];
 *   running_mean[l], running_var[l], num_batches_tracked[l]: may be NULL (track_running_stats=False).
 * forward:  out [B, S, Cout_2] POINT-major (like pcst_sa_mlp_max_f32).
 * backward: grad_out [B, S, Cout_2]; grads->w[l] [Cout_l, Cin_l], bias[l] (exactly 0: a bias in front of BatchNorm
 *           has no gradient), gamma[l], beta[l] -- any of them may be NULL to skip it;
 *           grad_grouped (may be NULL) [B, S, K, 3 + D] = gradient w.r.t. the grouped input
 *           cat([xyz[idx] - new_xyz, feats[idx]], -1) in the reference's channel order (:99); the caller scatters it
 *           with pcst_index_points_bwd_f32. */
typedef struct {
    const float* w[3];
    const float* bias[3];
    const float* gamma[3];
    const float* beta[3];
    float* running_mean[3];
    float* running_var[3];
    int64_t* num_batches_tracked[3];
    int cout[3];
    float eps;      /* nn.BatchNorm2d default 1e-5 */
    float momentum; /* nn.BatchNorm2d default 0.1 */
} pcst_mlp3_train_t;
typedef struct {
    float* w[3];
    float* bias[3];
    float* gamma[3];
    float* beta[3];
} pcst_mlp3_grads_t;
size_t pcst_sa_mlp_train_saved_bytes(int B, int S, int K, int D, const int* cout /*[3]*/);
size_t pcst_sa_mlp_train_workspace_bytes(int B, int S, int K, int D, const int* cout /*[3]*/, int precision, int backward);
int pcst_sa_mlp_max_bnstats_bf16(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B, int N,
                                 int S, int K, int D, const pcst_mlp3_train_t* mlp, int precision, float* out, void* saved,
                                 size_t saved_bytes, void* ws, size_t ws_bytes, pcst_stream_t stream);
int pcst_sa_mlp_max_bwd_bf16(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B, int N,
                             int S, int K, int D, const pcst_mlp3_train_t* mlp, int precision, const void* saved,
                             size_t saved_bytes, const float* grad_out, const pcst_mlp3_grads_t* grads, float* grad_grouped,
                             void* ws, size_t ws_bytes, pcst_stream_t stream);

/* ---- nearest-neighbour minimum reduction ------------------------------------------------------
 * a [B,N,3], b [B,M,3] -> rowmin [B,N] = min_j D(a_i, b_j), rowarg [B,N] (optional, may be NULL) =
 * the lowest j attaining it.
 *  form 0: models/losses.py:36-41,53-58 -- D = clamp((|a|^2 + |b|^2) + (-2 * dot), min=0), squared
 *          distances; bit-exact with the reference's fp32 CPU per-point minima.  The second
 *          direction of the Chamfer loss is the same call with a and b swapped.
 *  form 1: evaluation/metrics.py:32 -- torch.cdist(a, b, p=2) row minima (Euclidean; ATen's
 *          mm path [-2a,|a|^2,1].[b,1,|b|^2], clamp_min(0), sqrt), a = cdist's x1.
 *  form 2: column minima of the same cdist matrix (evaluation/metrics.py:40): a = cdist's x2,
 *          b = cdist's x1. */
size_t pcst_nn_min_workspace_bytes(int B, int N, int M);
int pcst_nn_min_f32(const float* a, const float* b, int B, int N, int M, int form, float* rowmin,
                    int64_t* rowarg, void* ws, size_t ws_bytes, pcst_stream_t stream);

/* Both directions of the same pair matrix in ONE sweep (each pair evaluated once): the matrix of the second
 * direction is bit-for-bit the transpose of the first, for the loss (models/losses.py:36-41 vs 53-58) and for
 * torch.cdist (evaluation/metrics.py:32: dist.min(dim=2) / dist.min(dim=1)).
 *  form 0: rowmin [B,N] = pcst_nn_min_f32(a, b, form 0), colmin [B,M] = pcst_nn_min_f32(b, a, form 0);
 *  form 1: a = cdist's x1: rowmin = form 1 of (a, b), colmin = form 2 of (b, a).
 * Values are identical to the two one-directional calls; no argmin (use pcst_nn_min_f32 for backward). */
size_t pcst_nn_min_pair_workspace_bytes(int B, int N, int M);
int pcst_nn_min_pair_f32(const float* a, const float* b, int B, int N, int M, int form, float* rowmin,
                         float* colmin, void* ws, size_t ws_bytes, pcst_stream_t stream);

/* The same single sweep (loss form) WITH both argmins, for the backward of the Chamfer loss:
 * rowarg [B,N] = lowest j attaining rowmin, colarg [B,M] = lowest i attaining colmin -- identical to the outputs of
 * pcst_nn_min_f32(a, b, ..., rowarg) and pcst_nn_min_f32(b, a, ..., rowarg).  The sweep remembers only which
 * 32-candidate / 256-row block held each minimum; two small fix-up kernels recover the exact first index. */
size_t pcst_nn_min_pair_arg_workspace_bytes(int B, int N, int M);
int pcst_nn_min_pair_arg_f32(const float* a, const float* b, int B, int N, int M, float* rowmin, int64_t* rowarg,
                             float* colmin, int64_t* colarg, void* ws, size_t ws_bytes, pcst_stream_t stream);

/* Backward of chamfer_distance_chunked_optimized (autograd of models/losses.py:24-61).
 * pred [B,N,3], target [B,M,3]; arg_pt [B,N] / arg_tp [B,M] = the argmins returned by pcst_nn_min_f32
 * (form 0) for pred->target / target->pred; grad_out [B] = dL/d(chamfer[b]).
 * grad_pred [B,N,3], grad_target [B,M,3] are OVERWRITTEN (zeroed inside, then accumulated). */
int pcst_chamfer_bwd_f32(const float* pred, const float* target, const int64_t* arg_pt,
                         const int64_t* arg_tp, const float* grad_out, int B, int N, int M,
                         float* grad_pred, float* grad_target, pcst_stream_t stream);

/* ---- query-sharded Chamfer (multi-GPU): the kernels either side of its single collective ----------------
 * Rank r sweeps its [n_r x M] tile of the pair matrix with pcst_nn_min_pair_f32 (complete row minima of its queries, partial
 * column minima of all M targets).  _pack writes payload [B, P] with P = pcst_chamfer_shard_payload_floats(M) = M + 256:
 * colmin | 128 fp64 partial row sums (float pairs); the ranks all-gather the payloads ([G, B, P]); _finish forms
 * out [B] = sum_r rowsum_r / n_total + sum_m min_r colmin_r[m] / M (x 0.5 for form 1, the metric), fp64 sums added in a
 * fixed order: identical on every rank.  ws: B * 128 doubles, 256-byte aligned.
 * Replaces all_reduce(MIN) + all_reduce(SUM) + host-side reductions around models/losses.py:61 / evaluation/metrics.py:42. */
int pcst_chamfer_shard_payload_floats(int M);
int pcst_chamfer_shard_pack_f32(const float* rowmin, const float* colmin, int B, int n, int M, float* payload,
                                pcst_stream_t stream);
int pcst_chamfer_shard_finish_f32(const float* gathered, int G, int B, int M, long long n_total, int form, float* out, void* ws,
                                  size_t ws_bytes, pcst_stream_t stream);

/* ---- k nearest neighbours: sklearn NearestNeighbors(n_neighbors=k).kneighbors ------------------
 * call sites models/diffusion_model.py:146-147, evaluation/metrics.py:126-127,152-153.
 * query [B,Q,3], ref [B,R,3] fp32 -> idx [B,Q,k] int64, dist [B,Q,k] fp64 ascending (Euclidean);
 * distances are evaluated in fp64 as sqrt(((dx*dx)+(dy*dy))+(dz*dz)), ties to the lower index.
 * 1 <= k <= 16, k <= R.  Large self queries (query == ref, >= 16384 points) are searched through a three-level uniform grid
 * over the references (counting sort by cell, ring walk with an exact stopping bound: the reference's own search is a
 * kd-tree); results are identical to the sweep's. */
size_t pcst_knn_workspace_bytes(int B, int Q, int R, int k);
/* How many kernels pcst_knn_f32 launches for this shape (1-2: brute-force sweep (+ merge); 10: exact multi-level grid search
 * over the reference cloud, the default for large self queries (query == ref) -- same results bit for bit).  For launch
 * accounting. */
int pcst_knn_kernel_launches(int B, int Q, int R, int k, int self_query);
int pcst_knn_f32(const float* query, const float* ref, int B, int Q, int R, int k, int64_t* idx,
                 double* dist, void* ws, size_t ws_bytes, pcst_stream_t stream);

/* ---- inverse-distance interpolation of HierarchicalProcessor.upsample_knn ----------------------
 * models/diffusion_model.py:148-150: w = 1/(dist + 1e-8), w /= sum_k w, out = sum_k w_k * feat[idx_k],
 * evaluated in fp64 and rounded to fp32 once.  feat [B,R,C], idx/dist [B,Q,k] -> out [B,Q,C]. */
int pcst_knn_interpolate_f32(const float* feat, const int64_t* idx, const double* dist, int B, int R,
                             int Q, int k, int C, float* out, pcst_stream_t stream);

/* ---- NoisePredictor.forward: models/diffusion_model.py:38-61 ----------------------------------
 * The per-point denoiser MLP (3 -> 128 -> 256 -> F, + time / style conditioning, nblocks residual blocks F -> 2F -> F,
 * F -> 256 -> 128 -> 3) as ONE fused tcgen05 kernel: activations in shared memory / TMEM, the residual stream as the fp32
 * TMEM accumulator, weights packed once.  Inference (eval mode: Dropout is the identity).
 * All weights are the nn.Linear tensors, row-major [out, in] fp32; pe = point_encoder.{0,2,4}, blk_w1 / blk_w2 =
 * layers.{i}.{0,2}, out = output_mlp.{0,2,4}; time_w [F, time_dim], style_w [F, F].
 * Supported: feature_dim a multiple of 16 in [16, 256], time_dim even, nblocks <= 8.
 * points [B,N,3], timestep [B] int64, style [B,F] (the style feature AFTER any CFG masking) -> out [B,N,3].
 * bf16 operands, fp32 accumulation: within rtol 2e-2 of the reference's fp32 module. */
typedef struct {
    const float* pe_w[3];
    const float* pe_b[3];
    const float* time_w;
    const float* time_b;
    const float* style_w;
    const float* style_b;
    const float* blk_w1[8];
    const float* blk_b1[8];
    const float* blk_w2[8];
    const float* blk_b2[8];
    const float* out_w[3];
    const float* out_b[3];
    int feature_dim, time_dim, nblocks;
} pcst_noise_mlp_t;
size_t pcst_noise_predictor_packed_bytes(int feature_dim, int time_dim, int nblocks);
/* __global__ launches of one pcst_noise_predictor_pack_f32 call (0 = unsupported sizes); bookkeeping for launch counts */
int pcst_noise_predictor_pack_launches(int feature_dim, int time_dim, int nblocks);
/* Host-only consistency check of the kernel's step table for these sizes (dependencies between MMA steps and epilogues, buffer
 * ranges): 0 = consistent, -1 = unsupported sizes, > 0 = the violated rule (csrc/noise_mlp_tc.cu).  Test aid. */
int pcst_noise_predictor_plan_selfcheck(int feature_dim, int time_dim, int nblocks);
size_t pcst_noise_predictor_workspace_bytes(int B, int feature_dim, int nblocks);
int pcst_noise_predictor_pack_f32(const pcst_noise_mlp_t* mlp, void* packed, size_t packed_bytes, pcst_stream_t stream);
int pcst_noise_predictor_f32(const float* points, const int64_t* timestep, const float* style, int B, int N, int feature_dim,
                             int time_dim, int nblocks, const void* packed, float* out, void* ws, size_t ws_bytes,
                             pcst_stream_t stream);

/* ---- voxel-grid downsample: HierarchicalProcessor._voxel_grid_downsample_torch, models/diffusion_model.py:69-122 --
 * (the step in front of the encoder and the denoiser on every 120k-point scan.)
 * pcst_minmax_f32: out [B,6] = per-cloud (min x, min y, min z, max x, max y, max z), :78-79.
 * pcst_voxel_representatives_f32: the deterministic part, :86-93.  xyz_min [B,3] and voxel_size [B] come from the
 *   caller, which forms voxel_size with the reference's own scalar fp32 expression (:80-84).  Per cloud: voxel index
 *   floor((p - min) / voxel_size).int(), int32 hash (ix*73856093)^(iy*19349663)^(iz*83492791), torch.unique order
 *   (ascending signed hash), and per unique hash the truncated float32(sum of member indices) / float32(count).
 *   rep [B,N] int64: the first count[b] entries of row b are valid.  The random thinning / top-up (:95-112) draws
 *   from torch's generator and stays in the Python wrapper. */
int pcst_minmax_f32(const float* xyz, int B, int N, float* out, pcst_stream_t stream);
size_t pcst_voxel_representatives_workspace_bytes(int B, int N);
int pcst_voxel_representatives_f32(const float* xyz, int B, int N, const float* xyz_min, const float* voxel_size,
                                   int64_t* rep, int* count, void* ws, size_t ws_bytes, pcst_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCST_H_ */
