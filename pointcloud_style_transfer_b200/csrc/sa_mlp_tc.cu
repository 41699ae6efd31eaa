// sa_mlp_tc.cu -- SetAbstraction grouping + shared MLP + max-pool on tcgen05 tensor cores (sm_100a).
//
// Replaces index_points x2 + subtract + cat + apply_mlp (models/pointnet2_encoder.py:94-112, eval
// mode) with ONE kernel per set-abstraction stage: gather -> 3 x (GEMM, scale/shift, ReLU) -> max.
// A CTA owns a tile of 128 rows (rows = (b, s, k) flattened = 4 groups of K=32, 2 of K=64, ...) and
// runs a short table of GEMM steps over it.  Warp roles (192 threads):
//   warps 0-3  gather the layer-0 rows from HBM (features with 128-bit loads, then xyz - centroid) into
//              shared memory as the bf16 K-major no-swizzle UMMA operand [Kp/8][128 rows][8], later run the epilogues: read
//              the fp32 accumulator from TMEM (tcgen05.ld 32x32b, warp w owns lanes 32w..32w+31), apply the
//              folded conv-bias / BatchNorm scale+shift and ReLU, and either write the next step's bf16
//              operand back to shared memory or reduce the max over each group's rows;
//   warp 4     allocates TMEM; its lane 0 streams the pre-packed bf16 weights [Kp/8][N][8] through a
//              ring of shared-memory stages in 32-row K chunks with 1-D TMA bulk copies (UBLKCP),
//              running ahead of the math across steps (weights do not depend on data);
//   warp 5     lane 0 issues the tcgen05.mma (M=128, N<=256, K=16, bf16 x bf16 -> fp32 in TMEM), frees
//              ring stages and signals accumulator completion with tcgen05.commit -> mbarrier.
// Layers wider than 256 (the group_all stage: 259 -> 256 -> 512 -> F) are cut into N halves whose
// outputs feed the next layer as K halves accumulating into the same TMEM columns, so the widest
// operand in shared memory stays 128 x 272 bf16 and the accumulators fit the 512 TMEM columns.
// Stages with few row tiles (the group_all stage is ONE tile per scan) are additionally split along N over a
// thread-block cluster of C = 2/4/8 CTAs: every CTA holds the same 128 rows and computes an N/C slice of each
// step, so the weight stream, the MMAs and the epilogues all shrink by C.  After an epilogue a CTA pushes its
// bf16 activation slice (a contiguous block of the [K/8][128][8] operand) into every peer's shared memory with
// cp.async.bulk.shared::cluster (completion counted on the peer's mbarrier), and accumulator completion is
// multicast to all CTAs of the cluster with tcgen05.commit.multicast::cluster, which is also what makes it safe
// to overwrite a peer's operand buffer.
// Activations never touch HBM; weights are packed ONCE per parameter version (pcst_sa_mlp_pack_f32)
// and cached by the caller.  Bound: tensor pipe in the limit of many rows; at the reference's shapes
// (16 384 / 8 192 / 128 rows per scan) a stage is latency-bound.
// Precision: bf16 operands, fp32 accumulate/epilogue -> features within rtol 2e-2 / atol 2e-2 of
// the reference's fp32 result.
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace pcst {

constexpr int kTcMaxN = 256;        // one tcgen05.mma covers a whole step's width
constexpr int kTcStageBytes = 16 * 1024;  // ring stage: a K chunk of a step's weights, as many rows as fit
constexpr int kTcMaxStages = 4;
constexpr int kTcMaxSteps = 9;
constexpr int kTcMaxBlocks = 7;

struct TcStep {
    uint32_t a_off;     // shared-memory offset of the A operand
    uint32_t out_off;   // epi 1: shared-memory offset of the activation buffer written
    uint32_t w_off;     // offset of this step's weight block in the packed blob (sub-block of cluster rank 0)
    uint32_t w_stride;  // bytes between the sub-blocks of consecutive cluster ranks
    uint16_t kp;        // reduction length (multiple of 16)
    uint16_t ck;        // K rows per weight chunk (multiple of 16): ck * n * 2 bytes <= one ring stage
    uint16_t n;         // output channels of the step PER CTA (multiple of 16, <= 256)
    uint16_t tmem_col;  // first accumulator column
    uint16_t out_c0;    // epi 2: first output channel
    uint16_t ss_idx;    // first channel in the scale/shift tables
    uint8_t acc;        // accumulate onto the TMEM contents
    uint8_t epi;        // 0 = none (partial sum), 1 = scale/shift/ReLU -> bf16 operand, 2 = scale/shift/ReLU -> max-pool
};

struct TcArgs {
    const float* xyz;
    const float* feats;
    const __nv_bfloat16* feats_bf16;  // optional bf16 copy of feats (same rounding as the gather's own conversion): half the
                                      // bytes and half the loads of the gather; made per call at throughput shapes
    const float* new_xyz;
    const int64_t* idx;
    int N, S, K, D, rows;
    const unsigned char* blob;  // packed weights, then fp32 scale[total_ch], shift[total_ch]
    uint32_t ss_blob_off;
    int total_ch, kp0, cout;    // cout = channels of the pooled output row
    int nsteps;
    TcStep st[kTcMaxSteps];
    uint32_t off_ring, stage_bytes, nstages, off_ss, off_pool, tmem_cols;
    float* out;                 // [B*S, cout] point-major
    int pool_atomic;            // 0: every group lies inside one tile (plain stores); 1: atomicMax merge
    int early_gather;           // plain three-step plan whose layer-1 output overwrites layer 0's: the next tile is gathered early
    int transpose_pool;         // the pooled layer is computed TRANSPOSED (channels on the TMEM lanes, the tile's rows on the
                                // columns): the max over a group's rows is then a per-thread max over columns
    uint32_t cluster;           // CTAs per row tile (N split); 1 = no cluster
    unsigned long long* probe;  // profiling aid (pcst_sa_mlp_set_probe): [probe_tiles][16] SM-clock stamps per tile, or null
    int probe_tiles;
    int relaxed;                // throughput launch: the producer / MMA threads back off between barrier polls
    int ntiles;                 // row tiles; a CTA (cluster) walks tiles first, first + stride, ... (persistent when > grid)
};

// ---- weight pre-pack: fp32 W[n0 + n, k0 + k] (row-major [Cout, Cin]) -> bf16 [kp/8][nlen][8], zero padded in K ----
// feat_first >= 0 (layer 0): the kernel's operand holds the D = feat_first feature channels FIRST and the three
// relative coordinates after them (16-byte aligned feature groups for the gather), i.e. operand row k is the
// reference's input channel 3 + k for k < D and k - D for D <= k < D + 3 (models/pointnet2_encoder.py:99 order).
__global__ void tc_pack_block_kernel(const float* __restrict__ w, int cin, int k0, int klen, int kp, int n0, int nlen,
                                     int feat_first, __nv_bfloat16* __restrict__ out) {
    const int total = kp * nlen;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kc = e / (nlen * 8);
        const int rem = e % (nlen * 8);
        const int n = rem / 8, ke = rem % 8;
        const int k = kc * 8 + ke;
        int src = k0 + k;
        if (feat_first >= 0) src = k < feat_first ? 3 + k : k - feat_first;
        out[e] = __float2bfloat16_rn(k < klen ? w[(size_t)(n0 + n) * cin + src] : 0.f);
    }
}
__global__ void tc_pack_ss_kernel(const float* __restrict__ scale, const float* __restrict__ shift, int n, int at,
                                  int total_ch, float* __restrict__ out) {
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        out[at + c] = scale[c];
        out[total_ch + at + c] = shift[c];
    }
}

// fp32 feature rows -> bf16 (round to nearest even, exactly what the gather does per element), 8 channels per thread
__global__ void tc_feats_to_bf16_kernel(const float* __restrict__ in, uint4* __restrict__ out, size_t n8) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(in) + 2 * i), b = __ldg(reinterpret_cast<const float4*>(in) + 2 * i + 1);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
        out[i] = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1), *reinterpret_cast<uint32_t*>(&h2),
                            *reinterpret_cast<uint32_t*>(&h3));
    }
}

// Stream-ordered boundary after a pack.  sa_mlp_tc_kernel is launched with programmatic stream serialization and reads
// the packed blob (weights by TMA, the scale/shift tables) BEFORE its griddepcontrol.wait, which is only safe when the
// blob's writers are not its immediate predecessor in the stream: this empty, ordinarily-launched kernel starts after the
// pack kernels have completed (and flushed), and it is the kernel a following PDL launch may overlap with.
__global__ void tc_pack_fence_kernel() {}

// Two instantiations.  <168, true>: CTAs walk several row tiles when there are more tiles than the machine holds (two
// CTAs per SM; barriers, TMEM and the scale/shift tables set up once, the weight ring streaming across tile boundaries).
// <80, false>: one tile per CTA within the register budget that lets FOUR CTAs share an SM.  A CTA's six warps land
// 2/2/1/1 on the four SM partitions (16 Ki registers each), so four CTAs put six warps on a partition: 6 x 32 x 80
// registers fit, 96 allow three CTAs, 112 two.  The narrow stages (SA1: 3 -> 64 -> 64 -> 128, 52 KiB of shared memory
// and 128 TMEM columns per CTA) are bound by the latency of a tile's serial chain, so every additional resident tile
// is worth more than the walk.  Measured at 32 x 16 384 points: 132 us walking with two CTAs per SM, 112 us with three
// (a 96-register build of this code), 93 us with four (the 78-register kernel earlier in the round; this 80-register
// build has not been on a GPU yet -- it differs from the measured 96-register one only in the register cap).
template <int kMaxReg, bool kWalk>
__global__ void __maxnreg__(kMaxReg)
sa_mlp_tc_kernel(const __grid_constant__ TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kTcMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kTcMaxStages];
    __shared__ __align__(8) uint64_t mma_bar;  // accumulator of an epilogue-bearing step is complete
    __shared__ __align__(8) uint64_t a_bar;    // an epilogue (or the gather) has finished: operand written, TMEM read
    __shared__ __align__(8) uint64_t x_bar[2]; // the peers' activation slices of an epilogue have landed (by event parity)
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, lane = tid & 31;
    // warp-uniform FOR THE COMPILER: the role branches below become uniform branches and the single-lane roles' loop
    // counters, table reads and descriptors live in uniform registers (see the MMA issuer)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t C = a.cluster;
    const uint32_t rank = C > 1 ? cluster_ctarank() : 0;
    const int tile_first = (int)(blockIdx.x / C), tile_stride = (int)(gridDim.x / C);
    // Programmatic dependent launch: let the next kernel of the stream start its own prologue now; this kernel's
    // prologue (barriers, TMEM, cluster rendezvous, the first weight chunks, the scale/shift tables -- nothing the
    // previous kernel produces) runs before griddepcontrol.wait, the gather of the previous kernel's output after it.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (tid == 0) {
        for (int s = 0; s < kTcMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&mma_bar, C);  // every CTA of the cluster commits to every CTA's barrier
        mbar_init(&a_bar, kTcEpiThreads / 32);  // one arrival per epilogue warp (after __syncwarp)
        mbar_init(&x_bar[0], 1);
        mbar_init(&x_bar[1], 1);
        fence_mbar_init();
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(a.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (C > 1) cluster_sync_all();  // every peer's barriers exist before anything is sent to them
    const uint32_t tmem_base = tmem_base_sh;

    // shared-window addresses formed once: a generic pointer to a __shared__ object is re-derived with S2UR SR_CgaCtaId at
    // every use in a cluster launch
    const uint32_t smem_base = tc_opaque_u32(smem_u32(smem));
    const uint32_t full0 = tc_opaque_u32(smem_u32(&full_bar[0])), empty0 = tc_opaque_u32(smem_u32(&empty_bar[0]));
    const uint32_t mma_bar_a = tc_opaque_u32(smem_u32(&mma_bar)), a_bar_a = tc_opaque_u32(smem_u32(&a_bar));
    const uint32_t x_bar_a = tc_opaque_u32(smem_u32(&x_bar[0]));
    const bool relaxed = kWalk && a.relaxed;
    auto wait_relaxed = [&](uint32_t bar, uint32_t parity) {
        if (relaxed) {
            uint32_t ok;
            for (;;) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(bar), "r"(parity)
                    : "memory");
                if (ok) break;
                __nanosleep(64);
            }
        } else {
            mbar_wait_addr(bar, parity);
        }
    };

    if (warp == 4) {
        // ================= weight producer =================
        if (elect_one_sync()) {
            // the ring never drains between tiles: the next tile's first chunks load under this tile's epilogues
            const uint32_t nstages = (uint32_t)a.nstages;
            uint32_t stage = 0, round = 0;   // ring position; how often the ring has wrapped
            for (int tile = tile_first; tile < a.ntiles; tile += kWalk ? tile_stride : a.ntiles)
                for (int s = 0; s < a.nsteps; ++s) {
                    const TcStep& st = a.st[s];
                    const unsigned char* src = a.blob + st.w_off + (size_t)rank * st.w_stride;
                    const uint32_t kp = (uint32_t)st.kp, ckf = (uint32_t)st.ck, row_bytes = (uint32_t)st.n * 2u;
                    for (uint32_t k0 = 0; k0 < kp; k0 += ckf) {
                        if (round > 0) wait_relaxed(empty0 + stage * 8u, (round - 1u) & 1u);
                        const uint32_t bytes = min(ckf, kp - k0) * row_bytes;
                        mbar_expect_tx_addr(full0 + stage * 8u, bytes);
                        tma_load_1d_addr(smem_base + a.off_ring + stage * a.stage_bytes, src, bytes, full0 + stage * 8u);
                        src += bytes;
                        if (++stage == nstages) { stage = 0; ++round; }
                    }
                }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        // The whole warp walks the loop converged and one elected lane issues.  Issued from a lone lane of a divergent
        // branch, every tcgen05.mma operand is a per-thread value for the compiler: a register -> uniform-register move
        // each and an ELECT / BRA.U.ANY waterfall loop around every UTCHMMA, ~200 cycles per MMA against a tensor-pipe
        // floor of 32-128 (measured on the denoiser kernel, csrc/noise_mlp_tc.cu).  Converged, the loop runs on the uniform
        // datapath and a stage's MMAs issue back to back.
        //
        // Events on a_bar, per tile: the gather, then one per epilogue EXCEPT the tile's last (pooled) one -- the
        // next tile's gather event follows it in the same threads and stands for both (two arrivals nobody waits
        // between would let the barrier run two phases ahead of this thread's parity wait).
        const bool leader = elect_one_sync();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);     // SBO = 128 B; descriptor version 1 (bit 46)
        const uint32_t a_step = (2u * kTcM * 16u) >> 4;        // one K = 16 slice of A: two K groups of 128 rows
        const uint32_t ring_lo = (smem_base + a.off_ring) >> 4, stage_lo = a.stage_bytes >> 4;
        const uint32_t nstages = (uint32_t)a.nstages;
        uint32_t stage = 0, round = 0, seen = 0, need = 0;
        uint32_t xseen = 0;                   // operand-writing epilogues whose REMOTE slices have been awaited
        for (int tile = tile_first; tile < a.ntiles; tile += kWalk ? tile_stride : a.ntiles) {
            ++need;                           // this tile's gather
            int xstep = 0;                    // step scanned up to while counting those epilogues
            for (int s = 0; s < a.nsteps; ++s) {
                const TcStep& st = a.st[s];
                while (seen < need) {
                    wait_relaxed(a_bar_a, seen & 1u);
                    ++seen;
                }
                if (C > 1) {
                    // every operand-writing epilogue before this step: (C - 1) peers each push 128 x n x 2 bytes
                    for (; xstep < s; ++xstep) {
                        if (a.st[xstep].epi != 1) continue;
                        const uint32_t xb = x_bar_a + (xseen & 1u) * 8u;
                        if (leader) mbar_expect_tx_addr(xb, (C - 1u) * (uint32_t)a.st[xstep].n * (kTcM * 2u));
                        __syncwarp();
                        mbar_wait_addr(xb, (xseen >> 1) & 1u);
                        ++xseen;
                    }
                }
                tc_fence_after();
                const uint32_t n = (uint32_t)st.n, kp = (uint32_t)st.kp, ckf = (uint32_t)st.ck;
                const uint32_t idesc = umma_idesc_bf16(kTcM, (int)n);
                const uint32_t d_addr = tmem_base + st.tmem_col;
                const uint32_t w_step = n * 2u;                // one K = 16 slice of the weights, 16-byte units
                const uint32_t w_lbo = n << 16;                // LBO = n * 16 bytes
                uint32_t a_lo = (((smem_base + st.a_off) >> 4) & 0x3FFFu) | (((kTcM * 16u) >> 4) << 16);
                uint32_t accum = st.acc;
                // pooled layer, transposed: D^T[channel block of 128][the tile's 128 rows] = W[block, :] x A^T -- the
                // weights are the M-side operand (rows cb * 128 ... of every K group), the activations the N-side one
                const bool transposed = a.transpose_pool && st.epi != 1;
                const uint32_t idesc_t = umma_idesc_bf16(kTcM, kTcM);
                for (uint32_t k0 = 0; k0 < kp; k0 += ckf) {
                    mbar_wait_addr(full0 + stage * 8u, round & 1u);
                    tc_fence_after();
                    const uint32_t nk = min(ckf, kp - k0) >> 4;
                    if (leader) {
                        const uint32_t w_lo = ((ring_lo + stage * stage_lo) & 0x3FFFu) | w_lbo;
                        auto mma = [&](uint32_t kk, bool acc) {
                            umma_bf16(d_addr, ((uint64_t)desc_hi << 32) | (a_lo + kk * a_step),
                                      ((uint64_t)desc_hi << 32) | (w_lo + kk * w_step), idesc, acc);
                        };
                        if (transposed) {
                            auto mma_t = [&](uint32_t kk, uint32_t cb, bool acc) {
                                umma_bf16(d_addr + cb * 128u, ((uint64_t)desc_hi << 32) | (w_lo + kk * w_step + cb * 128u),
                                          ((uint64_t)desc_hi << 32) | (a_lo + kk * a_step), idesc_t, acc);
                            };
                            if (n == 256 && nk == 2) {
                                mma_t(0, 0, accum != 0); mma_t(0, 1, accum != 0); mma_t(1, 0, true); mma_t(1, 1, true);
                            } else if (n == 128 && nk == 4) {
                                mma_t(0, 0, accum != 0); mma_t(1, 0, true); mma_t(2, 0, true); mma_t(3, 0, true);
                            } else {
                                for (uint32_t kk = 0; kk < nk; ++kk)
                                    for (uint32_t cb = 0; cb < (n >> 7); ++cb) mma_t(kk, cb, (accum | kk) != 0);
                            }
                        } else if (nk == 4) {
                            mma(0, accum != 0); mma(1, true); mma(2, true); mma(3, true);
                        } else if (nk == 2) {
                            mma(0, accum != 0); mma(1, true);
                        } else {
                            for (uint32_t kk = 0; kk < nk; ++kk) mma(kk, (accum | kk) != 0);
                        }
                        umma_commit_addr(empty0 + stage * 8u);  // the stage is free once these MMAs have read it
                    }
                    __syncwarp();
                    a_lo += nk * a_step;
                    accum = 1;
                    if (++stage == nstages) { stage = 0; ++round; }
                }
                if (st.epi) {
                    if (leader) {
                        if (C > 1) umma_commit_multicast_addr(mma_bar_a, (uint16_t)((1u << C) - 1u));
                        else umma_commit_addr(mma_bar_a);
                    }
                    __syncwarp();
                    if (s != a.nsteps - 1) ++need;
                }
            }
        }
    } else {
        // ================= gather + epilogues (128 threads, thread = row = TMEM lane) =================
        float* ss = reinterpret_cast<float*>(smem + a.off_ss);  // scale[total_ch], shift[total_ch]
        {
            const float* src = reinterpret_cast<const float*>(a.blob + a.ss_blob_off);
            for (int c = tid; c < 2 * a.total_ch; c += kTcEpiThreads) ss[c] = src[c];
        }
        asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous kernel's output (idx, feats, xyz) is complete
        const uint32_t taddr_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int m = tid;  // this thread's row = its TMEM lane
        uint32_t mma_phase = 0, xev = 0;
        // The gather is a chain of dependent loads (index -> feature row); when a CTA walks several tiles the NEXT tile's
        // index is requested before this tile's epilogues, so that the chain of the next gather starts one latency shorter.
        long long idx_next = 0;
        if (kWalk && a.idx && tile_first * kTcM + tid < a.rows) idx_next = a.idx[tile_first * kTcM + tid];
        // gather of row tile gt into the layer-0 operand buffer (tiles are gathered in walk order: idx_next is gt's index)
        auto gather_tile = [&](int gt) {
            {
                // Layer-0 operand, one thread per row: operand row k = feature channel k (k < D), then the three
                // relative coordinates, then zero padding up to kp0.  Features are read with 16 independent 128-bit
                // loads in flight per thread and stored as 16-byte (8 x bf16) core-matrix rows: thread m writes
                // [(k / 8) * 128 + m], so a warp's stores are contiguous (no bank conflicts).
                unsigned char* A0b = smem + a.st[0].a_off;
                __nv_bfloat16* A0 = reinterpret_cast<__nv_bfloat16*>(A0b);
                const int D = a.D, kp0 = a.kp0;
                const int row = gt * kTcM + tid;
                const bool valid = row < a.rows;
                int j = 0, bs = 0;
                if (valid) {
                    bs = row / a.K;
                    if (a.idx) {
                        const int64_t jj = kWalk ? idx_next : a.idx[row];
                        j = jj < 0 ? 0 : (jj >= a.N ? a.N - 1 : (int)jj);
                    } else {
                        j = row % a.K;  // group_all: row k of cloud b is point k
                    }
                }
                const int b = bs / a.S;
                float rel[3] = {0.f, 0.f, 0.f};
                if (valid) {
                    const float* p = a.xyz + ((size_t)b * a.N + j) * 3;
                    rel[0] = p[0]; rel[1] = p[1]; rel[2] = p[2];
                    if (a.new_xyz) {
                        const float* c = a.new_xyz + (size_t)bs * 3;
                        rel[0] = __fsub_rn(rel[0], c[0]); rel[1] = __fsub_rn(rel[1], c[1]); rel[2] = __fsub_rn(rel[2], c[2]);
                    }
                }
                const float* frow = D > 0 ? a.feats + ((size_t)b * a.N + j) * D : nullptr;
                const bool vec = D > 0 && (D % 8) == 0 && ((reinterpret_cast<uintptr_t>(a.feats) & 15) == 0);
                int kdone = 0;
                if (a.feats_bf16) {
                    // bf16 rows (the host guarantees D % 8 == 0): 16 bytes = one K group, stored as loaded
                    const uint4* frow16 = reinterpret_cast<const uint4*>(a.feats_bf16 + ((size_t)b * a.N + j) * D);
                    const int G = D / 8;
                    for (int g0 = 0; g0 < G; g0 += 16) {
                        uint4 v[16];
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            v[u] = make_uint4(0u, 0u, 0u, 0u);
                            if (valid && g0 + u < G) v[u] = __ldg(frow16 + g0 + u);
                        }
#pragma unroll
                        for (int u = 0; u < 16; ++u)
                            if (g0 + u < G) *reinterpret_cast<uint4*>(A0b + ((size_t)(g0 + u) * kTcM + tid) * 16) = v[u];
                    }
                    kdone = D;
                } else if (vec) {
                    const int G = D / 8;
                    for (int g0 = 0; g0 < G; g0 += 8) {
                        float4 v[16];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            v[2 * u] = v[2 * u + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (valid && g0 + u < G) {
                                v[2 * u] = __ldg(reinterpret_cast<const float4*>(frow + (g0 + u) * 8));
                                v[2 * u + 1] = __ldg(reinterpret_cast<const float4*>(frow + (g0 + u) * 8 + 4));
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            if (g0 + u < G) {
                                __nv_bfloat162 h0 = __floats2bfloat162_rn(v[2 * u].x, v[2 * u].y);
                                __nv_bfloat162 h1 = __floats2bfloat162_rn(v[2 * u].z, v[2 * u].w);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * u + 1].x, v[2 * u + 1].y);
                                __nv_bfloat162 h3 = __floats2bfloat162_rn(v[2 * u + 1].z, v[2 * u + 1].w);
                                *reinterpret_cast<uint4*>(A0b + ((size_t)(g0 + u) * kTcM + tid) * 16) =
                                    make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                               *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
                            }
                        }
                    }
                    kdone = D;
                }
                if (kdone == D && (D % 8) == 0) {
                    // the tail starts on a K-group boundary: (rel x, rel y, rel z, 0 ...) and zero groups, one 16-byte store each
                    // (element-wise 2-byte stores of a warp land 8 to a bank)
                    __nv_bfloat162 xy = __floats2bfloat162_rn(rel[0], rel[1]), z0 = __floats2bfloat162_rn(rel[2], 0.f);
                    *reinterpret_cast<uint4*>(A0b + ((size_t)(D >> 3) * kTcM + tid) * 16) =
                        make_uint4(*reinterpret_cast<uint32_t*>(&xy), *reinterpret_cast<uint32_t*>(&z0), 0u, 0u);
                    for (int g = (D >> 3) + 1; g < (kp0 >> 3); ++g)
                        *reinterpret_cast<uint4*>(A0b + ((size_t)g * kTcM + tid) * 16) = make_uint4(0u, 0u, 0u, 0u);
                    kdone = kp0;
                }
                for (int k = kdone; k < kp0; ++k) {  // feature tail (unaligned D), relative coordinates, zero padding
                    float v = 0.f;
                    if (k < D) v = valid ? __ldg(frow + k) : 0.f;
                    else if (k - D < 3) v = (k - D) == 0 ? rel[0] : ((k - D) == 1 ? rel[1] : rel[2]);
                    A0[((size_t)(k >> 3) * kTcM + tid) * 8 + (k & 7)] = __float2bfloat16_rn(v);
                }
            }
            if (kWalk && a.idx) {
                const long long rn = (long long)(gt + tile_stride) * kTcM + tid;
                if (gt + tile_stride < a.ntiles && rn < a.rows) idx_next = a.idx[rn];
            }
        };
        bool gathered = false;  // the tile's operand is already in place (gathered under the previous tile's last MMAs)
        for (int tile = tile_first; tile < a.ntiles; tile += kWalk ? tile_stride : a.ntiles) {
            const int row0 = tile * kTcM;
            // stamps: 0 tile start, then per epilogue-bearing step (accumulator ready, epilogue done); 15 = SM id
            unsigned long long* stamp = (kWalk && a.probe && tid == 0 && rank == 0 && tile < a.probe_tiles) ? a.probe + (size_t)tile * 16 : nullptr;
            int nstamp = 0;
            if (stamp) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                stamp[15] = smid;
                stamp[nstamp++] = clock64();
            }
            if (!gathered) gather_tile(tile);
            gathered = false;
            fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
            if (tile == tile_first) epi_bar_sync();  // the scale/shift tables are complete for every epilogue thread
            tc_fence_before();    // (later tiles) the previous tile's accumulator reads precede the MMAs this releases
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(a_bar_a);  // the layer-0 operand is in place

            for (int s = 0; s < a.nsteps; ++s) {
                const TcStep& st = a.st[s];
                if (!st.epi) continue;
                if (kWalk && a.early_gather && s == a.nsteps - 1 && tile + tile_stride < a.ntiles) {
                    // The layer-0 operand buffer has been free since the first step's MMAs (layer 1 writes its output over
                    // layer 0's, not into it): gather the NEXT tile now, under this tile's last MMAs, instead of after its
                    // pooled epilogue.  The arrival that publishes it stays at the top of the next iteration.
                    gather_tile(tile + tile_stride);
                    gathered = true;
                }
                mbar_wait_addr(mma_bar_a, mma_phase & 1u);  // all C CTAs have finished the MMAs up to this step
                ++mma_phase;
                tc_fence_after();
                if (stamp && nstamp < 14) stamp[nstamp++] = clock64();
                const int slice0 = (int)rank * st.n;  // first channel of this CTA's N slice inside the step
                const float* sc = ss + st.ss_idx + slice0;
                const float* sh = ss + a.total_ch + st.ss_idx + slice0;
                const uint32_t taddr = taddr_lane + st.tmem_col;
                if (st.epi == 1) {
                    unsigned char* outp = smem + st.out_off + (size_t)(slice0 >> 3) * kTcM * 16;  // this slice's K groups
                    // two 16-column loads in flight per wait (ptxas sinks a prefetch issued after the wait below the
                    // arithmetic of the previous block, so the software-pipelined walk is kept for the pooled steps only)
                    for (int c0 = 0; c0 < st.n; c0 += 32) {
                        uint32_t r[2][16];
                        tmem_ld16_issue(taddr + (uint32_t)c0, r[0]);
                        const bool two = c0 + 16 < st.n;
                        if (two) tmem_ld16_issue(taddr + (uint32_t)c0 + 16u, r[1]);
                        tmem_ld_wait(r[0]);
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h == 1 && !two) break;
                            const int cb = c0 + 16 * h;
                            uint32_t packed[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                // two channels per step: one packed FFMA2 for scale/shift, one cvt.rn.relu.bf16x2 for ReLU +
                                // rounding + packing (upper half = second channel)
                                const float2 s2 = *reinterpret_cast<const float2*>(&sc[cb + 2 * i]);
                                const float2 t2 = *reinterpret_cast<const float2*>(&sh[cb + 2 * i]);
                                const float2 y = __ffma2_rn(make_float2(__uint_as_float(r[h][2 * i]), __uint_as_float(r[h][2 * i + 1])), s2, t2);
                                asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(packed[i]) : "f"(y.y), "f"(y.x));
                            }
                            // channels cb..cb+7 and cb+8..cb+15 are two K chunks of the next operand
                            uint4* d0 = reinterpret_cast<uint4*>(outp + ((size_t)(cb >> 3) * kTcM + m) * 16);
                            uint4* d1 = reinterpret_cast<uint4*>(outp + ((size_t)((cb >> 3) + 1) * kTcM + m) * 16);
                            *d0 = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                            *d1 = make_uint4(packed[4], packed[5], packed[6], packed[7]);
                        }
                    }
                    fence_proxy_async();
                    if (C > 1) {
                        // push the slice (contiguous: n / 8 K-groups x 128 rows x 16 B) into every peer's operand buffer
                        epi_bar_sync();
                        if (tid == 0) {
                            const uint32_t src = smem_u32(outp), bytes = (uint32_t)st.n * (kTcM * 2u);
                            const uint32_t bar = x_bar_a + (xev & 1u) * 8u;
                            for (uint32_t q = 0; q < C; ++q)
                                if (q != rank) bulk_copy_to_peer(mapa_shared(src, q), src, bytes, mapa_shared(bar, q));
                        }
                        ++xev;
                    }
                } else {
                    // ---- max over each group's K rows ----
                    const int row = row0 + m;
                    const bool valid = row < a.rows;
                    if (a.transpose_pool) {
                        // This thread's TMEM lane is a CHANNEL, the accumulator's 128 columns are the tile's rows: scale /
                        // shift are two scalars, the max over a group is a running max over its K columns (no warp
                        // reductions, no staging buffer, no block barrier), and a warp stores 32 consecutive channels of a
                        // pooled row.  relu(max) == max(relu): the ReLU is applied once per pooled value.
                        const int rows_here = a.rows - row0 < kTcM ? a.rows - row0 : kTcM;   // valid rows = valid columns
                        const int g0 = row0 / a.K;                                           // first group of the tile
                        auto pooled_t = [&](auto masked) {
                            for (int cb = 0; cb < st.n / 128; ++cb) {
                                const int ch = cb * 128 + m;
                                const float scl = sc[ch], sft = sh[ch];
                                float mx = __int_as_float(0xff800000);
                                tmem_for_each16(taddr + (uint32_t)cb * 128u, kTcM, [&](const uint32_t (&r)[16], int c0) {
#pragma unroll
                                    for (int i = 0; i < 16; ++i) {
                                        const float y = __fmaf_rn(__uint_as_float(r[i]), scl, sft);
                                        if (!decltype(masked)::value || c0 + i < rows_here) mx = fmaxf(mx, y);
                                    }
                                    if (((c0 + 16) % a.K) == 0) {   // the last 16 rows of a group (K is 32, 64 or 128)
                                        const int g = g0 + (c0 + 16) / a.K - 1;
                                        if ((size_t)g * a.K < (size_t)a.rows)
                                            a.out[(size_t)g * a.cout + st.out_c0 + slice0 + ch] = fmaxf(mx, 0.f);
                                        mx = __int_as_float(0xff800000);
                                    }
                                });
                            }
                        };
                        if (row0 + kTcM <= a.rows) pooled_t(std::false_type{});
                        else pooled_t(std::true_type{});
                    } else if (!a.pool_atomic) {
                        // K in {32, 64, 128}: the warp's 32 rows belong to one group and every group lies in this tile
                        float* pool = reinterpret_cast<float*>(smem + a.off_pool);  // [4 warps][n]
                        // relu(max over rows) == max over rows of relu: the warp reduces the RAW scale/shift results as
                        // SIGNED integers (any non-negative float beats every negative one and non-negative floats order
                        // like their bits; an all-negative column yields some negative value) and the ReLU is applied
                        // once, where the pooled value is read below.  Four columns per step: two LDS.128 for scale and
                        // shift, two FFMA2, four warp reductions, one 16-byte store of the (uniform) results.
                        if constexpr (!kWalk) {
                            // narrow build (80 registers): 16 independent reductions per accumulator block, each
                            // lane keeping one of them
                            for (int c0 = 0; c0 < st.n; c0 += 16) {
                                uint32_t r[16];
                                tmem_ld16_issue(taddr + (uint32_t)c0, r);
                                tmem_ld_wait(r);
                                unsigned int mine = 0;
#pragma unroll
                                for (int i = 0; i < 16; ++i) {
                                    float y = __fmaf_rn(__uint_as_float(r[i]), sc[c0 + i], sh[c0 + i]);
                                    y = (valid && y > 0.f) ? y : 0.f;  // post-ReLU values are >= +0: bits order like unsigned
                                    const unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
                                    if (lane == i) mine = mx;
                                }
                                if (lane < 16) pool[warp * st.n + c0 + lane] = __uint_as_float(mine);
                            }
                        } else {
                            // relu(max over rows) == max over rows of relu: the warp reduces the RAW scale/shift results
                            // as SIGNED integers (any non-negative float beats every negative one and non-negative floats
                            // order like their bits; an all-negative column yields some negative value) and the ReLU is
                            // applied once, where the pooled value is read below.  Per 16 columns: LDS.128 for scale and
                            // shift, 8 FFMA2, 16 independent warp reductions, then four 16-byte stores of the (uniform)
                            // results; the next 16 columns are already loading from TMEM.
                            float* prow = pool + warp * st.n;
                            auto pooled = [&](auto masked) {
                                tmem_for_each16(taddr, st.n, [&](const uint32_t (&r)[16], int c0) {
                                    int mx[16];
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const float4 s4 = *reinterpret_cast<const float4*>(&sc[c0 + 4 * q]);
                                        const float4 t4 = *reinterpret_cast<const float4*>(&sh[c0 + 4 * q]);
                                        const float2 y01 =
                                            __ffma2_rn(make_float2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])),
                                                       make_float2(s4.x, s4.y), make_float2(t4.x, t4.y));
                                        const float2 y23 =
                                            __ffma2_rn(make_float2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])),
                                                       make_float2(s4.z, s4.w), make_float2(t4.z, t4.w));
                                        int v[4] = {__float_as_int(y01.x), __float_as_int(y01.y), __float_as_int(y23.x),
                                                    __float_as_int(y23.y)};
                                        if (decltype(masked)::value) {  // rows past the end (last tile only): below every value
#pragma unroll
                                            for (int i = 0; i < 4; ++i) v[i] = valid ? v[i] : (int)0x80000000;
                                        }
#pragma unroll
                                        for (int i = 0; i < 4; ++i) mx[4 * q + i] = __reduce_max_sync(0xffffffffu, v[i]);
                                    }
                                    if (lane == 0) {
#pragma unroll
                                        for (int q = 0; q < 4; ++q)
                                            *reinterpret_cast<int4*>(prow + c0 + 4 * q) =
                                                make_int4(mx[4 * q], mx[4 * q + 1], mx[4 * q + 2], mx[4 * q + 3]);
                                    }
                                });
                            };
                            if (row0 + kTcM <= a.rows) pooled(std::false_type{});
                            else pooled(std::true_type{});
                        }
                        epi_bar_sync();
                        const int wpg = a.K / 32;       // warps per group
                        const int groups = kTcM / a.K;  // groups per tile
                        const int g0 = row0 / a.K;      // first group (= b * S + s) of the tile
                        for (int e = tid; e < groups * st.n; e += kTcEpiThreads) {
                            const int g = e / st.n, c = e % st.n;
                            if ((size_t)(g0 + g) * a.K >= (size_t)a.rows) continue;
                            float v = pool[(g * wpg) * st.n + c];
                            for (int w = 1; w < wpg; ++w) v = fmaxf(v, pool[(g * wpg + w) * st.n + c]);
                            a.out[(size_t)(g0 + g) * a.cout + st.out_c0 + slice0 + c] = v > 0.f ? v : 0.f;  // the ReLU
                        }
                        epi_bar_sync();  // pool is reused by the next pooled step
                    } else {
                        const bool warp_uniform_group = (a.K % 32) == 0;
                        const int wrow = row0 + warp * 32;
                        const int g = (warp_uniform_group ? wrow : (valid ? row : 0)) / a.K;
                        unsigned int* obase = reinterpret_cast<unsigned int*>(a.out) + (size_t)g * a.cout + st.out_c0 + slice0;
                        tmem_for_each16(taddr, st.n, [&](const uint32_t (&r)[16], int c0) {
                            unsigned int mine = 0;
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                float y = __fmaf_rn(__uint_as_float(r[i]), sc[c0 + i], sh[c0 + i]);
                                y = (valid && y > 0.f) ? y : 0.f;
                                if (warp_uniform_group) {
                                    const unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
                                    if (lane == i) mine = mx;
                                } else if (valid) {
                                    atomicMax(obase + c0 + i, __float_as_uint(y));
                                }
                            }
                            if (warp_uniform_group && lane < 16 && wrow < a.rows) atomicMax(obase + c0 + lane, mine);
                        });
                    }
                }
                if (stamp && nstamp < 14) stamp[nstamp++] = clock64();
                if (s != a.nsteps - 1) {  // the last (pooled) step: the next tile's gather event stands for it
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_addr(a_bar_a);  // operand written / accumulator drained: later steps may proceed
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();  // no CTA leaves while a peer may still push slices or commits into it
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols)
                     : "memory");
    }
}

// ---- host side: the step table, the shared-memory layout and the packed-blob layout ----------------
struct TcBlock {
    int layer, k0, klen, kp, n0, nlen;  // logical block: W[n0 .. n0+nlen, k0 .. k0+klen], K padded to kp
    uint32_t w_off, sub_bytes;          // blob offset of rank 0's sub-block; bytes between ranks' sub-blocks
};
struct TcPlan {
    bool ok, reuse_h;
    int nsteps, nblocks, kp0, total_ch, C;
    int n[3];
    TcStep st[kTcMaxSteps];
    TcBlock blk[kTcMaxBlocks];
    uint32_t off_ring, stage_bytes, nstages, off_ss, off_pool, smem_bytes, tmem_cols, ss_blob_off;
    size_t blob_bytes;
};

// C = CTAs per row tile (N split over a cluster): every step computes n / C channels per CTA.
// dense = many more row tiles than SMs: a small weight ring (2 stages of 8 KiB) so that 2-4 CTAs share an SM and one
// tile's epilogue overlaps another's MMAs; otherwise a deep ring for the latency of a single tile.
static TcPlan tc_plan(int D, const int* cout, int C, bool dense = false, int dense_stage_bytes = 0) {
    if (dense && dense_stage_bytes == 0) {
        // Stage size of the dense (many tiles per SM) plan: every stage costs the issuer a barrier wait, a commit and ~300
        // cycles of loop overhead, so take 16 KiB stages when they leave as many CTAs per SM as 8 KiB ones do.
        const TcPlan small = tc_plan(D, cout, C, true, kTcStageBytes / 2), big = tc_plan(D, cout, C, true, kTcStageBytes);
        auto per_sm = [](const TcPlan& q) {
            const uint32_t n = 228u * 1024u / (q.smem_bytes + 1024u);
            return n < 4u ? n : 4u;
        };
        return (big.ok && (!small.ok || per_sm(big) >= per_sm(small))) ? big : small;
    }
    TcPlan p = {};
    p.ok = false;
    p.C = C;
    if (C != 1 && C != 2 && C != 4 && C != 8) return p;
    const int n0 = cout[0], n1 = cout[1], n2 = cout[2];
    for (int l = 0; l < 3; ++l) {
        p.n[l] = cout[l];
        if (cout[l] <= 0 || (cout[l] % 32) != 0 || cout[l] > 2 * kTcMaxN) return p;
    }
    if (n0 > kTcMaxN) return p;
    const int h1 = n1 > kTcMaxN ? 2 : 1, h2 = n2 > kTcMaxN ? 2 : 1;  // N halves of layers 1 and 2
    const int n1h = n1 / h1, n2h = n2 / h2;
    if ((n0 % (16 * C)) || (n1h % (16 * C)) || (n2h % (16 * C))) return p;  // per-CTA slices: multiples of 16 columns
    const int n0c = n0 / C, n1c = n1h / C, n2c = n2h / C;
    p.kp0 = (int)align_up((size_t)(3 + D), 16);
    p.total_ch = n0 + n1 + n2;

    // weight blocks in the blob: C consecutive sub-blocks [kp/8][nlen/C][8] per logical block
    uint32_t woff = 0;
    auto add_block = [&](int layer, int k0, int klen, int kp, int nb0, int nlen) -> int {
        TcBlock& b = p.blk[p.nblocks];
        b.layer = layer; b.k0 = k0; b.klen = klen; b.kp = kp; b.n0 = nb0; b.nlen = nlen; b.w_off = woff;
        b.sub_bytes = (uint32_t)align_up((size_t)kp * (nlen / C) * 2, 128);
        woff += b.sub_bytes * C;
        return p.nblocks++;
    };
    const int b0 = add_block(0, 0, 3 + D, p.kp0, 0, n0);
    int b1[2] = {0, 0}, b2[2][2] = {{0, 0}, {0, 0}};
    for (int kh = 0; kh < h1; ++kh) b1[kh] = add_block(1, 0, n0, n0, kh * n1h, n1h);
    for (int fh = 0; fh < h2; ++fh)
        for (int kh = 0; kh < h1; ++kh) b2[kh][fh] = add_block(2, kh * n1h, n1h, n1h, fh * n2h, n2h);
    p.ss_blob_off = woff;
    p.blob_bytes = (size_t)woff + (size_t)2 * p.total_ch * sizeof(float);

    // shared memory (operand buffers hold the FULL width: the peers' slices are pushed into them)
    const int max_nc = n0c > n1c ? (n0c > n2c ? n0c : n2c) : (n1c > n2c ? n1c : n2c);
    // K rows per chunk of a step: as many as fit a ring stage (narrow per-CTA slices get long chunks, so that the
    // number of dependent TMA round trips stays small)
    const int stage_target = dense ? dense_stage_bytes : kTcStageBytes;
    auto chunk_rows = [stage_target](int kp, int nc) {
        int ck = stage_target / (nc * 2) / 16 * 16;
        if (ck < 16) ck = 16;
        return ck < kp ? ck : kp;
    };
    p.stage_bytes = 0;
    {
        const int kps[3] = {p.kp0, n0, n1h}, ncs[3] = {n0c, n1c, n2c};
        for (int l = 0; l < 3; ++l) {
            const uint32_t sb = (uint32_t)chunk_rows(kps[l], ncs[l]) * ncs[l] * 2;
            if (sb > p.stage_bytes) p.stage_bytes = sb;
        }
        p.stage_bytes = (uint32_t)align_up(p.stage_bytes, 128);
    }
    // Plain plans of one CTA per tile (no N / K halves, no cluster): layer 1's output OVERWRITES layer 0's (its readers, the
    // layer-1 MMAs, are complete when that epilogue starts), so the layer-0 input buffer is free from the first step's
    // MMAs on and the next tile can be gathered into it early.  Otherwise layer 1's (half) output goes to the input buffer.
    // Only where there are feature rows to gather (D > 0): the coordinates-only first stage runs 19 % SLOWER with the shared
    // buffer (measured: 64.7 -> 76.8 us at 32 x 16 384 points) and has next to nothing to gather early.
    p.reuse_h = (h1 * h2 == 1) && C == 1 && D > 0 && tuning("sa_mlp.reuse_h", 0) != 2;
    const uint32_t size_x = (uint32_t)kTcM * (p.reuse_h ? p.kp0 : (p.kp0 > n1h ? p.kp0 : n1h)) * 2;  // layer-0 input (later a layer-1 half)
    const uint32_t size_h = (uint32_t)kTcM * (p.reuse_h ? (n0 > n1 ? n0 : n1) : n0) * 2;               // layer-0 output
    const uint32_t size_ss = (uint32_t)align_up((size_t)2 * p.total_ch * sizeof(float), 128);
    const uint32_t size_pool = (uint32_t)align_up((size_t)4 * n2c * sizeof(float), 128);
    const uint32_t fixed = size_x + size_h + size_ss + size_pool;
    const uint32_t limit = 225 * 1024;
    auto nchunks = [&](int kp, int nc) { const int ck = chunk_rows(kp, nc); return (kp + ck - 1) / ck; };
    int total_chunks = nchunks(p.kp0, n0c) + h2 * h1 * (nchunks(n0, n1c) + nchunks(n1h, n2c));
    p.nstages = kTcMaxStages;
    while (p.nstages > 2 && (fixed + p.nstages * p.stage_bytes > limit || (int)p.nstages > total_chunks)) --p.nstages;
    if (dense) {
        // the deepest ring that does not cost a co-resident CTA (4 is the register-file bound of the narrow build)
        auto ctas = [&](uint32_t ns) {
            const uint32_t per_sm = 228u * 1024u / (fixed + ns * p.stage_bytes + 1024u);
            return per_sm < 4u ? per_sm : 4u;
        };
        p.nstages = kTcMaxStages;
        while (p.nstages > 2 && ctas(p.nstages) < ctas(2)) --p.nstages;
    }
    if (fixed + p.nstages * p.stage_bytes > limit) return p;
    p.off_ring = 0;
    const uint32_t off_x = p.nstages * p.stage_bytes;
    const uint32_t off_h = off_x + size_x;
    p.off_ss = off_h + size_h;
    p.off_pool = p.off_ss + size_ss;
    p.smem_bytes = p.off_pool + size_pool;

    // steps (n = per-CTA slice width; ss_idx / out_c0 = first channel of the whole step, the kernel adds rank * n)
    const bool split = h1 * h2 > 1;
    const int l1col = split ? (int)align_up((size_t)(n0c > n2c ? n0c : n2c), 32) : 0;  // layer-1 accumulator next to layer 2's
    auto add_step = [&](uint32_t a_off, uint32_t out_off, int blk, int kp, int n, int col, int out_c0, int ss_idx,
                        int acc, int epi) {
        TcStep& s = p.st[p.nsteps++];
        s.a_off = a_off; s.out_off = out_off; s.w_off = p.blk[blk].w_off; s.w_stride = p.blk[blk].sub_bytes;
        s.kp = (uint16_t)kp; s.n = (uint16_t)n; s.ck = (uint16_t)chunk_rows(kp, n);
        s.tmem_col = (uint16_t)col; s.out_c0 = (uint16_t)out_c0; s.ss_idx = (uint16_t)ss_idx;
        s.acc = (uint8_t)acc; s.epi = (uint8_t)epi;
    };
    add_step(off_x, off_h, b0, p.kp0, n0c, 0, 0, 0, 0, 1);
    for (int fh = 0; fh < h2; ++fh)
        for (int kh = 0; kh < h1; ++kh) {
            // the layer-1 half's output lands in buffer X at K offset 0 (it is the whole K range of the next step)
            add_step(off_h, p.reuse_h ? off_h : off_x, b1[kh], n0, n1c, l1col, 0, n0 + kh * n1h, 0, 1);
            add_step(p.reuse_h ? off_h : off_x, 0, b2[kh][fh], n1h, n2c, 0, fh * n2h, n0 + n1 + fh * n2h, kh > 0, kh == h1 - 1 ? 2 : 0);
        }
    const int cols = (split ? l1col + n1c : 0) > max_nc ? l1col + n1c : max_nc;
    p.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
    p.ok = true;
    return p;
}

bool sa_mlp_tc_supported(int D, const int* cout) { return tc_plan(D, cout, 1).ok; }

// CTAs per row tile: split N over a cluster while the launch would otherwise leave most SMs idle.
int sa_mlp_tc_pick_cluster(long tiles, int D, const int* cout) {
    int best = 1;
    for (int C = 2; C <= 8; C *= 2)
        if (tiles * C <= num_sms() && tc_plan(D, cout, C).ok) best = C;
    return best;
}
size_t sa_mlp_tc_blob_bytes(int D, const int* cout, int C) {
    const TcPlan p = tc_plan(D, cout, C);
    return p.ok ? p.blob_bytes : 0;
}

int sa_mlp_tc_pack(const pcst_mlp3_t* mlp, int D, int C, void* blob, cudaStream_t stream) {
    const TcPlan p = tc_plan(D, mlp->cout, C);
    if (!p.ok) {
        set_error("sa_mlp pack (tensor-core path): unsupported layer widths / cluster size");
        return PCST_ERR_UNSUPPORTED;
    }
    const int cin[3] = {3 + D, p.n[0], p.n[1]};
    for (int i = 0; i < p.nblocks; ++i) {
        const TcBlock& b = p.blk[i];
        const int nc = b.nlen / C;
        const int total = b.kp * nc;
        for (int r = 0; r < C; ++r) {
            tc_pack_block_kernel<<<(total + 255) / 256, 256, 0, stream>>>(
                mlp->w[b.layer], cin[b.layer], b.k0, b.klen, b.kp, b.n0 + r * nc, nc, b.layer == 0 ? D : -1,
                reinterpret_cast<__nv_bfloat16*>((char*)blob + b.w_off + (size_t)r * b.sub_bytes));
            PCST_CUDA(cudaGetLastError());
        }
    }
    float* ss = reinterpret_cast<float*>((char*)blob + p.ss_blob_off);
    int at = 0;
    for (int l = 0; l < 3; ++l) {
        tc_pack_ss_kernel<<<(p.n[l] + 255) / 256, 256, 0, stream>>>(mlp->scale[l], mlp->shift[l], p.n[l], at, p.total_ch, ss);
        PCST_CUDA(cudaGetLastError());
        at += p.n[l];
    }
    tc_pack_fence_kernel<<<1, 32, 0, stream>>>();
    PCST_CUDA(cudaGetLastError());
    return PCST_OK;
}

static unsigned long long* g_tc_probe = nullptr;
static int g_tc_probe_tiles = 0;
void sa_mlp_tc_set_probe(unsigned long long* buf, int tiles) {
    g_tc_probe = buf;
    g_tc_probe_tiles = buf ? tiles : 0;
}

// bf16 copy of the feature tensor: worth its extra launch when every feature row is gathered many times and the launch is
// not on a latency-critical path (many more row tiles than SMs)
static bool tc_wants_bf16_feats(size_t rows, int B, int N, int D) {
    return D >= 32 && (D % 8) == 0 && (rows + kTcM - 1) / kTcM > (size_t)2 * num_sms() && rows >= (size_t)4 * B * N &&
           tuning("sa_mlp.bf16_feats", 0) != 2;
}
size_t sa_mlp_tc_workspace_bytes(int B, int N, int S, int K, int D) {
    return tc_wants_bf16_feats((size_t)B * S * K, B, N, D) ? align_up((size_t)B * N * D * 2, 256) : 0;
}
int sa_mlp_tc_launches(int B, int N, int S, int K, int D) { return tc_wants_bf16_feats((size_t)B * S * K, B, N, D) ? 2 : 1; }

int sa_mlp_tc_run(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                  int K, int D, const int* cout, int C, const void* blob, float* out, void* ws, size_t ws_bytes,
                  cudaStream_t stream) {
    const size_t rows_sz = (size_t)B * S * K;
    const TcPlan p = tc_plan(D, cout, C, /*dense=*/(rows_sz + kTcM - 1) / kTcM > (size_t)2 * num_sms());
    if (!p.ok) {
        set_error("sa_mlp_max (tensor-core path): unsupported layer widths / cluster size");
        return PCST_ERR_UNSUPPORTED;
    }
    if (rows_sz >= (1u << 30)) {
        set_error("sa_mlp_max: B*S*K too large");
        return PCST_ERR_INVALID;
    }
    TcArgs a = {};
    a.xyz = xyz; a.feats = feats; a.new_xyz = new_xyz; a.idx = idx;
    a.N = N; a.S = S; a.K = K; a.D = D; a.rows = (int)rows_sz;
    a.feats_bf16 = nullptr;
    if (feats && ws && ((uintptr_t)ws & 255) == 0 && ((uintptr_t)feats & 15) == 0 && tc_wants_bf16_feats(rows_sz, B, N, D) &&
        ws_bytes >= sa_mlp_tc_workspace_bytes(B, N, S, K, D)) {
        const size_t n8 = (size_t)B * N * D / 8;
        unsigned blocks = (unsigned)((n8 + 255) / 256);
        if (blocks > 8u * (unsigned)num_sms()) blocks = 8u * (unsigned)num_sms();
        tc_feats_to_bf16_kernel<<<blocks, 256, 0, stream>>>(feats, (uint4*)ws, n8);
        PCST_CUDA(cudaGetLastError());
        a.feats_bf16 = (const __nv_bfloat16*)ws;
    }
    a.blob = (const unsigned char*)blob;
    a.ss_blob_off = p.ss_blob_off;
    a.total_ch = p.total_ch; a.kp0 = p.kp0; a.cout = cout[2];
    a.nsteps = p.nsteps;
    for (int s = 0; s < p.nsteps; ++s) a.st[s] = p.st[s];
    a.off_ring = p.off_ring; a.stage_bytes = p.stage_bytes; a.nstages = p.nstages;
    a.off_ss = p.off_ss; a.off_pool = p.off_pool; a.tmem_cols = p.tmem_cols;
    a.out = out;
    a.pool_atomic = !(K == 32 || K == 64 || K == 128);
    a.cluster = (uint32_t)C;
    a.early_gather = p.reuse_h && p.nsteps == 3 && tuning("sa_mlp.early_gather", 0) != 2;
    {   // layer-2 steps (the ones that are not operand-writing) all have the same per-CTA width
        int n2c = 0;
        for (int s = 0; s < p.nsteps; ++s)
            if (p.st[s].epi == 2) n2c = p.st[s].n;
        a.transpose_pool = !a.pool_atomic && n2c > 0 && (n2c % 128) == 0 && tuning("sa_mlp.transpose_pool", 0) != 2;
    }
    if (a.pool_atomic) PCST_CUDA(cudaMemsetAsync(out, 0, (size_t)B * S * cout[2] * sizeof(float), stream));
    const int tiles = (a.rows + kTcM - 1) / kTcM;
    // the narrow build pays off when tiles queue up on every SM and the stage's shared memory allows three or four CTAs
    const int regs = tuning("sa_mlp.regs", 0);  // 0 = auto, else 80 / 168 (A/B measurements)
    const bool narrow = regs ? regs == 80 : (tiles > 2 * num_sms() && (p.smem_bytes + 1024u) * 3u <= 228u * 1024u);
    auto kernel = narrow ? sa_mlp_tc_kernel<80, false> : sa_mlp_tc_kernel<168, true>;
    a.probe = g_tc_probe; a.probe_tiles = g_tc_probe_tiles;
    a.relaxed = tiles > num_sms() && tuning("sa_mlp.backoff", 1) == 1;
    PCST_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    a.ntiles = tiles;
    // More tiles than the machine holds at once (C == 1 by construction): as many CTAs as are co-resident, each walking
    // tiles with stride gridDim -- barriers, TMEM and the scale/shift tables are set up once and the weight ring keeps
    // streaming across tile boundaries.  TMEM columns bound the co-residency too (the occupancy API does not know).
    unsigned grid = (unsigned)tiles * C;
    // Plans with N / K halves (layers wider than 256) keep one CTA per tile: their step tables mix partial-sum and pooled
    // steps mid-sequence, and the walk has only been exercised on GPUs with the plain three-step table.
    if (C == 1 && !narrow && p.nsteps == 3 && tiles > num_sms() && tuning("sa_mlp.persistent", 1) == 1) {
        // co-resident CTAs per SM from the kernel's own footprint: shared memory (228 KiB per SM at the maximum
        // carve-out, 1 KiB reserved per CTA), registers (64 Ki per SM, allocated per warp in units of 256) and TMEM
        // columns (512 per SM, which the occupancy API does not model)
        cudaFuncAttributes fa;
        PCST_CUDA(cudaFuncGetAttributes(&fa, kernel));
        const int regs_per_cta = (kTcThreads / 32) * ((fa.numRegs * 32 + 255) / 256 * 256);
        int occ = (int)(228u * 1024u / (p.smem_bytes + (uint32_t)fa.sharedSizeBytes + 1024u));
        if (occ > 65536 / regs_per_cta) occ = 65536 / regs_per_cta;
        if (occ > 512 / (int)p.tmem_cols) occ = 512 / (int)p.tmem_cols;
        if (occ < 1) occ = 1;
        if ((unsigned)(occ * num_sms()) < grid) grid = (unsigned)(occ * num_sms());
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = tuning("sa_mlp.pdl", 1) == 1 ? 2 : 1;  // 2 = off (A/B measurements)
    PCST_CUDA(cudaLaunchKernelEx(&cfg, kernel, a));
    return PCST_OK;
}

}  // namespace pcst
