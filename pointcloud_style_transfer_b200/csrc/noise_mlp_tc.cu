// noise_mlp_tc.cu -- NoisePredictor (models/diffusion_model.py:38-61) as ONE fused tcgen05 kernel per call (sm_100a).
//
// The reference evaluates, per point, a chain of nn.Linear layers on cuBLAS with separate bias / ReLU / residual kernels:
//   point_feat = Linear(3,128) -> ReLU -> Linear(128,256) -> ReLU -> Linear(256,F)
//   x = point_feat + time_proj(TimeEmbedding(t)) + style_proj(style)              (per batch element vectors, :55-57)
//   6 x:  x = Linear(2F,F)(ReLU(Linear(F,2F)(x))) + x                             (Dropout is identity in eval, :58-59)
//   out = Linear(F,256) -> ReLU -> Linear(256,128) -> ReLU -> Linear(128,3)
// = 3.5 MFLOP per point and ~20 activation round trips through HBM.  Here a CTA owns 128 points and runs the whole chain
// with activations in shared memory (bf16 K-major UMMA operands) and TMEM:
//   * the residual stream x lives in TMEM columns [0, F) as the fp32 ACCUMULATOR for the whole chain: the second Linear of
//     every block accumulates straight onto it (tcgen05.mma with accumulate), so the residual add costs nothing and x is
//     never rounded to bf16 except when it is read as the next GEMM's operand;
//   * every bias that lands on x (point_encoder.4, time / style projections, the blocks' second biases) is folded into a
//     per-(batch element, stage) shift vector computed once per call by a tiny prep kernel; the operand read of stage i
//     adds shift_i;
//   * the hidden 2F activations are produced and consumed in CHUNKS of at most 128 columns that alternate between two
//     TMEM accumulator slots (columns [256, 384) and [384, 512)) and two 32 KiB operand slots in shared memory, and the
//     MMA issuer runs one chunk ahead:  W1(c0) W1(c1) W2(c0) W1(c2) W2(c1) W1(c3) W2(c2) W2(c3)  -- the epilogue of a
//     chunk (TMEM -> +bias -> ReLU -> bf16 -> shared) runs under the MMAs of the next one instead of in series with them.
//     Every step carries the number of epilogue events the issuer must have seen before it may run (its operand written,
//     the accumulator slot it overwrites read out); the host derives these counts from the buffers a step touches;
//   * eight epilogue warps: a warp reads only the TMEM lanes of its quadrant (warp % 4), so two warps share a quadrant
//     and split an accumulator's 16-column groups between them; biases and shift vectors are staged in shared memory once;
//   * weights are packed once (bf16 [K/8][N][8] blocks in step order) and streamed through a 4-stage TMA ring; they stay
//     L2-resident (3.3 MB for F = 256).  Every 128-row tile needs ALL of them (3.3 MB per tile, 64 bytes per SM and clock at
//     the tensor floor -- more than the L2 delivers to 148 SMs at once), so the kernel can run as thread-block clusters of
//     2 or 4 CTAs that walk their tiles in lockstep and share every stage: each CTA fetches 1/C of it and MULTICASTS it
//     into all shared memories (cp.async.bulk ... .multicast::cluster), and a stage is refilled once all MMA issuers have
//     committed it (tcgen05.commit.multicast::cluster onto every CTA's empty barrier).  Tuning key noise.cluster.
// Warp roles: warps 0-7 = operand build + epilogues (thread = row = TMEM lane), warp 8 = TMEM allocation + weight
// producer, warp 9 = MMA issuer.
// Precision: bf16 operands, fp32 accumulation / residual stream / biases -> within rtol 2e-2 of the reference's fp32 module.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pcst {

constexpr int kNpMaxSteps = 80;
constexpr int kNpMaxStages = 6;   // weight-ring stages: as many as fit beside the operands (tuning noise.stages)
constexpr int kNpStageBytes = 16 * 1024;
constexpr int kNpMaxBlocks = 8;
constexpr int kNpEpiWarps = 8;
constexpr int kNpEpiThreads = kNpEpiWarps * 32;
constexpr int kNpThreads = kNpEpiThreads + 64;
constexpr int kNpSlotBytes = kTcM * 128 * 2;   // a 128-row x 128-K bf16 operand slot
constexpr int kNpSlotCols = 128;               // TMEM columns of a hidden accumulator slot
constexpr int kNpHiddenCol = 256;              // first TMEM column of the two slots (x occupies [0, F))

struct NpStep {
    uint32_t a_off;    // shared-memory offset of the A operand
    uint32_t out_off;  // epi 1, 2: operand buffer written
    uint32_t w_off;    // offset of the step's weight block [kp/8][n][8] in the blob
    uint32_t bias_off; // epi 1, 3: first bias (float index into the bias table)
    uint16_t kp, n, tmem_col;
    uint16_t wait_ev;  // epilogue events (the input operand is event 0) the MMA issuer must have seen before this step
    uint8_t acc;       // accumulate onto the TMEM contents
    uint8_t epi;       // 0 none; 1 relu(acc + bias) -> bf16 operand; 2 (acc + shift[stage]) -> bf16 operand; 3 acc + bias -> out
    uint8_t stage;     // epi 2: which shift vector
    uint8_t bar;       // epi != 0: accumulator barrier committed after the step (0 / 1 = hidden slots, 2 = x)
    // the issue loop's constants, precomputed: the lone issuing warp retires ~1 instruction per 4-5 cycles, so every
    // instruction between two tcgen05.mma counts against the 64-cycle floor of an M = 128, N = 128 MMA
    uint32_t idesc;    // instruction descriptor of the step's MMAs
    uint8_t nchunks;   // weight-ring stages the step consumes
    uint8_t nk;        // K = 16 slices (MMAs) per full stage
    uint8_t nk_last;   // ... in the last stage
    uint8_t pad_;
};

struct NpArgs {
    const float* pts;   // [B, N, 3]
    float* out;         // [B, N, 3]
    int B, N, F, tiles_per_b;
    const unsigned char* blob;
    const float* bias;   // bias table inside the blob
    const float* shift;  // [B][nstage][F]
    int nstage, nsteps, bias_floats;
    uint32_t off_ring, off_vec;   // shared-memory offsets: weight ring; bias table followed by the batch element's shifts
    int nstages;                  // ring stages in use
    int vec_in_smem;              // biases / shifts staged in shared memory (else read through L1 from global memory)
    uint32_t cluster;    // CTAs sharing every weight stage by multicast (1, 2 or 4)
    int ntiles;          // real row tiles; CTAs past them (the grid is padded to whole clusters) only keep the lockstep
    int probe;           // timing probe: 64 = clock stamps of the first steps, printed by one CTA
    NpStep st[kNpMaxSteps];
};

__host__ __device__ __forceinline__ int np_chunk_rows(int n, int kp) {
    int ck = kNpStageBytes / (n * 2) / 16 * 16;
    if (ck < 16) ck = 16;
    return ck < kp ? ck : kp;
}

// Walk the 16-column groups first, first + 2, ... (< ngroups) of an accumulator in batches of four: the batch's loads are
// all in flight before the one wait (an epilogue warp has a single partner on its scheduler, so exposed TMEM-load
// latency per group is what an epilogue costs).
template <typename F>
__device__ __forceinline__ void tmem_for_each16_alternate(uint32_t taddr, int first, int ngroups, F&& f) {
    for (int g0 = first; g0 < ngroups; g0 += 8) {
        uint32_t r[4][16];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (g0 + 2 * j < ngroups) tmem_ld16_issue(taddr + (uint32_t)(g0 + 2 * j) * 16u, r[j]);
#pragma unroll
        for (int j = 0; j < 4; ++j) tmem_ld_wait(r[j]);   // one real wait; the rest only tie the registers to it
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (g0 + 2 * j < ngroups) f(r[j], (g0 + 2 * j) * 16);
    }
}

__global__ void __launch_bounds__(kNpThreads)
noise_mlp_kernel(const __grid_constant__ NpArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t full_bar[kNpMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kNpMaxStages];
    __shared__ __align__(8) uint64_t acc_bar[3];  // an epilogue-bearing step's accumulator is complete (slot 0, slot 1, x)
    __shared__ __align__(8) uint64_t a_bar[2];    // epilogue event e (operand written / accumulator read out) -> a_bar[e & 1]
    __shared__ uint32_t tmem_base_sh;
    // the step table is walked by every role with a run-time index: from shared memory (an indexed read of the kernel
    // parameters is a constant-cache access per field and was measured at ~2 000 cycles per step)
    __shared__ __align__(16) NpStep steps[kNpMaxSteps];
    __shared__ long long stamp[5][kNpMaxSteps];   // probe 64: issuer commit, epilogue wake, epilogue arrive, issuer wake,
                                                  // cycles the issuer waited for weight stages (per step)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler: role branches and their loop
                                                              // counters stay in uniform registers
    const long long t_entry = clock64();
    for (int i = tid; i < a.nsteps * (int)(sizeof(NpStep) / 4); i += kNpThreads)
        reinterpret_cast<uint32_t*>(steps)[i] = reinterpret_cast<const uint32_t*>(a.st)[i];
    const uint32_t C = a.cluster;
    const uint32_t rank = C > 1 ? cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << C) - 1u);
    const bool real_tile = (int)blockIdx.x < a.ntiles;
    const int b = real_tile ? blockIdx.x / a.tiles_per_b : 0;
    const int row0 = real_tile ? (blockIdx.x % a.tiles_per_b) * kTcM : a.N;   // a padding CTA owns no valid row
    if (tid == 0) {
        for (int s = 0; s < kNpMaxStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], C);   // every CTA of the cluster commits to every CTA's empty barrier
        }
        for (int s = 0; s < 3; ++s) mbar_init(&acc_bar[s], 1);
        mbar_init(&a_bar[0], kNpEpiWarps);   // one arrival per epilogue WARP: hundreds of per-thread arrivals on one
        mbar_init(&a_bar[1], kNpEpiWarps);   // mbarrier serialise (measured ~1 us per event with 256 of them)
        fence_mbar_init();
    }
    if (warp == kNpEpiWarps) tmem_alloc(&tmem_base_sh, 512);
    // biases of the whole chain and this batch element's shift vectors: read by every epilogue, staged once
    const float* bias_v = a.bias;
    const float* shift_v = a.shift + (size_t)b * a.nstage * a.F;
    if (a.vec_in_smem) {
        float* vec = reinterpret_cast<float*>(smem + a.off_vec);
        for (int i = tid; i < a.bias_floats; i += kNpThreads) vec[i] = __ldg(bias_v + i);
        for (int i = tid; i < a.nstage * a.F; i += kNpThreads) vec[a.bias_floats + i] = __ldg(shift_v + i);
        bias_v = vec;
        shift_v = vec + a.bias_floats;
    }
    const uint32_t nstages = (uint32_t)a.nstages;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (C > 1) cluster_sync_all();  // every peer's barriers exist before anything is multicast to them
    const uint32_t tmem_base = tmem_base_sh;
    const long long t_setup = clock64();
    // shared-window addresses formed once (see tc_common.cuh)
    const uint32_t smem_base = tc_opaque_u32(smem_u32(smem));
    const uint32_t full0 = tc_opaque_u32(smem_u32(&full_bar[0])), empty0 = tc_opaque_u32(smem_u32(&empty_bar[0]));
    const uint32_t acc0 = tc_opaque_u32(smem_u32(&acc_bar[0])), abar0 = tc_opaque_u32(smem_u32(&a_bar[0]));

    if (warp == kNpEpiWarps) {
        if (elect_one_sync()) {
            uint32_t stage = 0, round = 0;   // ring position; how many times the ring has wrapped
            for (int s = 0; s < a.nsteps; ++s) {
                const NpStep& st = a.st[s];
                const uint32_t nchunks = st.nchunks, slice_bytes = (uint32_t)st.n * 32u;   // one K = 16 slice of the weights
                const unsigned char* src = a.blob + st.w_off;
                for (uint32_t c = 0; c < nchunks; ++c) {
                    if (round > 0) mbar_wait_addr(empty0 + stage * 8u, (round - 1u) & 1u);
                    const uint32_t bytes = (c + 1 == nchunks ? st.nk_last : st.nk) * slice_bytes;
                    mbar_expect_tx_addr(full0 + stage * 8u, bytes);   // the whole stage: this CTA's slice + the peers'
                    const uint32_t dst = smem_base + a.off_ring + stage * kNpStageBytes;
                    if (C > 1) {
                        const uint32_t slice = bytes / C;   // bytes is a multiple of 512
                        tma_load_1d_multicast_addr(dst + rank * slice, src + rank * slice, slice, full0 + stage * 8u, cmask);
                    } else {
                        tma_load_1d_addr(dst, src, bytes, full0 + stage * 8u);
                    }
                    src += bytes;
                    if (++stage == nstages) { stage = 0; ++round; }
                }
            }
        }
    } else if (warp == kNpEpiWarps + 1) {
        // The whole warp walks the loop converged and one elected lane issues: every operand of tcgen05.mma / commit is then
        // warp-uniform for the compiler (kernel parameters with a uniform index, uniform counters), which keeps the
        // descriptors in uniform registers.  Issued from a lone lane of a divergent branch the same code costs a
        // register -> uniform-register move per operand and an ELECT / BRA.U.ANY waterfall loop around every UTCHMMA:
        // measured ~200 cycles per MMA against a tensor-pipe floor of 64.
        const bool leader = elect_one_sync();
        const uint32_t desc_hi = (128u >> 4) | (1u << 14);              // SBO = 128 B; descriptor version 1 (bit 46)
        const uint32_t a_step = (2u * kTcM * 16u) >> 4;                 // one K = 16 slice of A: two K groups of 128 rows
        const uint32_t ring_lo = (smem_base + a.off_ring) >> 4;
        uint32_t stage = 0, round = 0, seen = 0;
        for (int s = 0; s < a.nsteps; ++s) {
            const NpStep& st = a.st[s];
            const uint32_t wait_ev = st.wait_ev;
            while (seen < wait_ev) {   // events complete in order; event e lives on a_bar[e & 1], phase e >> 1
                mbar_wait_addr(abar0 + (seen & 1u) * 8u, (seen >> 1) & 1u);
                ++seen;
            }
            if ((a.probe & 64) && leader) stamp[3][s] = clock64();
            tc_fence_after();
            const uint32_t n = st.n, idesc = st.idesc;
            const uint32_t d_addr = tmem_base + st.tmem_col;
            const uint32_t w_step = n * 2u;                              // one K = 16 slice of the weights, 16-byte units
            const uint32_t w_lbo = n << 16;                              // LBO = n * 16 bytes
            uint32_t a_lo = (((smem_base + st.a_off) >> 4) & 0x3FFFu) | (((kTcM * 16u) >> 4) << 16);
            uint32_t accum = st.acc;
            const uint32_t nchunks = st.nchunks, nk_full = st.nk, nk_last = st.nk_last;
            long long waited = 0;
            for (uint32_t c = 0; c < nchunks; ++c) {
                if (a.probe & 64) {
                    const long long t0 = clock64();
                    mbar_wait_addr(full0 + stage * 8u, round & 1u);
                    waited += clock64() - t0;
                } else {
                    mbar_wait_addr(full0 + stage * 8u, round & 1u);
                }
                tc_fence_after();
                const uint32_t nk = c + 1 == nchunks ? nk_last : nk_full;
                if (leader) {
                    const uint32_t w_lo = ((ring_lo + stage * (kNpStageBytes >> 4)) & 0x3FFFu) | w_lbo;
                    auto mma = [&](uint32_t kk, bool acc) {
                        umma_bf16(d_addr, ((uint64_t)desc_hi << 32) | (a_lo + kk * a_step),
                                  ((uint64_t)desc_hi << 32) | (w_lo + kk * w_step), idesc, acc);
                    };
                    if (nk == 4) {          // N = 128: the common stage, straight-line
                        mma(0, accum != 0); mma(1, true); mma(2, true); mma(3, true);
                    } else if (nk == 2) {   // N = 256
                        mma(0, accum != 0); mma(1, true);
                    } else {
                        for (uint32_t kk = 0; kk < nk; ++kk) mma(kk, (accum | kk) != 0);
                    }
                    if (C > 1) umma_commit_multicast_addr(empty0 + stage * 8u, cmask);
                    else umma_commit_addr(empty0 + stage * 8u);
                }
                __syncwarp();
                a_lo += nk * a_step;
                accum = 1;
                if (++stage == nstages) { stage = 0; ++round; }
            }
            if ((a.probe & 64) && leader) {
                stamp[4][s] = waited;
                stamp[0][s] = clock64();
            }
            if (st.epi && leader) umma_commit_addr(acc0 + st.bar * 8u);
            __syncwarp();
        }
    } else {
        const int quad = warp & 3, part = warp >> 2;   // TMEM lane quadrant; which of the quadrant's two warps
        const int m = quad * 32 + lane, row = row0 + m;
        const bool valid = row < a.N;
        // input operand: K = 16 = (x, y, z, 0 ...), two 16-byte K groups per row
        if (part == 0) {
            float x = 0.f, y = 0.f, z = 0.f;
            if (valid) {
                const float* p = a.pts + ((size_t)b * a.N + row) * 3;
                x = p[0]; y = p[1]; z = p[2];
            }
            unsigned char* A0 = smem + steps[0].a_off;
            __nv_bfloat162 xy = __floats2bfloat162_rn(x, y), z0 = __floats2bfloat162_rn(z, 0.f);
            *reinterpret_cast<uint4*>(A0 + (size_t)m * 16) =
                make_uint4(*reinterpret_cast<uint32_t*>(&xy), *reinterpret_cast<uint32_t*>(&z0), 0u, 0u);
            *reinterpret_cast<uint4*>(A0 + ((size_t)kTcM + m) * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(abar0);   // event 0

        const uint32_t taddr_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint32_t ev = 1, acc_phase = 0;   // bit i of acc_phase = parity of acc_bar[i]'s next completion
        for (int s = 0; s < a.nsteps; ++s) {
            const NpStep st = steps[s];
            if (!st.epi) continue;
            mbar_wait_addr(acc0 + st.bar * 8u, (acc_phase >> st.bar) & 1u);
            if ((a.probe & 64) && tid == 0) stamp[1][s] = clock64();
            acc_phase ^= 1u << st.bar;
            tc_fence_after();
            const uint32_t taddr = taddr_lane + st.tmem_col;
            if (st.epi == 3) {
                if (part == 0) {
                    uint32_t rr[16];
                    tmem_ld16_issue(taddr, rr);
                    tmem_ld_wait(rr);
                    if (valid) {
                        float* o = a.out + ((size_t)b * a.N + row) * 3;
                        const float* bs = bias_v + st.bias_off;
                        o[0] = __uint_as_float(rr[0]) + bs[0];
                        o[1] = __uint_as_float(rr[1]) + bs[1];
                        o[2] = __uint_as_float(rr[2]) + bs[2];
                    }
                }
                continue;
            }
            // per-channel additive vector of this epilogue: a bias (epi 1) or the batch element's shift of the stage (epi 2)
            const float* add = st.epi == 1 ? bias_v + st.bias_off : shift_v + (int)st.stage * a.F;
            unsigned char* outp = smem + st.out_off;
            const bool relu = st.epi == 1;
            tmem_for_each16_alternate(taddr, part, st.n / 16, [&](uint32_t (&r)[16], int cb) {
                uint32_t packed[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 ad = *reinterpret_cast<const float4*>(add + cb + 4 * i);
                    float y0 = __uint_as_float(r[4 * i]) + ad.x, y1 = __uint_as_float(r[4 * i + 1]) + ad.y;
                    float y2 = __uint_as_float(r[4 * i + 2]) + ad.z, y3 = __uint_as_float(r[4 * i + 3]) + ad.w;
                    if (relu) {
                        y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); y2 = fmaxf(y2, 0.f); y3 = fmaxf(y3, 0.f);
                    }
                    __nv_bfloat162 p0 = __floats2bfloat162_rn(y0, y1), p1 = __floats2bfloat162_rn(y2, y3);
                    packed[2 * i] = *reinterpret_cast<uint32_t*>(&p0);
                    packed[2 * i + 1] = *reinterpret_cast<uint32_t*>(&p1);
                }
                *reinterpret_cast<uint4*>(outp + ((size_t)(cb >> 3) * kTcM + m) * 16) =
                    make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *reinterpret_cast<uint4*>(outp + ((size_t)((cb >> 3) + 1) * kTcM + m) * 16) =
                    make_uint4(packed[4], packed[5], packed[6], packed[7]);
            });
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if ((a.probe & 64) && tid == 0) stamp[2][s] = clock64();
            if (lane == 0) mbar_arrive_addr(abar0 + (ev & 1u) * 8u);
            ++ev;
        }
    }
    const long long t_role = clock64();
    tc_fence_before();
    __syncthreads();
    if (C > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast or commit into it
    if (warp == kNpEpiWarps) tmem_dealloc(tmem_base, 512);
    if ((a.probe & 64) && lane == 0 && (blockIdx.x == 0 || blockIdx.x == 300))
        printf("cta %d warp %d: setup %lld role %lld exit %lld cycles\n", (int)blockIdx.x, warp, t_setup - t_entry,
               t_role - t_setup, clock64() - t_role);
    if ((a.probe & 64) && tid == 0 && blockIdx.x == 300)
        for (int s2 = 0; s2 < a.nsteps && s2 < 24; ++s2)
            printf("step %d epi %d wait_ev %d: issuer woke %lld done %lld (waited for weights %lld) | epilogue woke %lld arrived %lld\n",
                   s2, (int)steps[s2].epi, (int)steps[s2].wait_ev, stamp[3][s2] - t_setup, stamp[0][s2] - t_setup, stamp[4][s2],
                   stamp[1][s2] - t_setup, stamp[2][s2] - t_setup);
}

// ---- prep: per batch element, shift[b][0] = b_pe2 + time_proj(TimeEmbedding(t_b)) + style_proj(style_b);
//            shift[b][i] = shift[b][i-1] + b2_{i-1}    (models/diffusion_model.py:15-26,55-59)
// grid (F / 8, B), one warp per output channel: the rows of the two projection matrices are read coalesced.
__global__ void __launch_bounds__(256)
noise_prep_kernel(const long long* __restrict__ timestep, const float* __restrict__ style, int F, int T, int nblocks,
                  const float* __restrict__ time_w, const float* __restrict__ time_b, const float* __restrict__ style_w,
                  const float* __restrict__ style_b, const float* __restrict__ pe2_b, const float* __restrict__ b2 /*[nblocks][F]*/,
                  float* __restrict__ shift /*[B][nblocks + 1][F]*/) {
    extern __shared__ float emb[];  // [T]
    const int b = blockIdx.y;
    const int half = T / 2;
    const float t = (float)timestep[b];
    const float scale = logf(10000.0f) / (float)(half - 1);
    for (int j = threadIdx.x; j < half; j += blockDim.x) {
        const float f = expf((float)j * -scale);
        const float v = t * f;
        emb[j] = sinf(v);
        emb[half + j] = cosf(v);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= F) return;
    // the reference's two dot products, each summed in index order per lane stripe and then across lanes: not the
    // sequential order of a CPU GEMV, well inside the kernel's bf16 tolerance
    float acc = 0.f, acs = 0.f;
    for (int j = lane; j < T; j += 32) acc = fmaf(time_w[(size_t)c * T + j], emb[j], acc);
    for (int j = lane; j < F; j += 32) acs = fmaf(style_w[(size_t)c * F + j], style[(size_t)b * F + j], acs);
    float v = acc + acs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
        float s = pe2_b[c] + time_b[c] + style_b[c] + v;
        float* o = shift + (size_t)b * (nblocks + 1) * F + c;
        o[0] = s;
        for (int i = 0; i < nblocks; ++i) {
            s += b2[(size_t)i * F + c];
            o[(size_t)(i + 1) * F] = s;
        }
    }
}

// fp32 W[n0 + n, k0 + k] (row-major [rows, cin]) -> bf16 [kp/8][nlen][8], zero padded in K and N
__global__ void noise_pack_block_kernel(const float* __restrict__ w, int rows, int cin, int n0, int nlen, int k0, int klen, int kp,
                                        __nv_bfloat16* __restrict__ out) {
    const int total = kp * nlen;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kc = e / (nlen * 8);
        const int rem = e % (nlen * 8);
        const int n = rem / 8, k = kc * 8 + (rem % 8);
        float v = 0.f;
        if (n0 + n < rows && k < klen) v = w[(size_t)(n0 + n) * cin + k0 + k];
        out[e] = __float2bfloat16_rn(v);
    }
}
__global__ void noise_copy_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int pad_to) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < pad_to; i += gridDim.x * blockDim.x) dst[i] = i < n ? src[i] : 0.f;
}

// ---- host: the step table and the blob layout -------------------------------------------------------------------------------
struct NpPlan {
    bool ok;
    int F, T, nblocks, nsteps;
    NpStep st[kNpMaxSteps];
    // which reference matrix a step's weight block is cut from
    struct Src { int which, block, n0, nlen, k0, klen; } src[kNpMaxSteps];  // which: 0-2 pe, 3 blk_w1, 4 blk_w2, 5-7 out
    uint32_t bias_floats;  // bias table: pe0 | pe1 | blk b1 x nblocks | out0 | out1 | out2(16)
    uint32_t off_bias, off_time_w, off_time_b, off_style_w, off_style_b, off_pe2_b, off_b2;
    size_t blob_bytes;
    uint32_t off_x, off_h, off_ring, off_vec, smem_bytes;
    int nstages, vec_in_smem;
    int pack_launches;
};

static NpPlan np_plan(int F, int T, int nblocks, int stages_wanted = 0) {
    NpPlan p = {};
    p.ok = false;
    if (F < 16 || F > 256 || (F % 16) != 0 || T < 4 || (T % 2) != 0 || T > 1024 || nblocks < 0 || nblocks > kNpMaxBlocks) return p;
    p.F = F; p.T = T; p.nblocks = nblocks;
    p.off_x = 0;                       // the x operand, K <= 256: two slots
    p.off_h = 2 * kNpSlotBytes;        // the two hidden operand slots (contiguous: together one K = 256 operand)
    p.off_ring = 4 * kNpSlotBytes;
    // ring depth: measured (profiles/r02/noise_stages_ab.log) 3 / 4 / 5 stages = 327 / 317 / 340 us per call: the stream is
    // not bound by bytes in flight, and the fifth stage only fits by evicting the biases and shifts from shared memory
    // (which costs more than it gains): as many stages as fit BESIDE the staged vectors, unless tuned
    p.off_vec = 0;
    uint32_t woff = 0, boff = 0;
    // hazard bookkeeping: which epilogue event last wrote a shared-memory range / last read a TMEM column range
    struct Range { uint32_t lo, hi; int ev; };
    Range smem_w[2 * kNpMaxSteps]; int n_sw = 0;
    Range tmem_r[2 * kNpMaxSteps]; int n_tr = 0;
    int events = 1;                    // event 0 = the input operand
    smem_w[n_sw++] = {p.off_x, p.off_x + 2u * kTcM * 16u, 1};
    int seen = 0;                      // events the issuer has waited for so far (waits are cumulative)
    auto add = [&](uint32_t a_off, uint32_t out_off, int kp, int n, int col, int acc, int epi, int stage, uint32_t bias_off,
                   int which, int block, int n0, int nlen, int k0, int klen) {
        NpStep& s = p.st[p.nsteps];
        s.a_off = a_off; s.out_off = out_off; s.w_off = woff; s.bias_off = bias_off;
        s.kp = (uint16_t)kp; s.n = (uint16_t)n; s.tmem_col = (uint16_t)col;
        s.acc = (uint8_t)acc; s.epi = (uint8_t)epi; s.stage = (uint8_t)stage;
        s.bar = (uint8_t)(col == 0 ? 2 : col == kNpHiddenCol ? 0 : 1);
        {
            const int ck = np_chunk_rows(n, kp);
            s.nchunks = (uint8_t)((kp + ck - 1) / ck);
            s.nk = (uint8_t)(ck / 16);
            s.nk_last = (uint8_t)((kp - (s.nchunks - 1) * ck) / 16);
            s.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);  // = umma_idesc_bf16
        }
        // RAW on the operand (its latest writer) and WAR / RMW on the accumulator columns (their latest reader)
        int need = 0;
        const uint32_t alo = a_off, ahi = a_off + (uint32_t)kp * kTcM * 2u;
        for (int i = 0; i < n_sw; ++i)
            if (smem_w[i].lo < ahi && alo < smem_w[i].hi && smem_w[i].ev > need) need = smem_w[i].ev;
        for (int i = 0; i < n_tr; ++i)
            if (tmem_r[i].lo < (uint32_t)(col + n) && (uint32_t)col < tmem_r[i].hi && tmem_r[i].ev > need) need = tmem_r[i].ev;
        if (epi == 1 || epi == 2) {
            // this step's epilogue is event `events`; two events share a barrier two apart: the older one must have been
            // seen by the issuer before the newer one can complete, i.e. before this step is issued
            if (events >= 2 && need < events - 1) need = events - 1;
        }
        if (need > seen) seen = need;
        s.wait_ev = (uint16_t)seen;
        if (epi == 1 || epi == 2) {
            const int e = events++;
            smem_w[n_sw++] = {out_off, out_off + (uint32_t)n * kTcM * 2u, e + 1};
            tmem_r[n_tr++] = {(uint32_t)col, (uint32_t)(col + n), e + 1};
        }
        p.src[p.nsteps] = {which, block, n0, nlen, k0, klen};
        woff += (uint32_t)align_up((size_t)kp * n * 2, 128);
        ++p.nsteps;
    };
    const uint32_t X0 = p.off_x, X1 = p.off_x + kNpSlotBytes, H0 = p.off_h, H1 = p.off_h + kNpSlotBytes;
    const int T0 = kNpHiddenCol, T1 = kNpHiddenCol + kNpSlotCols;
    // point encoder: 3 -> 128 -> 256 (two N chunks) -> F
    add(X0, H0, 16, 128, T0, 0, 1, 0, boff, 0, 0, 0, 128, 0, 3);                       boff += 128;
    add(H0, X0, 128, 128, T1, 0, 1, 0, boff, 1, 0, 0, 128, 0, 128);
    add(H0, X1, 128, 128, T0, 0, 1, 0, boff + 128, 1, 0, 128, 128, 0, 128);           boff += 256;
    add(X0, X0, 256, F, 0, 0, 2, 0, 0, 2, 0, 0, F, 0, 256);                            // x = pe2(.) ; operand read adds shift_0
    // residual blocks: hidden chunks of <= 128 columns, the issuer one chunk ahead of the epilogues
    int cw[8], co[8], nc = 0;
    for (int o = 0; o < 2 * F; o += kNpSlotCols) { co[nc] = o; cw[nc] = 2 * F - o < kNpSlotCols ? 2 * F - o : kNpSlotCols; ++nc; }
    for (int i = 0; i < nblocks; ++i) {
        auto w1 = [&](int c) {   // relu(W1[co:co+cw, :] x + b1[co:co+cw]) -> hidden slot c & 1
            add(X0, (c & 1) ? H1 : H0, F, cw[c], (c & 1) ? T1 : T0, 0, 1, 0, boff + co[c], 3, i, co[c], cw[c], 0, F);
        };
        auto w2 = [&](int c) {   // x += W2[:, co:co+cw] h_c ; the last chunk's epilogue re-reads x as the next operand
            add((c & 1) ? H1 : H0, X0, cw[c], F, 0, 1, c == nc - 1 ? 2 : 0, i + 1, 0, 4, i, 0, F, co[c], cw[c]);
        };
        w1(0);
        for (int c = 1; c < nc; ++c) { w1(c); w2(c - 1); }
        w2(nc - 1);
        boff += 2 * F;
    }
    // output MLP: F -> 256 (two N chunks) -> 128 -> 3
    add(X0, H0, F, 128, T0, 0, 1, 0, boff, 5, 0, 0, 128, 0, F);
    add(X0, H1, F, 128, T1, 0, 1, 0, boff + 128, 5, 0, 128, 128, 0, F);               boff += 256;
    add(H0, X0, 256, 128, T0, 0, 1, 0, boff, 6, 0, 0, 128, 0, 256);                   boff += 128;
    add(X0, 0, 128, 16, T1, 0, 3, 0, boff, 7, 0, 0, 16, 0, 128);                      boff += 16;
    p.bias_floats = boff;
    {
        const uint32_t vec_bytes = (uint32_t)align_up((size_t)(boff + (nblocks + 1) * F) * sizeof(float), 128);
        const uint32_t limit = 227u * 1024u - 8u * 1024u;   // static shared memory (step table, barriers, stamps) + slack
        const bool tuned = stages_wanted >= 2 && stages_wanted <= kNpMaxStages;
        int stages = tuned ? stages_wanted : kNpMaxStages;
        while (stages > 2 && p.off_ring + (uint32_t)stages * kNpStageBytes + (tuned ? 0u : vec_bytes) > limit) --stages;
        p.nstages = stages;
        p.off_vec = p.off_ring + (uint32_t)stages * kNpStageBytes;
        p.vec_in_smem = p.off_vec + vec_bytes <= limit;
        p.smem_bytes = p.vec_in_smem ? p.off_vec + vec_bytes : p.off_vec;
    }
    if (p.nsteps > kNpMaxSteps) return p;
    p.pack_launches = p.nsteps + 5 + nblocks;
    uint32_t off = woff;
    auto take = [&](size_t floats) { const uint32_t o = off; off += (uint32_t)align_up(floats * sizeof(float), 256); return o; };
    p.off_bias = take(boff);
    p.off_time_w = take((size_t)F * T);
    p.off_time_b = take(F);
    p.off_style_w = take((size_t)F * F);
    p.off_style_b = take(F);
    p.off_pe2_b = take(F);
    p.off_b2 = take((size_t)(nblocks > 0 ? nblocks : 1) * F);
    p.blob_bytes = off;
    p.ok = true;
    return p;
}

}  // namespace pcst

using namespace pcst;

extern "C" size_t pcst_noise_predictor_packed_bytes(int feature_dim, int time_dim, int nblocks) {
    const NpPlan p = np_plan(feature_dim, time_dim, nblocks);
    return p.ok ? p.blob_bytes : 0;
}
extern "C" int pcst_noise_predictor_pack_launches(int feature_dim, int time_dim, int nblocks) {
    const NpPlan p = np_plan(feature_dim, time_dim, nblocks);
    return p.ok ? p.pack_launches : 0;
}
// Host-side self-check of the step table (no GPU): the schedule's invariants for ANY supported (feature_dim, blocks), also the
// ones no GPU test runs.  0 = consistent; otherwise the number of the first violated rule:
//   1 a step waits for an event that no earlier step produces (deadlock)        2 wait counts decrease
//   3 two events on the same barrier could both complete before the issuer has seen the first (parity ambiguity)
//   4 a step may read its operand before the epilogue that writes it has finished (RAW)
//   5 a step may overwrite accumulator columns an earlier epilogue still reads (WAR)
//   6 an operand / accumulator range leaves its buffer                           7 the chain does not end in the output step
extern "C" int pcst_noise_predictor_plan_selfcheck(int feature_dim, int time_dim, int nblocks) {
    const NpPlan p = np_plan(feature_dim, time_dim, nblocks);
    if (!p.ok) return -1;
    int produced = 1;   // event 0 = the input operand
    int prev_wait = 0;
    struct W { uint32_t lo, hi; int ev; };
    W writes[kNpMaxSteps + 1]; int nw = 0;
    W reads[kNpMaxSteps + 1]; int nr = 0;
    writes[nw++] = {p.off_x, p.off_x + 2u * kTcM * 16u, 0};
    for (int s = 0; s < p.nsteps; ++s) {
        const NpStep& st = p.st[s];
        if ((int)st.wait_ev > produced) return 1;
        if ((int)st.wait_ev < prev_wait) return 2;
        prev_wait = st.wait_ev;
        const uint32_t alo = st.a_off, ahi = st.a_off + (uint32_t)st.kp * kTcM * 2u;
        if (ahi > p.off_ring || (uint32_t)st.tmem_col + st.n > 512u) return 6;
        for (int i = 0; i < nw; ++i)
            if (writes[i].lo < ahi && alo < writes[i].hi && writes[i].ev >= (int)st.wait_ev) return 4;
        for (int i = 0; i < nr; ++i)
            if (reads[i].lo < (uint32_t)st.tmem_col + st.n && (uint32_t)st.tmem_col < reads[i].hi && reads[i].ev >= (int)st.wait_ev) return 5;
        if (st.epi == 1 || st.epi == 2) {
            const int e = produced++;
            if (e >= 2 && (int)st.wait_ev < e - 1) return 3;
            if (st.out_off + (uint32_t)st.n * kTcM * 2u > p.off_ring) return 6;
            writes[nw++] = {st.out_off, st.out_off + (uint32_t)st.n * kTcM * 2u, e};
            reads[nr++] = {st.tmem_col, (uint32_t)st.tmem_col + st.n, e};
        }
    }
    if (p.nsteps == 0 || p.st[p.nsteps - 1].epi != 3) return 7;
    return 0;
}
extern "C" size_t pcst_noise_predictor_workspace_bytes(int B, int feature_dim, int nblocks) {
    if (B <= 0 || feature_dim <= 0 || nblocks < 0) return 0;
    return align_up((size_t)B * (nblocks + 1) * feature_dim * sizeof(float), 256);
}

extern "C" int pcst_noise_predictor_pack_f32(const pcst_noise_mlp_t* m, void* packed, size_t packed_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(m && packed, "null pointer");
    const NpPlan p = np_plan(m->feature_dim, m->time_dim, m->nblocks);
    if (!p.ok) {
        set_error("pcst_noise_predictor_pack_f32: unsupported sizes (feature_dim a multiple of 16 in [16, 256], nblocks <= 8)");
        return PCST_ERR_UNSUPPORTED;
    }
    PCST_CHECK_ARG(packed_bytes >= p.blob_bytes && ((uintptr_t)packed & 255) == 0, "packed buffer too small or misaligned");
    const int F = p.F;
    char* blob = (char*)packed;
    for (int s = 0; s < p.nsteps; ++s) {
        const auto& src = p.src[s];
        const float* w = nullptr;
        int rows = 0, cin = 0;
        switch (src.which) {
            case 0: w = m->pe_w[0]; rows = 128; cin = 3; break;
            case 1: w = m->pe_w[1]; rows = 256; cin = 128; break;
            case 2: w = m->pe_w[2]; rows = F; cin = 256; break;
            case 3: w = m->blk_w1[src.block]; rows = 2 * F; cin = F; break;
            case 4: w = m->blk_w2[src.block]; rows = F; cin = 2 * F; break;
            case 5: w = m->out_w[0]; rows = 256; cin = F; break;
            case 6: w = m->out_w[1]; rows = 128; cin = 256; break;
            default: w = m->out_w[2]; rows = 3; cin = 128; break;
        }
        PCST_CHECK_ARG(w, "null weight pointer");
        const int total = p.st[s].kp * p.st[s].n;
        noise_pack_block_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, rows, cin, src.n0, src.nlen, src.k0, src.klen, p.st[s].kp,
                                                                         (__nv_bfloat16*)(blob + p.st[s].w_off));
        PCST_CUDA(cudaGetLastError());
    }
    // bias table in step order
    float* bias = (float*)(blob + p.off_bias);
    uint32_t at = 0;
    auto put = [&](const float* src, int n, int pad) -> int {
        if (!src) {
            set_error("pcst_noise_predictor_pack_f32: null bias pointer");
            return PCST_ERR_INVALID;
        }
        noise_copy_kernel<<<(pad + 255) / 256, 256, 0, stream>>>(src, n, bias + at, pad);
        at += pad;
        return check_cuda(cudaGetLastError(), "noise_copy_kernel");
    };
    int st;
    if ((st = put(m->pe_b[0], 128, 128)) != PCST_OK) return st;
    if ((st = put(m->pe_b[1], 256, 256)) != PCST_OK) return st;
    for (int i = 0; i < p.nblocks; ++i)
        if ((st = put(m->blk_b1[i], 2 * F, 2 * F)) != PCST_OK) return st;
    if ((st = put(m->out_b[0], 256, 256)) != PCST_OK) return st;
    if ((st = put(m->out_b[1], 128, 128)) != PCST_OK) return st;
    if ((st = put(m->out_b[2], 3, 16)) != PCST_OK) return st;
    auto copyf = [&](const float* src, size_t n, uint32_t off) -> int {
        if (!src) {
            set_error("pcst_noise_predictor_pack_f32: null pointer");
            return PCST_ERR_INVALID;
        }
        return check_cuda(cudaMemcpyAsync(blob + off, src, n * sizeof(float), cudaMemcpyDeviceToDevice, stream), "cudaMemcpyAsync");
    };
    if ((st = copyf(m->time_w, (size_t)F * p.T, p.off_time_w)) != PCST_OK) return st;
    if ((st = copyf(m->time_b, F, p.off_time_b)) != PCST_OK) return st;
    if ((st = copyf(m->style_w, (size_t)F * F, p.off_style_w)) != PCST_OK) return st;
    if ((st = copyf(m->style_b, F, p.off_style_b)) != PCST_OK) return st;
    if ((st = copyf(m->pe_b[2], F, p.off_pe2_b)) != PCST_OK) return st;
    for (int i = 0; i < p.nblocks; ++i)
        if ((st = copyf(m->blk_b2[i], F, p.off_b2 + (uint32_t)(i * F * sizeof(float)))) != PCST_OK) return st;
    return PCST_OK;
}

extern "C" int pcst_noise_predictor_f32(const float* points, const int64_t* timestep, const float* style, int B, int N,
                                        int feature_dim, int time_dim, int nblocks, const void* packed, float* out, void* ws,
                                        size_t ws_bytes, pcst_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    PCST_CHECK_ARG(points && timestep && style && packed && out && ws, "null pointer");
    PCST_CHECK_ARG(B > 0 && N > 0, "B and N must be positive");
    const NpPlan p = np_plan(feature_dim, time_dim, nblocks, tuning("noise.stages", 0));
    if (!p.ok) {
        set_error("pcst_noise_predictor_f32: unsupported sizes");
        return PCST_ERR_UNSUPPORTED;
    }
    if (ws_bytes < pcst_noise_predictor_workspace_bytes(B, feature_dim, nblocks) || ((uintptr_t)ws & 255) || ((uintptr_t)packed & 255)) {
        set_error("pcst_noise_predictor_f32: workspace too small or misaligned");
        return PCST_ERR_WORKSPACE;
    }
    const char* blob = (const char*)packed;
    const int F = p.F;
    noise_prep_kernel<<<dim3((F + 7) / 8, B), 256, p.T * sizeof(float), stream>>>(
        (const long long*)timestep, style, F, p.T, p.nblocks, (const float*)(blob + p.off_time_w), (const float*)(blob + p.off_time_b),
        (const float*)(blob + p.off_style_w), (const float*)(blob + p.off_style_b), (const float*)(blob + p.off_pe2_b),
        (const float*)(blob + p.off_b2), (float*)ws);
    PCST_CUDA(cudaGetLastError());
    NpArgs a = {};
    a.pts = points; a.out = out;
    a.B = B; a.N = N; a.F = F;
    a.tiles_per_b = (N + kTcM - 1) / kTcM;
    a.blob = (const unsigned char*)blob;
    a.bias = (const float*)(blob + p.off_bias);
    a.shift = (const float*)ws;
    a.nstage = p.nblocks + 1;
    a.nsteps = p.nsteps;
    a.bias_floats = (int)p.bias_floats;
    for (int s = 0; s < p.nsteps; ++s) a.st[s] = p.st[s];
    a.off_ring = p.off_ring;
    a.off_vec = p.off_vec;
    a.nstages = p.nstages;
    a.vec_in_smem = p.vec_in_smem;
    PCST_CUDA(cudaFuncSetAttribute(noise_mlp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    const long tiles = (long)B * a.tiles_per_b;
    PCST_CHECK_ARG(tiles < (1L << 30), "too many rows");
    int C = tuning("noise.cluster", 0);
    if (C != 2 && C != 4) C = 1;   // measured (profiles/r02/noise_cluster_ab.log): the serial MMA -> epilogue chain per tile, not
                                   // the L2, bounds the kernel today, and the cluster's lockstep costs 18 %: opt-in
    a.cluster = (uint32_t)C;
    a.ntiles = (int)tiles;
    a.probe = tuning("noise.probe", 0);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((tiles + C - 1) / C * C));
    cfg.blockDim = dim3(kNpThreads);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PCST_CUDA(cudaLaunchKernelEx(&cfg, noise_mlp_kernel, a));
    return PCST_OK;
}
