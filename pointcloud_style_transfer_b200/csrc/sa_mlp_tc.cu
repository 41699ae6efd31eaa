// sa_mlp_tc.cu -- SetAbstraction grouping + shared MLP + max-pool on tcgen05 tensor cores (sm_100a).
//
// Replaces index_points x2 + subtract + cat + apply_mlp (models/pointnet2_encoder.py:94-112, eval
// mode) with ONE kernel per set-abstraction stage: gather -> 3 x (GEMM, scale/shift, ReLU) -> max.
// A CTA owns a tile of 128 rows (rows = (b, s, k) flattened = 4 groups of K=32, 2 of K=64, ...).
//   * layer-0 input rows are gathered from HBM (xyz - centroid, features), converted to bf16 and laid
//     out in shared memory as the K-major, no-swizzle UMMA operand [Kp/8][128 rows][8] (one 8x16-byte
//     core matrix = 128 contiguous bytes; SBO = 128 B between 8-row groups, LBO = 2048 B between
//     K chunks);
//   * the three weight matrices are pre-packed once per call to bf16 [Kp/8][Cout][8] (same canonical
//     layout) and brought in by one 1-D TMA bulk copy each, all resident for the CTA's lifetime;
//   * each layer is Kp/16 tcgen05.mma (M=128, N=Cout, K=16, bf16 x bf16 -> fp32) issued by one thread,
//     accumulating in TMEM (128 lanes x Cout columns); completion is signalled with tcgen05.commit on
//     an mbarrier;
//   * the epilogue reads the accumulator with tcgen05.ld (32 lanes x 16 columns per warp-instruction),
//     applies the folded conv-bias/BatchNorm scale+shift and ReLU in fp32, and either writes the next
//     layer's bf16 operand straight back into shared memory in the canonical layout, or (last layer)
//     reduces the max over each group's rows with REDUX and merges across warps/CTAs with atomicMax on
//     the IEEE bits (post-ReLU values are >= +0).
// Activations never touch HBM.  Bound: tensor pipe in the limit of many rows; at the reference's
// shapes (16 384 / 8 192 rows per scan) the stage is latency-bound.
// Precision: bf16 operands, fp32 accumulate/epilogue -> features within rtol 2e-2 / atol 2e-2 of
// the reference's fp32 result.
#include <cuda_bf16.h>

#include "common.cuh"

namespace pcst {

constexpr int kTcM = 128;        // rows per CTA = TMEM lanes
constexpr int kTcThreads = 128;  // 4 warps: warp w owns TMEM lanes [32w, 32w+32)
constexpr int kTcMaxN = 256;     // one tcgen05.mma covers the whole layer width

struct TcLayer {
    const __nv_bfloat16* w;  // packed [kp/8][n][8]
    const float* scale;
    const float* shift;
    int kp;  // padded reduction length (multiple of 16)
    int n;   // output channels (multiple of 32, <= 256)
    uint32_t smem_off;  // byte offset of the weights inside dynamic shared memory
};

struct TcArgs {
    const float* xyz;
    const float* feats;
    const float* new_xyz;
    const int64_t* idx;
    int N, S, K, D, rows;
    TcLayer L[3];
    uint32_t off_a, off_b, off_ss;  // activation ping-pong buffers, scale/shift staging
    uint32_t tmem_cols;
    unsigned int* out_bits;  // [B, Cout, S]
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // leading (K) byte offset between core matrices
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride (M/N) byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell)
    return d;                                            // base_offset 0, layout_type 0 = no swizzle
}
__device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
    return (1u << 4)                      // D format: fp32
           | (1u << 7)                    // A format: bf16
           | (1u << 10)                   // B format: bf16
           | ((uint32_t)(N >> 3) << 17)   // N
           | ((uint32_t)(M >> 4) << 24);  // M;  A and B both K-major (bits 15, 16 = 0)
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- weight pre-pack: fp32 [Cout, Cin] -> bf16 [kp/8][Cout][8], zero padded along K ------------
__global__ void tc_pack_weights_kernel(const float* __restrict__ w, int cout, int cin, int kp,
                                       __nv_bfloat16* __restrict__ out) {
    const int total = kp * cout;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kc = e / (cout * 8);        // K chunk
        const int rem = e % (cout * 8);
        const int n = rem / 8, ke = rem % 8;
        const int k = kc * 8 + ke;
        out[e] = __float2bfloat16_rn(k < cin ? w[(size_t)n * cin + k] : 0.f);
    }
}

__global__ void __launch_bounds__(kTcThreads)
sa_mlp_tc_kernel(const TcArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t wbar[3];
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = blockIdx.x * kTcM;

    if (tid == 0) {
        for (int l = 0; l < 3; ++l) mbar_init(&wbar[l], 1);
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)),
                     "r"(a.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_sh;

    // ---- weights: one TMA bulk copy per layer, all issued up front ----
    if (tid == 0) {
        for (int l = 0; l < 3; ++l) {
            const uint32_t bytes = (uint32_t)a.L[l].kp * a.L[l].n * 2u;
            mbar_arrive_expect_tx(&wbar[l], bytes);
            tma_load_1d(smem + a.L[l].smem_off, a.L[l].w, bytes, &wbar[l]);
        }
    }
    // ---- scale / shift of the three layers -> shared memory ----
    float* ss = reinterpret_cast<float*>(smem + a.off_ss);  // [3][2][kTcMaxN]
    for (int l = 0; l < 3; ++l)
        for (int c = tid; c < a.L[l].n; c += kTcThreads) {
            ss[(l * 2 + 0) * kTcMaxN + c] = a.L[l].scale[c];
            ss[(l * 2 + 1) * kTcMaxN + c] = a.L[l].shift[c];
        }

    // ---- layer-0 operand: warp-cooperative gather, fp32 -> bf16, canonical [kp/8][128][8] layout ----
    {
        __nv_bfloat16* A0 = reinterpret_cast<__nv_bfloat16*>(smem + a.off_a);
        const int kp0 = a.L[0].kp;
        const int cin = 3 + a.D;
        const int my_row = row0 + tid;
        int my_j = 0, my_bs = 0;
        if (my_row < a.rows) {
            my_bs = my_row / a.K;
            if (a.idx) {
                const int64_t jj = a.idx[my_row];
                my_j = jj < 0 ? 0 : (jj >= a.N ? a.N - 1 : (int)jj);
            } else {
                my_j = my_row % a.K;  // group_all: row k of cloud b is point k
            }
        }
        for (int r = 0; r < 32; ++r) {
            const int m = warp * 32 + r;
            const int row = row0 + m;
            const int j = __shfl_sync(0xffffffffu, my_j, r);
            const int bs = __shfl_sync(0xffffffffu, my_bs, r);
            const int b = bs / a.S;
            const bool valid = row < a.rows;
            const float* prow = a.xyz + ((size_t)b * a.N + j) * 3;
            const float* crow = a.new_xyz ? a.new_xyz + (size_t)bs * 3 : nullptr;
            const float* frow = a.feats ? a.feats + ((size_t)b * a.N + j) * a.D : nullptr;
            for (int c = lane; c < kp0; c += 32) {
                float v = 0.f;
                if (valid && c < cin) {
                    if (c < 3) {
                        v = prow[c];
                        if (crow) v = __fsub_rn(v, crow[c]);
                    } else {
                        v = __ldg(frow + (c - 3));
                    }
                }
                A0[((size_t)(c >> 3) * kTcM + m) * 8 + (c & 7)] = __float2bfloat16_rn(v);
            }
        }
    }
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const TcLayer& L = a.L[l];
        const uint32_t in_off = (l == 1) ? a.off_b : a.off_a;   // L0: A -> B, L1: B -> A, L2: A -> pooled
        const uint32_t out_off = (l == 0) ? a.off_b : a.off_a;
        if (tid == 0) {
            mbar_wait(&wbar[l], 0);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + in_off);
            const uint32_t w_addr = smem_u32(smem + L.smem_off);
            const uint32_t lbo_a = kTcM * 16, lbo_w = (uint32_t)L.n * 16;
            const uint32_t idesc = umma_idesc_bf16(kTcM, L.n);
            for (int kk = 0; kk < L.kp / 16; ++kk) {
                const uint64_t ad = umma_smem_desc(a_addr + (uint32_t)kk * 2u * lbo_a, lbo_a, 128);
                const uint64_t bd = umma_smem_desc(w_addr + (uint32_t)kk * 2u * lbo_w, lbo_w, 128);
                umma_bf16(tmem_base, ad, bd, idesc, kk > 0);
            }
            umma_commit(&mma_bar);  // arrives when every MMA above has completed
        }
        mbar_wait(&mma_bar, (uint32_t)(l & 1));
        tc_fence_after();

        const float* sc = ss + (l * 2 + 0) * kTcMaxN;
        const float* sh = ss + (l * 2 + 1) * kTcMaxN;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int m = tid;  // this thread's row = its TMEM lane
        if (l < 2) {
            unsigned char* outp = smem + out_off;
            for (int c0 = 0; c0 < L.n; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                uint32_t packed[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float y0 = __fmaf_rn(__uint_as_float(r[2 * i]), sc[c0 + 2 * i], sh[c0 + 2 * i]);
                    float y1 = __fmaf_rn(__uint_as_float(r[2 * i + 1]), sc[c0 + 2 * i + 1], sh[c0 + 2 * i + 1]);
                    y0 = y0 > 0.f ? y0 : 0.f;
                    y1 = y1 > 0.f ? y1 : 0.f;
                    __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                    packed[i] = *reinterpret_cast<uint32_t*>(&h);
                }
                // channels c0..c0+7 and c0+8..c0+15 are two K chunks of the next operand
                uint4* d0 = reinterpret_cast<uint4*>(outp + ((size_t)(c0 >> 3) * kTcM + m) * 16);
                uint4* d1 = reinterpret_cast<uint4*>(outp + ((size_t)((c0 >> 3) + 1) * kTcM + m) * 16);
                *d0 = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *d1 = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            tc_fence_after();
        } else {
            // ---- last layer: max over each group's K rows, merged with atomicMax on the fp32 bits ----
            const int row = row0 + m;
            const bool valid = row < a.rows;
            const bool warp_uniform_group = (a.K % 32) == 0;  // the warp's 32 rows belong to one group
            const int wrow = row0 + warp * 32;
            const int g = (warp_uniform_group ? wrow : (valid ? row : 0)) / a.K;
            const int b = g / a.S, s = g % a.S;
            unsigned int* obase = a.out_bits + ((size_t)b * L.n) * a.S + s;
            for (int c0 = 0; c0 < L.n; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(taddr + (uint32_t)c0, r);
                unsigned int mine = 0;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float y = __fmaf_rn(__uint_as_float(r[i]), sc[c0 + i], sh[c0 + i]);
                    y = (valid && y > 0.f) ? y : 0.f;
                    if (warp_uniform_group) {
                        const unsigned int mx = __reduce_max_sync(0xffffffffu, __float_as_uint(y));
                        if (lane == i) mine = mx;
                    } else if (valid) {
                        atomicMax(obase + (size_t)(c0 + i) * a.S, __float_as_uint(y));
                    }
                }
                if (warp_uniform_group && lane < 16 && wrow < a.rows)
                    atomicMax(obase + (size_t)(c0 + lane) * a.S, mine);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols)
                     : "memory");
    }
}

struct TcPlan {
    int kp[3], n[3];
    uint32_t off_w[3], off_a, off_b, off_ss, smem_bytes, tmem_cols;
    size_t ws_w[3], ws_total;
    bool ok;
};

static TcPlan tc_plan(int D, const pcst_mlp3_t* mlp) {
    TcPlan p = {};
    int cin = 3 + D;
    uint32_t off = 0;
    size_t ws = 0;
    int maxn = 0;
    p.ok = true;
    for (int l = 0; l < 3; ++l) {
        p.kp[l] = (int)align_up((size_t)cin, 16);
        p.n[l] = mlp->cout[l];
        if (p.n[l] > kTcMaxN || (p.n[l] % 32) != 0) p.ok = false;
        p.off_w[l] = off;
        off += (uint32_t)align_up((size_t)p.kp[l] * p.n[l] * 2, 128);
        p.ws_w[l] = ws;
        ws += align_up((size_t)p.kp[l] * p.n[l] * 2, 256);
        if (p.n[l] > maxn) maxn = p.n[l];
        cin = p.n[l];
    }
    const int a_k = p.kp[0] > p.n[1] ? p.kp[0] : p.n[1];  // buffer A: layer-0 input, later layer-1 output
    p.off_a = off;
    off += (uint32_t)kTcM * a_k * 2;
    p.off_b = off;
    off += (uint32_t)kTcM * p.n[0] * 2;
    p.off_ss = off;
    off += 3 * 2 * kTcMaxN * sizeof(float);
    p.smem_bytes = off;
    p.tmem_cols = maxn <= 32 ? 32 : maxn <= 64 ? 64 : maxn <= 128 ? 128 : 256;
    p.ws_total = ws;
    if (p.smem_bytes > 220 * 1024) p.ok = false;
    return p;
}

bool sa_mlp_max_tc_supported(int D, const pcst_mlp3_t* mlp) { return tc_plan(D, mlp).ok; }

size_t sa_mlp_max_tc_workspace(int B, int N, int S, int K, int D, const pcst_mlp3_t* mlp) {
    (void)B; (void)N; (void)S; (void)K;
    return tc_plan(D, mlp).ws_total + 256;
}

int sa_mlp_max_tc(const float* xyz, const float* feats, const float* new_xyz, const int64_t* idx, int B, int N, int S,
                  int K, int D, const pcst_mlp3_t* mlp, float* out, void* ws, size_t ws_bytes, cudaStream_t stream) {
    const TcPlan p = tc_plan(D, mlp);
    if (!p.ok) {
        set_error("sa_mlp_max (tensor-core path): needs Cout <= 256 (multiple of 32) and <= 220 KiB of shared memory");
        return PCST_ERR_UNSUPPORTED;
    }
    if (ws_bytes < p.ws_total) return PCST_ERR_WORKSPACE;
    const size_t rows_sz = (size_t)B * S * K;
    if (rows_sz >= (1u << 30)) {
        set_error("sa_mlp_max: B*S*K too large");
        return PCST_ERR_INVALID;
    }
    TcArgs a = {};
    a.xyz = xyz; a.feats = feats; a.new_xyz = new_xyz; a.idx = idx;
    a.N = N; a.S = S; a.K = K; a.D = D; a.rows = (int)rows_sz;
    int cin = 3 + D;
    for (int l = 0; l < 3; ++l) {
        __nv_bfloat16* wp = (__nv_bfloat16*)((char*)ws + p.ws_w[l]);
        const int total = p.kp[l] * p.n[l];
        tc_pack_weights_kernel<<<(total + 255) / 256, 256, 0, stream>>>(mlp->w[l], p.n[l], cin, p.kp[l], wp);
        PCST_CUDA(cudaGetLastError());
        a.L[l].w = wp; a.L[l].scale = mlp->scale[l]; a.L[l].shift = mlp->shift[l];
        a.L[l].kp = p.kp[l]; a.L[l].n = p.n[l]; a.L[l].smem_off = p.off_w[l];
        cin = p.n[l];
    }
    a.off_a = p.off_a; a.off_b = p.off_b; a.off_ss = p.off_ss; a.tmem_cols = p.tmem_cols;
    a.out_bits = (unsigned int*)out;
    PCST_CUDA(cudaMemsetAsync(out, 0, (size_t)B * p.n[2] * S * sizeof(float), stream));
    PCST_CUDA(cudaFuncSetAttribute(sa_mlp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
    const int grid = (a.rows + kTcM - 1) / kTcM;
    sa_mlp_tc_kernel<<<grid, kTcThreads, p.smem_bytes, stream>>>(a);
    return check_cuda(cudaGetLastError(), "sa_mlp_tc_kernel");
}

}  // namespace pcst
