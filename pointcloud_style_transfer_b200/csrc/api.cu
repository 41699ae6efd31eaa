// api.cu -- error plumbing, device check, tuning knobs and the point-packing pre-pass.
#include <stdarg.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace pcst {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return PCST_OK;
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return PCST_ERR_CUDA;
}

int num_sms() {
    static thread_local int cached_dev = -2, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();  // no device (plan queries on a build host): the B200 value
        return 148;
    }
    if (dev != cached_dev) {
        int n = 0;
        cached = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
        cached_dev = dev;
    }
    return cached;
}

static std::mutex g_tune_mu;
static std::map<std::string, int>& tune_map() {
    static std::map<std::string, int> m = {
        {"nn_min.variant", 0},        // 0 = auto
        {"nn_min.splits", 0},         // 0 = auto (candidate-range splits per row tile)
        {"fps.cluster", 0},           // 0 = auto (CTAs per cloud)
        {"fps.threads", 0},           // 0 = auto (32, 128 or 512 threads per CTA)
        {"fps.prune", 0},             // 0/1 = bounding-box skip test on, 2 = off
        {"fps.lookahead", 0},         // 0/2 = one sample per exchange (default); 1 = exact two-sample look-ahead kernel; 3, 4 = its probes
        {"knn.grid", 0},              // 0 = auto, 1 = always the exact grid search, 2 = always the brute-force sweep
        {"noise.cluster", 0},         // 0 / 1 = off, 2 / 4 = CTAs per cluster sharing the denoiser's weight stages by TMA multicast
        {"sa_mlp.reuse_h", 0},        // 0/1 = layer 1's output overwrites layer 0's in shared memory (plain one-CTA plans), 2 = off
        {"sa_mlp.early_gather", 0},   // 0/1 = the next row tile is gathered under the current tile's last MMAs (walk build), 2 = off
        {"sa_mlp.bf16_feats", 0},     // 0/1 = bf16 copy of the feature tensor for the gather at throughput shapes, 2 = off
        {"sa_mlp.transpose_pool", 0}, // 0/1 = pooled layer computed transposed (per-thread max over the rows), 2 = row-major + warp reductions
        {"noise.stages", 0},          // weight-ring stages of the denoiser kernel (0 = as many as fit beside the staged bias vectors: 4)
        {"noise.probe", 0},           // 64 = the denoiser kernel prints clock stamps of its first steps (one CTA)
        {"sa_mlp.variant", 0},
        {"sa_mlp.pdl", 0},            // 0/1 = programmatic dependent launch of the tcgen05 MLP kernel on, 2 = off
        {"sa_mlp.regs", 0},           // 0 = auto, 80 / 168 = register budget variant of the tcgen05 MLP kernel
        {"sa_mlp.backoff", 0},        // 0/1 = producer / MMA threads pause between barrier polls at throughput shapes, 2 = off
        {"sa_mlp.persistent", 0},     // 0/1 = CTAs walk several row tiles when tiles exceed the machine, 2 = one CTA per tile
    };
    return m;
}

int tuning(const char* key, int dflt) {
    std::lock_guard<std::mutex> lk(g_tune_mu);
    auto it = tune_map().find(key);
    if (it == tune_map().end() || it->second == 0) return dflt;
    return it->second;
}

// ---- pack: [B,N,3] fp32 -> [B,Npad] float4 (x, y, z, |p|^2), sentinel-padded -----------------
// HBM-bound, 12 B read + 16 B written per point; ~2 MB per 120k-point scan.
__global__ void pack_points_kernel(const float* __restrict__ xyz, int N, int Npad, float4* __restrict__ out) {
    const int b = blockIdx.y;
    const float* src = xyz + (size_t)b * N * 3;
    float4* dst = out + (size_t)b * Npad;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Npad; i += gridDim.x * blockDim.x) {
        float4 v;
        if (i < N) {
            v.x = src[3 * i];
            v.y = src[3 * i + 1];
            v.z = src[3 * i + 2];
            v.w = norm3_sq(v.x, v.y, v.z);
        } else {
            v = make_float4(0.f, 0.f, 0.f, __int_as_float(0x7f800000));
        }
        dst[i] = v;
    }
}

// ---- L2 prefetch: one prefetch.global.L2 per 128-byte line --------------------------------------
__global__ void l2_prefetch_kernel(const unsigned char* __restrict__ p, size_t bytes) {
    const size_t line = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    if (line < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + line));
}

// ---- FP32 pipe calibration: independent packed FMA chains, 2 * 2 * 8 * iters flop per thread -----------------
__global__ void __launch_bounds__(256)
fp32_probe_kernel(int iters, float* __restrict__ out) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(1.0f + threadIdx.x * 1e-6f + i, 0.5f + i);
    const float2 m = make_float2(0.999999f, 1.000001f), c = make_float2(1e-7f, -1e-7f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = __ffma2_rn(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}

int launch_pack(const float* xyz, int B, int N, int Npad, float4* out, cudaStream_t stream) {
    int blocks = (Npad + 255) / 256;
    if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
    pack_points_kernel<<<dim3(blocks, B), 256, 0, stream>>>(xyz, N, Npad, out);
    return check_cuda(cudaGetLastError(), "pack_points_kernel");
}

}  // namespace pcst

extern "C" {

const char* pcst_version(void) { return "pcst 0.1.0 (sm_100a)"; }

const char* pcst_last_error(void) { return pcst::g_err; }

int pcst_device_check(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        pcst::set_error("pcst_device_check: no CUDA device");
        return PCST_ERR_UNSUPPORTED;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        pcst::set_error("pcst_device_check: compute capability %d.x, libpcst needs 10.x (sm_100a)", major);
        return PCST_ERR_UNSUPPORTED;
    }
    return PCST_OK;
}

int pcst_l2_prefetch(const void* ptr, size_t bytes, pcst_stream_t stream_) {
    if (!ptr || bytes == 0) return PCST_OK;
    const size_t lines = (bytes + 127) / 128;
    const unsigned blocks = (unsigned)((lines + 255) / 256);
    pcst::l2_prefetch_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>((const unsigned char*)ptr, bytes);
    return pcst::check_cuda(cudaGetLastError(), "l2_prefetch_kernel");
}

long long pcst_fp32_probe(int iters, float* scratch, pcst_stream_t stream_) {
    if (iters <= 0 || !scratch) return 0;
    const int blocks = 8 * pcst::num_sms(), threads = 256;
    pcst::fp32_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream_>>>(iters, scratch);
    if (cudaGetLastError() != cudaSuccess) return 0;
    return (long long)blocks * threads * 8 * 2 * 2 * iters;  // 8 chains x 2 lanes x (mul + add) per iteration
}

int pcst_set_tuning(const char* key, int value) {
    if (!key) return PCST_ERR_INVALID;
    std::lock_guard<std::mutex> lk(pcst::g_tune_mu);
    auto it = pcst::tune_map().find(key);
    if (it == pcst::tune_map().end()) {
        pcst::set_error("pcst_set_tuning: unknown key '%s'", key);
        return PCST_ERR_INVALID;
    }
    it->second = value;
    return PCST_OK;
}

int pcst_get_tuning(const char* key, int* value) {
    if (!key || !value) return PCST_ERR_INVALID;
    std::lock_guard<std::mutex> lk(pcst::g_tune_mu);
    auto it = pcst::tune_map().find(key);
    if (it == pcst::tune_map().end()) {
        pcst::set_error("pcst_get_tuning: unknown key '%s'", key);
        return PCST_ERR_INVALID;
    }
    *value = it->second;
    return PCST_OK;
}

}  // extern "C"
