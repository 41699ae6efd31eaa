// sa_mlp_tc.cu -- SetAbstraction shared MLP + max-pool on the tcgen05 / TMEM tensor-core path.
// (placeholder: the bf16 tcgen05 chain is wired in a later step; precision == 1 is refused loudly
// rather than silently falling back to the fp32 path)
#include "common.cuh"

namespace pcst {

size_t sa_mlp_max_tc_workspace(int, int, int, int, int, const pcst_mlp3_t*) { return 256; }

int sa_mlp_max_tc(const float*, const float*, const float*, const int64_t*, int, int, int, int, int,
                  const pcst_mlp3_t*, float*, void*, size_t, cudaStream_t) {
    set_error("pcst_sa_mlp_max_f32: precision 1 (tcgen05 bf16) is not built in this version");
    return PCST_ERR_UNSUPPORTED;
}

}  // namespace pcst
