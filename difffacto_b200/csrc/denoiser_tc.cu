// placeholder -- replaced by the tcgen05 path
#include "denoiser.cuh"
namespace dfb200 {
size_t tc_stream_bytes_for(const NetDims&) { return 0; }
int tc_pack_stream(const PackLayout&, void*, cudaStream_t) { return DFB200_OK; }
int denoiser_forward_tc(const PackLayout&, const void*, int, int, const float*, const float*, const float*, const int*,
                        const float*, float*, Workspace&, cudaStream_t) {
  set_error("bf16 tcgen05 path not built");
  return DFB200_ERR_UNSUPPORTED;
}
}  // namespace dfb200
