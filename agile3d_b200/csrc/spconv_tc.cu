// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace ag3d {
bool spconv_tc_supported(int, int) { return false; }
int spconv_tc_launch(const float*, int, int, const int*, int, long long, const float*, int, const float*, const float*,
                     const float*, int, float*, int, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return AG3D_E_INVALID;
}
}  // namespace ag3d
