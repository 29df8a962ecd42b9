// placeholder until the tcgen05 path lands
#include "dist_common.cuh"
int hg_launch_dist_tc(hg_ctx *, const int16_t *, const int32_t *, uint32_t, uint32_t, const int16_t *, const int32_t *,
                      uint32_t, uint32_t, uint32_t, uint32_t, float, int, hg_hit *, uint64_t, unsigned long long *) {
  hg_set_error("tensor path not built");
  return HG_E_UNSUPPORTED;
}
