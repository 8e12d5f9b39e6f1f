#include "common.cuh"
#include "awr_b200.h"
extern "C" int awr_version(void) { return AWR_B200_VERSION; }
// 1: this build accumulates shared sums order-independently (bit-reproducible training steps; `make DET=1` -> libawr_b200_det.so)
extern "C" int awr_deterministic(void) {
#ifdef AWR_DETERMINISTIC
  return 1;
#else
  return 0;
#endif
}

// SMs the persistent tensor-core kernels may occupy (grid cap).  The data-parallel trainer lowers it while a gradient bucket's NCCL
// all-reduce is in flight, so the collective's CTAs find free SMs instead of queueing behind 148-CTA persistent grids.
static int g_sm_budget = 148;
int awr_sm_budget() { return g_sm_budget; }
extern "C" int awr_set_sm_budget(int n) {
  const int prev = g_sm_budget;
  if (n >= 8 && n <= 148) g_sm_budget = n;
  return prev;
}
