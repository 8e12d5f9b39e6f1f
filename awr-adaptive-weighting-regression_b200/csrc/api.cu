#include "common.cuh"
#include "awr_b200.h"
extern "C" int awr_version(void) { return AWR_B200_VERSION; }
