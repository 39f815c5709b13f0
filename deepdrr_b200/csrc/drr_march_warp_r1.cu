// The one-ray-per-lane instantiations of march_warp_kernel (see drr_march_warp.cu), as a translation unit of their own so that
// the two halves of the single-volume march compile in parallel.
#define MARCH_WARP_R1_UNIT
#include "drr_march_warp.cu"
