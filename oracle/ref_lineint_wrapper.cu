// TEST INFRASTRUCTURE -- not part of the product.
//
// Exposes the reference kernel's per-material line integrals (area_density[m], project_kernel.cu:
// 565-584) without touching its source: the reference file is #included from where it lies under
// /root/reference (oracle/Makefile passes the -I), with `expf` macro-replaced by the identity.  Run
// with n_bins = 1, pdf = {-1}, mu = one-hot(m), the kernel's Beer-Lambert tail (project_kernel.cu:
// 637-646) then stores photon_prob = area_density[m] exactly; the ray march is the reference's own
// compiled code.
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>
#define expf(x) (x)
#include <project_kernel.cu>
