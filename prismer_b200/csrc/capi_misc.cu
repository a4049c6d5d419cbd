// Library-level entry points.
#include "common.cuh"
#include "prismer_sm100.h"

extern "C" int prismer_abi_version(void) { return PRISMER_ABI_VERSION; }

extern "C" int prismer_check_device(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return PRISMER_ERR_CUDA;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return PRISMER_ERR_CUDA;
  return major == 10 ? PRISMER_OK : PRISMER_ERR_ARCH;
}
