// C-ABI runtime glue: error string, driver entry point for TMA descriptor encoding, device query.
#include "vx_host.h"

namespace vx {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

}  // namespace vx

extern "C" const char* vx_last_error() { return vx::last_error_buf(); }

extern "C" int vx_abi_version() { return 1; }

// Fails loudly unless the current device is an sm_100 part: there is no fallback path.
extern "C" int vx_require_sm100() {
  int dev = 0;
  cudaDeviceProp prop;
  VX_CHECK_CUDA(cudaGetDevice(&dev));
  VX_CHECK_CUDA(cudaGetDeviceProperties(&prop, dev));
  VX_REQUIRE(prop.major == 10, "vxb200 needs an sm_100a GPU (B200); found sm_%d%d (%s)", prop.major, prop.minor,
             prop.name);
  return 0;
}
