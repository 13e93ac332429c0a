// C-ABI runtime glue: error string, driver entry point for TMA descriptor encoding, device query.
#include "vx_host.h"
#include <cstdlib>

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

static int g_pdl = -1;   // -1: not read yet
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("VX_PDL");
    g_pdl = e ? (atoi(e) != 0) : 0;
  }
  return g_pdl != 0;
}

}  // namespace vx

// bring-up hook (csrc/vx_bringup.h): flip programmatic dependent launch inside one process (A/B tests)
extern "C" void vx_pdl_set(int on) { vx::g_pdl = on ? 1 : 0; }
extern "C" int vx_pdl_get() { return vx::pdl_enabled() ? 1 : 0; }

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
