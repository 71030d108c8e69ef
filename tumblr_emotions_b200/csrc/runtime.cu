// Runtime glue of libdeepsent: error channel, device binding, driver entry points for TMA maps.
#include <cuda.h>
#include <stdarg.h>
#include "common.cuh"
#include "tmap.cuh"
#ifdef DS_DEV
#include "../../include/deepsent_dev.h"
#endif

namespace ds {

std::string& last_error() {
  static thread_local std::string e;
  return e;
}

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error() = buf;
  return 1;
}

static int g_sm_count = 0;
int g_debug[16] = {0};
int g_pdl = 0;
PFN_encodeTiled g_encode_tiled = nullptr;
PFN_encodeIm2col g_encode_im2col = nullptr;

}  // namespace ds

extern "C" {

int ds_version(void) { return 200; }

const char* ds_last_error(void) { return ds::last_error().c_str(); }

int ds_init(int device) {
  DS_CUDA(cudaSetDevice(device));
  DS_CUDA(cudaFree(0));
  cudaDeviceProp prop;
  DS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return ds::fail("libdeepsent is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  ds::g_sm_count = prop.multiProcessorCount;
  cudaDriverEntryPointQueryResult q;
  void* fn = nullptr;
  DS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) return ds::fail("cuTensorMapEncodeTiled not available");
  ds::g_encode_tiled = reinterpret_cast<ds::PFN_encodeTiled>(fn);
  fn = nullptr;
  DS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !fn) return ds::fail("cuTensorMapEncodeIm2col not available");
  ds::g_encode_im2col = reinterpret_cast<ds::PFN_encodeIm2col>(fn);
  return 0;
}

int ds_sm_count(void) { return ds::g_sm_count; }

int ds_launch_count(void) { return ds::g_debug[15]; }

int ds_dependent_launch(int mode) {
  if (mode < 0 || mode > 7) return ds::fail("ds_dependent_launch: mode %d is not a combination of DS_PDL_* bits", mode);
  ds::g_pdl = mode;
  return 0;
}

#ifdef DS_DEV
// development build only (libdeepsent_dev.so, include/deepsent_dev.h): launch-policy overrides for the tuning tools and the
// per-kernel tests that force a mode.  The product library has no setter, so its policy table stays all-default.
int ds_debug_set(int key, int value) {
  if (key < 0 || key >= 15) return ds::fail("ds_debug_set: bad key %d", key);
  ds::g_debug[key] = value;
  return 0;
}

int ds_debug_get(int key) { return (key < 0 || key >= 16) ? 0 : ds::g_debug[key]; }
#endif

}  // extern "C"
