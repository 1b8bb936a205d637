// Error state and device gate behind the C ABI.
#include <cstdlib>

#include "common.cuh"

namespace mmr {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

mmr_status fail(mmr_status code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

// knob -> {environment variable, default}
static int g_tuning[MMR_TUNE_COUNT];
static bool g_tuning_init = false;
static unsigned g_tuning_generation = 0;   // bumped by mmr_set_tuning: captured CUDA graphs of a forward are keyed on it
static void tuning_init() {
  static const struct { const char* env; int def; } spec[MMR_TUNE_COUNT] = {
      {"MMR_GEMM_PAIR", 1}, {"MMR_GEMM_P16", 1}, {"MMR_GEMM_TAIL", 1}, {"MMR_GEMM_CLUSTER", 1}, {"MMR_GEMM_LN", 1}, {"MMR_PDL", 1}, {"MMR_ATTN_TMA", 0}, {"MMR_ATTN_TC", 2}, {"MMR_LN_ROW_CFG", 0}, {"MMR_LABEL_DEDUP", 1}, {"MMR_LX_MERGE", 1}, {"MMR_PRUNE_LAST", 1}, {"MMR_LX_QUERY_DEDUP", 1}};
  for (int i = 0; i < MMR_TUNE_COUNT; ++i) {
    const char* e = getenv(spec[i].env);
    g_tuning[i] = e ? atoi(e) : spec[i].def;
  }
  g_tuning_init = true;
}
int tuning(int knob) {
  if (!g_tuning_init) tuning_init();
  return (knob >= 0 && knob < MMR_TUNE_COUNT) ? g_tuning[knob] : 0;
}

mmr_status require_sm100() {
  static thread_local int cached_dev = -1;
  static thread_local mmr_status cached = MMR_ERR_ARCH;
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    return fail(MMR_ERR_ARCH, "no CUDA device available (%s); this library has no CPU fallback",
                cudaGetErrorString(e));
  }
  if (dev == cached_dev) {
    if (cached != MMR_OK) fail(cached, "device %d is not sm_100 (B200); no fallback path exists", dev);
    return cached;
  }
  int major = 0, minor = 0;
  MMR_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MMR_CUDA_OK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  cached_dev = dev;
  cached = (major == 10) ? MMR_OK : MMR_ERR_ARCH;
  if (cached != MMR_OK) {
    return fail(MMR_ERR_ARCH, "device %d is sm_%d%d, need sm_100 (B200); no fallback path exists", dev, major,
                minor);
  }
  return MMR_OK;
}

}  // namespace mmr

extern "C" const char* mmr_last_error(void) { return mmr::last_error_buf(); }
extern "C" int mmr_abi_version(void) { return 2; }
extern "C" int mmr_experimental_build(void) {
#ifdef MMR_EXPERIMENTAL
  return 1;
#else
  return 0;
#endif
}
extern "C" mmr_status mmr_set_tuning(int knob, int value) {
  if (knob < 0 || knob >= MMR_TUNE_COUNT) return mmr::fail(MMR_ERR_INVALID, "mmr_set_tuning: unknown knob %d", knob);
  if (!mmr::g_tuning_init) mmr::tuning_init();
  mmr::g_tuning[knob] = value;
  ++mmr::g_tuning_generation;
  return MMR_OK;
}
extern "C" unsigned mmr_tuning_generation(void) { return mmr::g_tuning_generation; }
extern "C" int mmr_get_tuning(int knob) { return (knob >= 0 && knob < MMR_TUNE_COUNT) ? mmr::tuning(knob) : -1; }
extern "C" mmr_status mmr_device_check(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    return mmr::fail(MMR_ERR_ARCH, "CUDA device %d not present; this library has no CPU fallback", device);
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (major != 10) return mmr::fail(MMR_ERR_ARCH, "device %d is sm_%d%d, need sm_100 (B200)", device, major, minor);
  return MMR_OK;
}
