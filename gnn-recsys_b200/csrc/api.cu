// Library-level entry points: error string, version, device probe.
#include <atomic>

#include "common.cuh"

namespace gr {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace gr

extern "C" const char* gr_last_error(void) { return gr::g_last_error.c_str(); }

extern "C" int gr_version(void) { return 100; }

extern "C" long long gr_launch_count(void) { return gr::g_launches.load(std::memory_order_relaxed); }

extern "C" int gr_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0, sms = 0, major = 0, minor = 0;
  GR_CUDA(cudaGetDevice(&dev));
  GR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  GR_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count_host) *sm_count_host = sms;
  if (cc_major_host) *cc_major_host = major;
  if (cc_minor_host) *cc_minor_host = minor;
  GR_REQUIRE(major == 10, GR_E_UNSUPPORTED, "this library is built for sm_100a (B200) only");
  return GR_OK;
}
