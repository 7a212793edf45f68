// Library-level helpers of libncn.so: version, error strings, device info.
#include "ncn_common.cuh"
#include <atomic>

namespace ncn {

int sm_count() {
  // cached per device (a process normally drives one GPU; 16 covers a full box)
  static std::atomic<int> cache[16];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < 16) {
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  if (dev >= 0 && dev < 16) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

}  // namespace ncn

extern "C" int ncn_version(void) { return NCN_VERSION; }

extern "C" const char* ncn_error_string(int code) {
  switch (code) {
    case NCN_OK: return "ok";
    case NCN_E_NULL: return "ncn: a required pointer is NULL";
    case NCN_E_SIZE: return "ncn: a size / count argument is out of range";
    case NCN_E_CONFIG: return "ncn: unsupported configuration";
    case NCN_E_ALIGN: return "ncn: pointer alignment requirement violated";
    case NCN_E_NCCL: return "ncn: NCCL failure";
    case NCN_E_UNSUPPORTED: return "ncn: unsupported";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "ncn: unknown error";
}

extern "C" int ncn_device_info(int* sm, int* cc_major, int* cc_minor) {
  int dev = 0;
  NCN_CUDA(cudaGetDevice(&dev));
  if (sm) NCN_CUDA(cudaDeviceGetAttribute(sm, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) NCN_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) NCN_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return NCN_OK;
}

// developer aid: writes the GPU global timer (ns) into slots[slot] when the stream reaches this point - a timeline of a
// captured CUDA graph, where CUDA events cannot be read (tools/timeline.py)
namespace ncn {
__global__ void stamp_kernel(unsigned long long* slots, int slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  slots[slot] = t;
}
}  // namespace ncn
extern "C" int ncn_debug_stamp(uint64_t* slots, int slot, ncn_stream_t stream) {
  NCN_CHECK_PTR(slots);
  ncn::stamp_kernel<<<1, 1, 0, ncn::as_stream(stream)>>>((unsigned long long*)slots, slot);
  NCN_LAUNCH_OK();
  return NCN_OK;
}
