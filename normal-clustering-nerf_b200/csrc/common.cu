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

static std::atomic<int> g_pdl{1};
int pdl_enabled() { return g_pdl.load(std::memory_order_relaxed); }

}  // namespace ncn

// programmatic dependent launch between the kernels of the training step's serial chain (1 = on, default); returns the old value
extern "C" int ncn_set_pdl(int on) { return ncn::g_pdl.exchange(on ? 1 : 0); }

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

// Sample-arena overflow guard of the sync-free training step (the reference sizes its sample arrays exactly after a host
// sync, models/csrc/raymarching.cu:302-305; the fused step marches into a fixed-capacity arena and keeps the count on the
// device).  When the march produced more samples than the arena holds the tail rays were dropped, so the step's gradient is
// wrong: the guard records it (state[0] = this step overflowed, state[1] += 1, state[2] = max sample count seen) and poisons
// the first gradient element with NaN - every optimizer variant (single GPU, NCCL all-reduce, sharded peer exchange) then
// takes its existing non-finite path: the update is skipped on EVERY rank and the gradient is zeroed.  The host reads
// `state` lazily and regrows the arena (ncn_b200.fused.FusedStep).
namespace ncn {
__global__ void step_guard_kernel(const int32_t* __restrict__ counter, int64_t capacity, int32_t* __restrict__ state,
                                  float* __restrict__ poison) {
  const int32_t n = counter[0];
  const bool over = (int64_t)n > capacity;
  state[0] = over ? 1 : 0;
  if (over) state[1] += 1;
  if (n > state[2]) state[2] = n;
  if (over && poison != nullptr) poison[0] = __int_as_float(0x7fc00000);
}
}  // namespace ncn
extern "C" int ncn_step_guard(const int32_t* counter, int64_t capacity, int32_t* state, float* poison, ncn_stream_t stream) {
  NCN_CHECK_PTR(counter); NCN_CHECK_PTR(state);
  NCN_CHECK_SIZE(capacity >= 0);
  ncn::step_guard_kernel<<<1, 1, 0, ncn::as_stream(stream)>>>(counter, capacity, state, poison);
  NCN_LAUNCH_OK();
  return NCN_OK;
}

// measurement aid: node census of a captured CUDA graph (cudaGraph_t passed as void*), counts = [kernel, memcpy, memset, other].
// Child graphs are descended into.  Used by bench.py to report the launches of one replayed step from the graph itself.
static int count_nodes(cudaGraph_t g, int* counts) {
  size_t n = 0;
  NCN_CUDA(cudaGraphGetNodes(g, nullptr, &n));
  if (n == 0) return NCN_OK;
  cudaGraphNode_t* nodes = new cudaGraphNode_t[n];
  cudaError_t e = cudaGraphGetNodes(g, nodes, &n);
  for (size_t i = 0; e == cudaSuccess && i < n; ++i) {
    cudaGraphNodeType t;
    e = cudaGraphNodeGetType(nodes[i], &t);
    if (e != cudaSuccess) break;
    if (t == cudaGraphNodeTypeKernel) counts[0]++;
    else if (t == cudaGraphNodeTypeMemcpy) counts[1]++;
    else if (t == cudaGraphNodeTypeMemset) counts[2]++;
    else if (t == cudaGraphNodeTypeGraph) {
      cudaGraph_t child;
      e = cudaGraphChildGraphNodeGetGraph(nodes[i], &child);
      if (e == cudaSuccess) { int rc = count_nodes(child, counts); if (rc) { delete[] nodes; return rc; } }
    } else counts[3]++;
  }
  delete[] nodes;
  return e == cudaSuccess ? NCN_OK : (int)e;
}
extern "C" int ncn_graph_node_counts(void* cuda_graph, int* counts4_host) {
  NCN_CHECK_PTR(cuda_graph); NCN_CHECK_PTR(counts4_host);
  for (int i = 0; i < 4; ++i) counts4_host[i] = 0;
  return count_nodes((cudaGraph_t)cuda_graph, counts4_host);
}
