// Test-only stand-in for cuda_backend.cu so the CPU half of libb200dp (kfd.cpp, allocator.cpp,
// labels.cpp, ctx.cpp) links under -fsanitize=address,undefined / -fsanitize=thread with plain
// g++.  Every cuda: entry point reports "no GPU"; the product library never contains this file.
#include "../../k8s-device-plugin_b200/csrc/internal.hpp"

namespace b2dp {
int cuda_backend_open(const CudaConfig&, CudaBackend**, std::string& err) { err = "sanitizer build: no cuda backend"; return B2DP_E_NOGPU; }
void cuda_backend_close(CudaBackend*) {}
int cuda_enumerate(CudaBackend*, std::vector<Device>&, std::string&) { return B2DP_E_NOGPU; }
int cuda_node_health(CudaBackend*) { return 0; }
int cuda_probe(CudaBackend*, const b2dp_probe_opts*, std::vector<b2dp_probe_result>&, std::string&) { return B2DP_E_NOGPU; }
int cuda_inject_fault(CudaBackend*, int, uint64_t, uint32_t, std::string&) { return B2DP_E_NOGPU; }
int cuda_probe_reset(CudaBackend*, int, std::string&) { return B2DP_E_NOGPU; }
int cuda_probe_peek(CudaBackend*, int, uint64_t, uint32_t*, uint64_t, std::string&) { return B2DP_E_NOGPU; }
int cuda_p2p_matrix(CudaBackend*, const b2dp_p2p_opts*, float*, int32_t*, uint64_t*, int, std::string&) { return B2DP_E_NOGPU; }
int cuda_device_count(CudaBackend*) { return 0; }
void cuda_label_source(CudaBackend*, LabelSource&) {}
float cuda_min_gbs(CudaBackend*) { return 0.f; }
std::string cuda_runtime_id(CudaBackend*, const std::string&, bool) { return ""; }
int cuda_set_ref(CudaBackend*, int, float, std::string&) { return B2DP_E_NOGPU; }
void cuda_prearm(CudaBackend*) {}
void cuda_set_health_event_callback(CudaBackend*, std::function<void()>) {}
}  // namespace b2dp
