// Test-only stand-in for cuda_backend.cu so the CPU half of libb200dp (kfd.cpp, allocator.cpp, labels.cpp, ctx.cpp and
// the NVML-driven units backend, csrc/units_backend.hpp) links under -fsanitize=address,undefined / -fsanitize=thread
// with plain g++.  In-process CUDA entry points report "no GPU"; `cuda:probe=off` (NVML enumeration, MIG devices) works
// against whatever NVML B2DP_NVML_LIBRARY names (tests/native/nvml_stub.cpp).  The product library never contains this file.
#include "../../k8s-device-plugin_b200/csrc/internal.hpp"
#include "../../k8s-device-plugin_b200/csrc/units_backend.hpp"

namespace b2dp {
class CudaBackend {
public:
    CudaConfig cfg;
    Nvml nvml;
    std::unique_ptr<UnitsBackend> units;
    std::string driver_version;
    std::mutex probe_mu;  // one fan-out at a time, like the product backend
};

int cuda_backend_open(const CudaConfig& cfg, CudaBackend** out, std::string& err) {
    auto be = std::make_unique<CudaBackend>();
    be->cfg = cfg;
    be->nvml.load();
    int mode = cfg.probe_mode;  // like cuda_backend.cu: a node with a MIG-enabled GPU is probed through helpers
    if (mode == 0 && cfg.mig_auto && be->nvml.ok && be->nvml.device_count && be->nvml.handle_by_index && be->nvml.mig_mode) {
        unsigned cnt = 0;
        if (be->nvml.device_count(&cnt) == 0)
            for (unsigned i = 0; i < cnt; ++i) {
                void* h = nullptr;
                unsigned cur = 0, pend = 0;
                if (be->nvml.handle_by_index(i, &h) == 0 && be->nvml.mig_mode(h, &cur, &pend) == 0 && cur == 1) mode = 1;
            }
    }
    if (mode == 0) { err = "sanitizer build: no in-process cuda backend (probe=off / probe=helpers only)"; return B2DP_E_NOGPU; }
    int rc = units_open(cfg, be->nvml, mode == 1, &be->units, err);
    if (rc != B2DP_OK) return rc;
    char buf[96] = {0};
    if (be->nvml.driver_version && be->nvml.driver_version(buf, sizeof buf) == 0) be->driver_version = buf;
    *out = be.release();
    return B2DP_OK;
}
void cuda_backend_close(CudaBackend* be) { if (be) { units_close(be->units.get()); delete be; } }
int cuda_enumerate(CudaBackend* be, std::vector<Device>& out, std::string&) {
    out.clear();
    for (auto& u : be->units->units) out.push_back(u.dev);
    return B2DP_OK;
}
int cuda_node_health(CudaBackend* be) { return be->units->units.empty() ? 0 : 1; }
int cuda_probe(CudaBackend* be, const b2dp_probe_opts* opts, std::vector<b2dp_probe_result>& out, std::string& err) {
    std::lock_guard<std::mutex> l(be->probe_mu);
    if (!be->units->helpers) { err = "probe=off"; return B2DP_E_UNSUPPORTED; }
    return units_probe(be->units.get(), opts, out, err);
}
int cuda_inject_fault(CudaBackend*, int, uint64_t, uint32_t, std::string&) { return B2DP_E_UNSUPPORTED; }
int cuda_probe_reset(CudaBackend*, int, std::string&) { return B2DP_E_UNSUPPORTED; }
int cuda_probe_peek(CudaBackend*, int, uint64_t, uint32_t*, uint64_t, std::string&) { return B2DP_E_UNSUPPORTED; }
int cuda_p2p_matrix(CudaBackend* be, const b2dp_p2p_opts*, float* gbs, int32_t* lt, uint64_t* mism, int n, std::string& err) {
    if (n != (int)be->units->units.size()) { err = "n must equal the device count"; return B2DP_E_INVAL; }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            gbs[(size_t)i * n + j] = 0; mism[(size_t)i * n + j] = 0;
            lt[(size_t)i * n + j] = i == j ? 0 : units_link_type(be->nvml, be->units->units[i], be->units->units[j]);
        }
    return B2DP_OK;
}
int cuda_device_count(CudaBackend* be) { return (int)be->units->units.size(); }
void cuda_label_source(CudaBackend* be, LabelSource& src) {
    src.native = true;
    src.driver_version = be->driver_version;
    for (auto& u : be->units->units) {
        src.family.push_back(u.family); src.product_name.push_back(u.name); src.device_id.push_back(u.pci_device_id);
        src.vbios.push_back(u.vbios); src.firmware.push_back(u.firmware); src.vram_bytes.push_back(u.vram); src.sm_count.push_back(u.sms);
    }
}
float cuda_min_gbs(CudaBackend*) { return 0.f; }
std::string cuda_runtime_id(CudaBackend* be, const std::string& id, bool) {
    for (auto& u : be->units->units) if (u.dev.id == id) return u.uuid;
    return "";
}
int cuda_set_ref(CudaBackend*, int, float, std::string&) { return B2DP_E_UNSUPPORTED; }
void cuda_prearm(CudaBackend*) {}
int cuda_describe(CudaBackend* be, int device, b2dp_probe_info* o, std::string& err) {
    if (device < 0 || device >= (int)be->units->units.size()) { err = "device index out of range"; return B2DP_E_INVAL; }
    const Unit& u = be->units->units[device];
    o->total_memory = (uint64_t)u.vram; o->sm_count = (int32_t)u.sms; o->usable = u.broken ? 0 : 1;
    o->via_helper = be->units->helpers ? 1 : 0; o->slot_bytes = u.slot_bytes; o->gbs_cal = u.gbs_cal; o->gbs_ref = u.gbs_ref;
    copy_str(o->uuid, sizeof o->uuid, u.uuid);
    copy_str(o->name, sizeof o->name, u.name);
    return B2DP_OK;
}
bool cuda_device_paths(CudaBackend* be, const std::string& id, std::vector<std::string>& out) {
    for (auto& u : be->units->units)
        if (u.dev.id == id) {
            out.push_back("/dev/nvidia" + std::to_string(u.parent_minor));
            if (u.cap_gi >= 0) out.push_back("/dev/nvidia-caps/nvidia-cap" + std::to_string(u.cap_gi));
            if (u.cap_ci >= 0) out.push_back("/dev/nvidia-caps/nvidia-cap" + std::to_string(u.cap_ci));
            return true;
        }
    return false;
}
void cuda_set_health_event_callback(CudaBackend*, std::function<void()>) {}
}  // namespace b2dp
