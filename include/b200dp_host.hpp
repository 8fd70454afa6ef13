// b200dp_host.hpp -- C++17 host-side mirror of the reference's Go packages above the C ABI (b200dp.h).
//
// The reference's host is compiled Go; no Go toolchain exists in the build image, so the host side that a
// cgo package would provide is written here in C++ with the reference's own names, argument meaning and
// error behaviour, so that tests can read like the reference's (tests/native/host_mirror_test.cpp follows
// internal/pkg/allocator/*_test.go, internal/pkg/amdgpu/amdgpu_test.go, internal/pkg/plugin/plugin_test.go).
// Header-only, no logic of its own: every function is one C-ABI call plus the array-growing / string
// conversions a binding needs.  Go `(value, error)` returns become `std::pair<T, Error>`; `Error` is falsy
// for nil and carries the reference's exact message text (b2dp_strerror).
//
//   package amdgpu    internal/pkg/amdgpu/amdgpu.go      -> namespace b200dp::amdgpu
//   package allocator internal/pkg/allocator/*.go        -> namespace b200dp::allocator
//   package exporter  internal/pkg/exporter/health.go    -> namespace b200dp::exporter
//   package plugin    internal/pkg/plugin/plugin.go      -> namespace b200dp::plugin
//   node labeller     cmd/k8s-node-labeller/main.go      -> namespace b200dp::labeller
#pragma once
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "b200dp.h"

namespace b200dp {

// Go `error`: nil == ok().  what() is the reference's message where the reference defines one.
struct Error {
    int code = B2DP_OK;
    std::string msg;
    explicit operator bool() const { return code != B2DP_OK; }  // `if err != nil`
    const std::string& what() const { return msg; }
    static Error from(int rc, b2dp_ctx* c = nullptr) {
        Error e;
        e.code = rc;
        if (rc != B2DP_OK) {
            e.msg = b2dp_strerror(rc);  // allocator codes: the reference's exact strings
            // codes whose static text is generic carry the detail recorded by the failing call on this thread
            const bool generic = rc == B2DP_E_CUDA || rc == B2DP_E_NOGPU || rc == B2DP_E_NODRIVER || rc == B2DP_E_INVAL ||
                                 rc == B2DP_E_UNSUPPORTED || rc == B2DP_E_IO;
            const char* le = generic ? b2dp_last_error(c) : nullptr;
            if (le && *le) e.msg = le;
        }
        return e;
    }
};

namespace detail {
inline std::vector<const char*> cstrs(const std::vector<std::string>& v) {
    std::vector<const char*> p;
    for (auto& s : v) p.push_back(s.c_str());
    return p;
}
// call(arr, cap, &n) until the caller-owned array is large enough (B2DP_E_NOSPC reports the need)
template <class T, class F>
inline int grow(std::vector<T>& out, F&& call, int first_cap = 64) {
    out.assign((size_t)first_cap, T{});
    int n = 0;
    int rc = call(out.data(), (int)out.size(), &n);
    if (rc == B2DP_E_NOSPC) {
        out.assign((size_t)n, T{});
        rc = call(out.data(), (int)out.size(), &n);
    }
    out.resize(rc == B2DP_OK ? (size_t)n : 0);
    return rc;
}
}  // namespace detail

// One process-wide backend handle ("kfd:<sysroot>" parity mode, "cuda:[opts]" on a B200 node).
class Context {
public:
    static std::pair<std::shared_ptr<Context>, Error> Open(const std::string& uri) {
        b2dp_ctx* h = nullptr;
        const int rc = b2dp_open(uri.c_str(), &h);
        if (rc != B2DP_OK) return {nullptr, Error::from(rc)};
        return {std::shared_ptr<Context>(new Context(h)), Error{}};
    }
    ~Context() { b2dp_close(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    b2dp_ctx* handle() const { return h_; }

private:
    explicit Context(b2dp_ctx* h) : h_(h) {}
    b2dp_ctx* h_;
};
using Ctx = std::shared_ptr<Context>;

// ---- package amdgpu ------------------------------------------------------------------------------------
namespace amdgpu {
// the per-device map of amdgpu.go:216/264 {card, renderD, devID, computePartitionType, memoryPartitionType, numaNode, nodeId}
struct GPU {
    int card = 0, renderD = 0, numaNode = 0, nodeId = 0;
    std::string devID, computePartitionType, memoryPartitionType;
};

// amdgpu.go:149-268 GetAMDGPUs() -- keyed by kubelet device id
inline std::map<std::string, GPU> GetAMDGPUs(const Ctx& c) {
    std::vector<b2dp_device> d;
    detail::grow(d, [&](b2dp_device* a, int cap, int* n) { return b2dp_enumerate(c->handle(), a, cap, n); });
    std::map<std::string, GPU> out;
    for (auto& x : d) out[x.id] = GPU{x.card, x.render_d, x.numa_node, x.node_id, x.dev_id, x.compute_partition, x.memory_partition};
    return out;
}
// amdgpu.go:442-463 ParseTopologyProperties(path, regexp `<key>\s(\d+)`)
inline std::pair<int64_t, Error> ParseTopologyProperties(const std::string& path, const std::string& key) {
    int64_t v = 0;
    const int rc = b2dp_parse_topology_property(path.c_str(), key.c_str(), &v);
    return {v, Error::from(rc)};
}
// amdgpu.go:101-146 GetDevIdsFromTopology(topoRootParam...)
inline std::map<int, std::string> GetDevIdsFromTopology(const std::string& topoRoot = "/sys/class/kfd/kfd") {
    std::vector<int32_t> minors(64);
    std::vector<char> ids(64 * 24);
    int n = 0;
    int rc = b2dp_dev_ids_from_topology(topoRoot.c_str(), minors.data(), (char (*)[24])ids.data(), 64, &n);
    if (rc == B2DP_E_NOSPC) {
        minors.assign((size_t)n, 0);
        ids.assign((size_t)n * 24, 0);
        rc = b2dp_dev_ids_from_topology(topoRoot.c_str(), minors.data(), (char (*)[24])ids.data(), n, &n);
    }
    std::map<int, std::string> out;
    for (int i = 0; rc == B2DP_OK && i < n; ++i) out[minors[(size_t)i]] = &ids[(size_t)i * 24];
    return out;
}
// amdgpu.go:496-538 GetNodeIdsFromTopology(topoRootParam...)
inline std::map<int, int> GetNodeIdsFromTopology(const std::string& topoRoot = "/sys/class/kfd/kfd") {
    std::vector<int32_t> minors(64), nodes(64);
    int n = 0;
    int rc = b2dp_node_ids_from_topology(topoRoot.c_str(), minors.data(), nodes.data(), 64, &n);
    if (rc == B2DP_E_NOSPC) {
        minors.assign((size_t)n, 0);
        nodes.assign((size_t)n, 0);
        rc = b2dp_node_ids_from_topology(topoRoot.c_str(), minors.data(), nodes.data(), n, &n);
    }
    std::map<int, int> out;
    for (int i = 0; rc == B2DP_OK && i < n; ++i) out[minors[(size_t)i]] = nodes[(size_t)i];
    return out;
}
// amdgpu.go:270-285 UniquePartitionConfigCount(GetAMDGPUs())
inline std::map<std::string, int> UniquePartitionConfigCount(const Ctx& c) {
    std::vector<b2dp_kv_count> kv;
    detail::grow(kv, [&](b2dp_kv_count* a, int cap, int* n) { return b2dp_partition_histogram(c->handle(), a, cap, n); });
    std::map<std::string, int> out;
    for (auto& x : kv) out[x.key] = x.count;
    return out;
}
// amdgpu.go:287-293 IsHomogeneous()
inline bool IsHomogeneous(const Ctx& c) {
    int32_t v = 0;
    return b2dp_is_homogeneous(c->handle(), &v) == B2DP_OK && v;
}
// amdgpu.go:295-328
inline bool IsComputePartitionSupported(const Ctx& c) { int32_t v = 0; return b2dp_partition_supported(c->handle(), 0, &v) == B2DP_OK && v; }
inline bool IsMemoryPartitionSupported(const Ctx& c) { int32_t v = 0; return b2dp_partition_supported(c->handle(), 1, &v) == B2DP_OK && v; }
// amdgpu.go:467-490 parseDebugFSFirmwareInfo(path) -> (feature versions, firmware versions)
inline std::pair<std::map<std::string, uint32_t>, std::map<std::string, uint32_t>> parseDebugFSFirmwareInfo(const std::string& path) {
    std::vector<b2dp_fw_entry> fw;
    detail::grow(fw, [&](b2dp_fw_entry* a, int cap, int* n) { return b2dp_parse_debugfs_firmware_info(path.c_str(), a, cap, n); });
    std::pair<std::map<std::string, uint32_t>, std::map<std::string, uint32_t>> out;
    for (auto& e : fw) { out.first[e.name] = e.feature; out.second[e.name] = e.firmware; }
    return out;
}
}  // namespace amdgpu

// ---- package allocator -----------------------------------------------------------------------------------
namespace allocator {
// allocator/device.go:56-65
struct Device {
    std::string Id;
    int NodeId = 0, NumaNode = 0;
    std::string DevId;
    int Card = 0, RenderD = 0;
    std::string ComputePartitionType, MemoryPartitionType;
};
inline b2dp_device to_abi(const Device& d) {
    b2dp_device x{};
    snprintf(x.id, sizeof x.id, "%s", d.Id.c_str());
    snprintf(x.dev_id, sizeof x.dev_id, "%s", d.DevId.c_str());
    x.card = d.Card; x.render_d = d.RenderD; x.node_id = d.NodeId; x.numa_node = d.NumaNode;
    snprintf(x.compute_partition, sizeof x.compute_partition, "%s", d.ComputePartitionType.c_str());
    snprintf(x.memory_partition, sizeof x.memory_partition, "%s", d.MemoryPartitionType.c_str());
    return x;
}

// allocator.go:27-30 Policy; besteffort_policy.go:45-151
class BestEffortPolicy {
public:
    BestEffortPolicy() { b2dp_allocator_new(&a_); }  // NewBestEffortPolicy()
    ~BestEffortPolicy() { b2dp_allocator_free(a_); }
    BestEffortPolicy(const BestEffortPolicy&) = delete;
    BestEffortPolicy& operator=(const BestEffortPolicy&) = delete;

    // besteffort_policy.go:70-86 Init(devs, topoDir)
    Error Init(const std::vector<Device>& devs, const std::string& topoDir) {
        std::vector<b2dp_device> d;
        for (auto& x : devs) d.push_back(to_abi(x));
        return Error::from(b2dp_allocator_init(a_, d.data(), (int)d.size(), topoDir.c_str()));
    }
    // besteffort_policy.go:88-151 Allocate(availableIds, requiredIds, size)
    std::pair<std::vector<std::string>, Error> Allocate(const std::vector<std::string>& available,
                                                        const std::vector<std::string>& required, int size) {
        auto av = detail::cstrs(available), rq = detail::cstrs(required);
        std::vector<char> out((available.size() + 1) * 64);
        int n = 0;
        const int rc = b2dp_allocator_allocate(a_, av.data(), (int)av.size(), rq.data(), (int)rq.size(), size,
                                               (char (*)[64])out.data(), (int)available.size() + 1, &n);
        std::vector<std::string> ids;
        for (int i = 0; rc == B2DP_OK && i < n; ++i) ids.emplace_back(&out[(size_t)i * 64]);
        return {ids, Error::from(rc)};
    }
    // device.go:220-252 fetchAllPairWeights result: p2pWeights[from][to]
    std::map<int, std::map<int, int>> PairWeights() const {
        std::vector<b2dp_pair_weight> pw;
        int rows = 0;
        detail::grow(pw, [&](b2dp_pair_weight* a, int cap, int* n) { return b2dp_allocator_pair_weights(a_, a, cap, n, &rows); }, 4096);
        std::map<int, std::map<int, int>> out;
        for (auto& w : pw) out[w.node_from][w.node_to] = w.weight;
        return out;
    }
    // device.go:287-304 len(groupPartitionsByDevId(devices))
    int GroupCount() const { int32_t g = 0; b2dp_allocator_group_count(a_, &g); return g; }
    // device.go:353-442 getCandidateDeviceSubsets: (number of candidates, best total weight)
    std::pair<std::pair<int, int>, Error> CandidateSubsets(const std::vector<std::string>& available,
                                                           const std::vector<std::string>& required, int size) const {
        auto av = detail::cstrs(available), rq = detail::cstrs(required);
        int32_t n = 0, best = 0;
        const int rc = b2dp_allocator_candidates(a_, av.data(), (int)av.size(), rq.data(), (int)rq.size(), size, &n, &best);
        return {{n, best}, Error::from(rc)};
    }

private:
    b2dp_allocator* a_ = nullptr;
};
inline std::unique_ptr<BestEffortPolicy> NewBestEffortPolicy() { return std::make_unique<BestEffortPolicy>(); }
}  // namespace allocator

// ---- package exporter ----------------------------------------------------------------------------------
namespace exporter {
constexpr const char* Healthy = "Healthy";      // pluginapi.Healthy
constexpr const char* Unhealthy = "Unhealthy";  // pluginapi.Unhealthy
struct PluginDevice { std::string ID, Health; int64_t NumaNode = 0; };  // pluginapi.Device (ID, Health, Topology.Nodes[0].ID)

// health.go:86-106 PopulatePerGPUDHealth(devs, defaultHealth) with the exporter's answer injectable:
// hMap == nullptr reproduces "socket absent / RPC failed".
inline void PopulatePerGPUDHealth(std::vector<PluginDevice>& devs, const std::string& defaultHealth,
                                  const std::map<std::string, std::string>* hMap) {
    std::vector<char> ids(devs.size() * 64 + 64, 0), src((hMap ? hMap->size() : 0) * 64 + 64, 0);
    std::vector<int32_t> sh, out(devs.size() + 1, 0);
    for (size_t i = 0; i < devs.size(); ++i) snprintf(&ids[i * 64], 64, "%s", devs[i].ID.c_str());
    size_t j = 0;
    if (hMap)
        for (auto& kv : *hMap) { snprintf(&src[j++ * 64], 64, "%s", kv.first.c_str()); sh.push_back(kv.second == Healthy); }
    sh.push_back(0);
    b2dp_merge_health((const char (*)[64])ids.data(), (int)devs.size(), defaultHealth == Healthy, hMap != nullptr,
                      (const char (*)[64])src.data(), sh.data(), (int)j, out.data());
    for (size_t i = 0; i < devs.size(); ++i) devs[i].Health = out[i] ? Healthy : Unhealthy;
}
}  // namespace exporter

// ---- package plugin ----------------------------------------------------------------------------------------
namespace plugin {
inline int countGPUDevFromTopology(const std::string& topoRoot = "/sys/class/kfd/kfd") {  // plugin.go:123-159
    int32_t v = 0;
    b2dp_count_gpu_dev_from_topology(topoRoot.c_str(), &v);
    return v;
}
inline bool simpleHealthCheck(const std::string& topoRoot = "/sys/class/kfd/kfd") {  // plugin.go:161-206
    int32_t v = 0;
    return b2dp_simple_health_check(topoRoot.c_str(), &v) == B2DP_OK && v;
}

struct DevicePluginOptions { bool PreStartRequired = false, GetPreferredAllocationAvailable = false; };
struct ContainerPreferredAllocationRequest { std::vector<std::string> AvailableDeviceIDs, MustIncludeDeviceIDs; int AllocationSize = 0; };
struct DeviceSpec { std::string ContainerPath, HostPath, Permissions; };
struct ContainerAllocateResponse { std::vector<DeviceSpec> Devices; };

// plugin.go:41-48 AMDGPUPlugin: one per resource name ("gpu" or "<compute>_<memory>")
class AMDGPUPlugin {
public:
    AMDGPUPlugin(Ctx c, std::string resource) : Resource(std::move(resource)), ctx_(std::move(c)) {}
    std::string Resource;
    bool allocatorInitError = false;

    // plugin.go:82-91: allocator init failure degrades to the kubelet's default allocation
    Error Start() {
        const int rc = b2dp_start(ctx_->handle());
        allocatorInitError = rc != B2DP_OK;
        return Error{};
    }
    Error Stop() { return Error{}; }  // plugin.go:117-119
    // plugin.go:210-217
    DevicePluginOptions GetDevicePluginOptions() const { return {false, !allocatorInitError}; }
    // plugin.go:222-224
    Error PreStartContainer() { return Error{}; }

    // plugin.go:229-330.  `send(serialized ListAndWatchResponse)` is s.Send; `next()` blocks for the next
    // event and returns true for a heartbeat tick (<-p.Heartbeat), false for the stop signal (<-p.signal).
    Error ListAndWatch(const std::function<void(const std::string&)>& send, const std::function<bool()>& next,
                       uint32_t heartbeat_flags = 0) {
        auto cycle = [&](uint32_t flags) -> Error {
            b2dp_cycle_opts o{};
            o.flags = flags;
            std::string buf(1 << 14, '\0');
            size_t len = 0;
            b2dp_cycle_stats st{};
            int rc = b2dp_list_and_watch(ctx_->handle(), Resource.c_str(), &o, (uint8_t*)buf.data(), buf.size(), &len, &st);
            if (rc == B2DP_E_NOSPC) {
                buf.assign(len, '\0');
                rc = b2dp_list_and_watch(ctx_->handle(), Resource.c_str(), &o, (uint8_t*)buf.data(), buf.size(), &len, &st);
            }
            if (rc != B2DP_OK) return Error::from(rc, ctx_->handle());
            last_stats = st;
            if (st.n_devices || st.homogeneous) send(buf.substr(0, len));  // plugin.go:296-298
            return Error{};
        };
        if (Error e = cycle(B2DP_LW_INITIAL)) return e;
        while (next())
            if (Error e = cycle(B2DP_LW_HEARTBEAT | heartbeat_flags)) return e;
        return Error{};  // returning unregisters the plugin (plugin.go:327-329)
    }
    b2dp_cycle_stats last_stats{};

    // plugin.go:337-351
    std::pair<std::vector<std::vector<std::string>>, Error> GetPreferredAllocation(
        const std::vector<ContainerPreferredAllocationRequest>& reqs) {
        std::vector<std::vector<std::string>> out;
        for (auto& r : reqs) {
            auto av = detail::cstrs(r.AvailableDeviceIDs), mi = detail::cstrs(r.MustIncludeDeviceIDs);
            std::vector<char> ids((r.AvailableDeviceIDs.size() + 1) * 64);
            int n = 0;
            const int rc = b2dp_preferred_allocation(ctx_->handle(), av.data(), (int)av.size(), mi.data(), (int)mi.size(),
                                                     r.AllocationSize, (char (*)[64])ids.data(), (int)av.size() + 1, &n);
            if (rc != B2DP_OK) {  // plugin.go:341-344
                Error e;
                e.code = rc;
                e.msg = std::string("unable to get preferred allocation list. Error:") + b2dp_strerror(rc);
                return {{}, e};
            }
            out.emplace_back();
            for (int i = 0; i < n; ++i) out.back().emplace_back(&ids[(size_t)i * 64]);
        }
        return {out, Error{}};
    }
    // plugin.go:356-393
    std::pair<std::vector<ContainerAllocateResponse>, Error> Allocate(const std::vector<std::vector<std::string>>& reqs) {
        std::vector<ContainerAllocateResponse> out;
        for (auto& ids : reqs) {
            auto p = detail::cstrs(ids);
            std::vector<b2dp_devspec> specs;
            const int rc = detail::grow(specs, [&](b2dp_devspec* a, int cap, int* n) {
                return b2dp_device_specs(ctx_->handle(), p.data(), (int)p.size(), a, cap, n);
            });
            if (rc != B2DP_OK) return {{}, Error::from(rc, ctx_->handle())};
            out.emplace_back();
            for (auto& s : specs) out.back().Devices.push_back({s.container_path, s.host_path, s.permissions});
        }
        return {out, Error{}};
    }

private:
    Ctx ctx_;
};

// plugin.go:398-438
class AMDGPULister {
public:
    explicit AMDGPULister(Ctx c) : ctx_(std::move(c)) {}
    std::string GetResourceNamespace() const { return "amd.com"; }  // plugin.go:406-408
    std::unique_ptr<AMDGPUPlugin> NewPlugin(const std::string& resourceLastName) { return std::make_unique<AMDGPUPlugin>(ctx_, resourceLastName); }

private:
    Ctx ctx_;
};

// cmd/k8s-device-plugin/main.go:53-91 getResourceList(strategy)
inline std::pair<std::vector<std::string>, Error> getResourceList(const Ctx& c, const std::string& strategy) {
    std::vector<char> names(64 * 64);
    int n = 0;
    int rc = b2dp_resource_list(c->handle(), strategy.c_str(), (char (*)[64])names.data(), 64, &n);
    if (rc == B2DP_E_NOSPC) {
        names.assign((size_t)n * 64, 0);
        rc = b2dp_resource_list(c->handle(), strategy.c_str(), (char (*)[64])names.data(), n, &n);
    }
    std::vector<std::string> out;
    for (int i = 0; rc == B2DP_OK && i < n; ++i) out.emplace_back(&names[(size_t)i * 64]);
    return {out, Error::from(rc, c->handle())};
}
}  // namespace plugin

// ---- cmd/k8s-node-labeller/main.go ----------------------------------------------------------------------------
namespace labeller {
using Labels = std::map<std::string, std::string>;
inline Labels to_map(const std::vector<b2dp_label>& v) { Labels m; for (auto& l : v) m[l.key] = l.value; return m; }
// main.go:87-108 createLabels(kind, entries)
inline Labels createLabels(const std::string& kind, const std::map<std::string, int>& entries) {
    std::vector<b2dp_kv_count> kv;
    for (auto& e : entries) { b2dp_kv_count x{}; snprintf(x.key, sizeof x.key, "%s", e.first.c_str()); x.count = e.second; kv.push_back(x); }
    std::vector<b2dp_label> out;
    detail::grow(out, [&](b2dp_label* a, int cap, int* n) { return b2dp_create_labels(kind.c_str(), kv.data(), (int)kv.size(), a, cap, n); });
    return to_map(out);
}
// main.go:383-397 generateLabels(): enabled = the labeller's bool flags that are set
inline Labels generateLabels(const Ctx& c, const std::vector<std::string>& enabled) {
    std::string csv;
    for (auto& e : enabled) csv += (csv.empty() ? "" : ",") + e;
    std::vector<b2dp_label> out;
    detail::grow(out, [&](b2dp_label* a, int cap, int* n) { return b2dp_generate_labels(c->handle(), csv.c_str(), a, cap, n); }, 256);
    return to_map(out);
}
// main.go:55-74 removeOldNodeLabels(node) on the node's label map
inline void removeOldNodeLabels(Labels& nodeLabels) {
    std::vector<b2dp_label> v;
    for (auto& kv : nodeLabels) { b2dp_label l{}; snprintf(l.key, sizeof l.key, "%s", kv.first.c_str()); snprintf(l.value, sizeof l.value, "%s", kv.second.c_str()); v.push_back(l); }
    int n = 0;
    if (b2dp_remove_old_node_labels(v.data(), (int)v.size(), &n) != B2DP_OK) return;
    v.resize((size_t)n);
    nodeLabels = to_map(v);
}
}  // namespace labeller

}  // namespace b200dp
