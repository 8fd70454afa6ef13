// The reference's own unit tests, re-written against the C++ host mirror (include/b200dp_host.hpp) of its
// Go packages -- same test names, same fixtures, same expected values:
//   internal/pkg/amdgpu/amdgpu_test.go:122-163      TestParseTopologyProperties
//   internal/pkg/amdgpu/amdgpu_test.go:165-218      TestParseDebugFSFirmwareInfo
//   internal/pkg/amdgpu/amdgpu_test.go:220-264      TestRenderDevIdsFromTopology (key set + grouping; values per amdgpu.go:139-142)
//   internal/pkg/plugin/plugin_test.go:23-30        TestCountGPUDevFromTopology
//   internal/pkg/allocator/device_test.go:80-169    TestPairWeightCalculation / TestGroupPartitionsByDevId / TestGetSubsetsMethod
//   internal/pkg/allocator/besteffort_policy_test.go:25-216  TestBestPolicyAllocator (all 17 cases)
//   cmd/k8s-node-labeller/main_test.go:11-125       TestInitLabelLists / TestRemoveOldNodeLabels
// plus the plugin-level flow (Start / ListAndWatch / GetPreferredAllocation / Allocate) on a sysroot.
// usage: host_mirror_test <fixtures_dir> [<sysroot built around topo-mi300-cpx>]
#include <algorithm>
#include <cstdio>
#include <set>
#include <string>
#include <vector>

#include "../../include/b200dp_host.hpp"

using namespace b200dp;

static int g_failed = 0, g_checks = 0;
static const char* g_test = "";
#define CHECK(c) do { ++g_checks; if (!(c)) { fprintf(stderr, "--- FAIL: %s (%s:%d): %s\n", g_test, __FILE__, __LINE__, #c); ++g_failed; } } while (0)
#define RUN(fn) do { g_test = #fn; fn(); printf("ok   %s\n", #fn); } while (0)

static std::string testdata;

// device_test.go:43-67 getTestDevices
static std::vector<allocator::Device> getTestDevices(int devCount, int partitionCountPerDev, int numanodeCount, int startNodeId, int endNodeId) {
    std::vector<allocator::Device> res;
    int nodeId = startNodeId;
    for (int i = 0; i < devCount; ++i) {
        const int numa = devCount / numanodeCount;
        for (int j = 0; j < partitionCountPerDev; ++j) {
            std::string id = "amdgpu_xcp_" + std::to_string(i * 8 + j);
            if (j == 0) id = "test" + std::to_string(i + 1);
            if (nodeId > endNodeId) break;
            allocator::Device d;
            d.Id = id; d.NodeId = nodeId; d.NumaNode = i / numa; d.DevId = std::to_string(i);
            res.push_back(d);
            ++nodeId;
        }
    }
    return res;
}
static std::vector<std::string> ids_of(const std::vector<allocator::Device>& devs) {
    std::vector<std::string> v;
    for (auto& d : devs) v.push_back(d.Id);
    return v;
}
static std::vector<std::string> T(std::initializer_list<int> n) { std::vector<std::string> v; for (int i : n) v.push_back("test" + std::to_string(i)); return v; }
static std::vector<std::string> X(std::initializer_list<int> n) { std::vector<std::string> v; for (int i : n) v.push_back("amdgpu_xcp_" + std::to_string(i)); return v; }
static std::vector<std::string> cat(std::vector<std::string> a, const std::vector<std::string>& b) { a.insert(a.end(), b.begin(), b.end()); return a; }
static std::vector<std::string> sorted(std::vector<std::string> v) { std::sort(v.begin(), v.end()); return v; }

// ---- amdgpu_test.go -----------------------------------------------------------------------------------
static void TestParseTopologyProperties() {
    const std::string tp = testdata + "/topology-parsing/topology/nodes";
    auto r = amdgpu::ParseTopologyProperties(tp + "/1/mem_banks/0/properties", "size_in_bytes");
    CHECK(!r.second && r.first == 17163091968LL);
    r = amdgpu::ParseTopologyProperties(tp + "/1/mem_banks/0/properties", "flags");
    CHECK(!r.second && r.first == 0);
    r = amdgpu::ParseTopologyProperties(tp + "/2/properties", "simd_count");
    CHECK(!r.second && r.first == 256);
    r = amdgpu::ParseTopologyProperties(tp + "/2/properties", "simd_id_base");
    CHECK(!r.second && r.first == 2147487744LL);
    r = amdgpu::ParseTopologyProperties(tp + "/2/properties", "asdf");
    CHECK(bool(r.second));  // "Topology property not found"
    r = amdgpu::ParseTopologyProperties(tp + "/1/properties", "unique_id");  // > MaxInt64: strconv.ErrRange, clamped
    CHECK(r.second.code == B2DP_E_RANGE && r.first == INT64_MAX);
}

static void TestParseDebugFSFirmwareInfo() {
    auto r = amdgpu::parseDebugFSFirmwareInfo(testdata + "/debugfs-parsing/amdgpu_firmware_info");
    const std::map<std::string, uint32_t> expFeat = {{"VCE", 0}, {"UVD", 0}, {"MC", 0}, {"ME", 35}, {"PFP", 35}, {"CE", 35}, {"RLC", 0},
                                                     {"MEC", 33}, {"MEC2", 33}, {"SOS", 0}, {"ASD", 0}, {"SMC", 0}, {"SDMA0", 40}, {"SDMA1", 40}};
    const std::map<std::string, uint32_t> expFw = {{"VCE", 0x352d0400}, {"UVD", 0x01571100}, {"MC", 0}, {"ME", 0x94}, {"PFP", 0xa4},
                                                   {"CE", 0x4a}, {"RLC", 0x58}, {"MEC", 0x160}, {"MEC2", 0x160}, {"SOS", 0x161a92},
                                                   {"ASD", 0x16129a}, {"SMC", 0x1c2800}, {"SDMA0", 0x197}, {"SDMA1", 0x197}};
    CHECK(r.first == expFeat);
    CHECK(r.second == expFw);
}

static void TestRenderDevIdsFromTopology() {
    auto got = amdgpu::GetDevIdsFromTopology(testdata + "/topology-parsing-mi308");
    CHECK(got.size() == 32);
    std::map<std::string, std::vector<int>> groups;
    for (auto& kv : got) groups[kv.second].push_back(kv.first);
    CHECK(groups.size() == 8);
    int base = 128;
    const char* want[8] = {"0000:0a:00:0", "0000:80:00:0", "0000:a4:00:0", "0000:c8:00:0", "0001:0b:00:0", "0001:81:00:0", "0001:a5:00:0", "0001:c9:00:0"};
    for (int g = 0; g < 8; ++g, base += 8) {
        CHECK(got[base] == want[g]);
        CHECK((groups[want[g]] == std::vector<int>{base, base + 1, base + 2, base + 3}));
    }
    CHECK(amdgpu::GetDevIdsFromTopology(testdata + "/topology-parsing").empty());  // no `domain` key in that capture
    auto nodes = amdgpu::GetNodeIdsFromTopology(testdata + "/topology-parsing");
    CHECK((nodes == std::map<int, int>{{128, 1}, {129, 2}}));
}

// ---- plugin_test.go -------------------------------------------------------------------------------------
static void TestCountGPUDevFromTopology() {
    CHECK(plugin::countGPUDevFromTopology(testdata + "/topology-parsing") == 2);
    CHECK(plugin::countGPUDevFromTopology(testdata + "/topology-parsing-mi308") == 32);
    CHECK(plugin::countGPUDevFromTopology(testdata + "/topo-mi300-cpx") == 63);
    CHECK(!plugin::simpleHealthCheck(testdata + "/topology-parsing"));  // capture has no gfx_target_version
    CHECK(plugin::simpleHealthCheck(testdata + "/topo-mi300-cpx"));
}

// ---- device_test.go -------------------------------------------------------------------------------------
static const std::string mi308() { return testdata + "/topology-parsing-mi308/topology/nodes"; }
static const std::string mi210() { return testdata + "/topo-mi210-xgmi-pcie/nodes"; }
static const std::string cpx() { return testdata + "/topo-mi300-cpx/topology/nodes"; }

static void TestPairWeightCalculation() {
    {
        allocator::BestEffortPolicy p;
        Error err = p.Init({}, mi308());  // device_test.go:80-88: an empty device list is an error
        CHECK(bool(err) && err.what() == "Devices list is empty. Unable to calculate pair wise weights");
    }
    allocator::BestEffortPolicy p;
    CHECK(!p.Init(getTestDevices(4, 8, 2, 2, 33), mi308()));
    auto w = p.PairWeights();
    CHECK(w.size() == 31);  // device_test.go:105
    std::map<int, int> hist;
    int pairs = 0;
    for (auto& r : w) for (auto& c : r.second) { hist[c.second]++; ++pairs; }
    CHECK(pairs == 496 && (hist == std::map<int, int>{{30, 112}, {40, 128}, {50, 256}}));
}

static void TestGroupPartitionsByDevId() {
    allocator::BestEffortPolicy p;
    CHECK(!p.Init(getTestDevices(4, 8, 2, 2, 33), mi308()));
    CHECK(p.GroupCount() == 4);  // device_test.go:110-123
}

static void TestGetSubsetsMethod() {  // device_test.go:125-169
    auto devices = getTestDevices(4, 8, 2, 2, 33);
    allocator::BestEffortPolicy p;
    CHECK(!p.Init(devices, mi308()));
    auto r = p.CandidateSubsets(ids_of(devices), {}, 3);
    CHECK(!r.second && r.first.first == 4);
    r = p.CandidateSubsets(ids_of(devices), {}, 12);
    CHECK(!r.second && r.first.first == 12);
}

// ---- besteffort_policy_test.go:25-216 -----------------------------------------------------------------------------
struct Case {
    const char* name;
    std::string topo;
    std::vector<allocator::Device> devices;
    int size;
    std::vector<std::string> available, filtered, required, expected;  // available empty = all; expected empty = size only
};

static void TestBestPolicyAllocator() {
    const auto d308 = getTestDevices(4, 8, 2, 2, 33), d210 = getTestDevices(8, 1, 2, 2, 9), dcpx = getTestDevices(8, 8, 2, 2, 64);
    const auto sameNuma = T({3, 4, 5, 6, 7, 8});
    const std::vector<Case> cases = {
        {"mi308 1 partition", mi308(), d308, 1, {}, {}, {}, {}},
        {"mi308 3 partitions", mi308(), d308, 3, {}, {}, {}, {}},
        {"mi308 12 partitions", mi308(), d308, 12, {}, {}, {}, {}},
        {"mi210 1 gpu", mi210(), d210, 1, {}, {}, {}, T({1})},
        {"mi210 3 gpus", mi210(), d210, 3, {}, {}, {}, T({1, 2, 3})},
        {"mi210 5 gpus", mi210(), d210, 5, {}, {}, {}, T({1, 2, 3, 4, 5})},
        {"mi210 3 gpus, same numa available", mi210(), d210, 3, sameNuma, {}, {}, T({5, 6, 7})},
        {"cpx 1 partition", cpx(), dcpx, 1, {}, {}, {}, T({8})},
        {"cpx 3 partitions", cpx(), dcpx, 3, {}, {}, {}, cat(T({8}), X({57, 58}))},
        {"cpx 5 partitions", cpx(), dcpx, 5, {}, {}, {}, cat(T({8}), X({57, 58, 59, 60}))},
        {"cpx 3, same numa available", cpx(), dcpx, 3, sameNuma, {}, {}, T({5, 6, 7})},
        {"cpx 3, same numa available, required", cpx(), dcpx, 3, sameNuma, {}, T({5}), T({5, 6, 7})},
        {"cpx 30 partitions", cpx(), dcpx, 30, {}, {}, {}, {}},
        {"cpx 8 partitions", cpx(), dcpx, 8, {}, {}, {}, cat(T({1}), X({1, 2, 3, 4, 5, 6, 7}))},
        {"cpx 7 partitions", cpx(), dcpx, 7, {}, {}, {}, cat(T({8}), X({57, 58, 59, 60, 61, 62}))},
        {"cpx 4 after 3 taken", cpx(), dcpx, 4, {}, cat(T({8}), X({57, 58})), {}, X({59, 60, 61, 62})},
        {"cpx 10 after 6 taken", cpx(), dcpx, 10, {}, cat(T({1, 2, 3, 4, 8}), X({57})), {},
         cat(T({5}), X({33, 34, 35, 36, 37, 38, 39, 58, 59}))},
    };
    for (const auto& tc : cases) {
        auto pol = allocator::NewBestEffortPolicy();
        Error err = pol->Init(tc.devices, tc.topo);
        CHECK(!err);
        std::vector<std::string> avail;
        for (auto& id : tc.available.empty() ? ids_of(tc.devices) : tc.available)
            if (std::find(tc.filtered.begin(), tc.filtered.end(), id) == tc.filtered.end()) avail.push_back(id);
        auto r = pol->Allocate(avail, tc.required, tc.size);
        if (r.second || (int)r.first.size() != tc.size || (!tc.expected.empty() && sorted(r.first) != sorted(tc.expected))) {
            fprintf(stderr, "--- FAIL: TestBestPolicyAllocator/%s: err=%s got %zu ids\n", tc.name, r.second.what().c_str(), r.first.size());
            ++g_failed;
        }
        ++g_checks;
    }
    // besteffort_policy.go:36-43: the fixed error strings
    auto pol = allocator::NewBestEffortPolicy();
    CHECK(pol->Allocate(ids_of(d210), {}, 2).second.what() == "Init method must be called before Allocate");
    CHECK(!pol->Init(d210, mi210()));
    CHECK(pol->Allocate(ids_of(d210), {}, 0).second.what() == "allocation size can not be negative");
    CHECK(pol->Allocate(T({1, 2}), {}, 3).second.what() == "available devices count less than allocation size");
    CHECK(pol->Allocate(ids_of(d210), T({1, 2, 3}), 2).second.what() == "must_include devices size is more than allocation size");
}

// ---- cmd/k8s-node-labeller/main_test.go ---------------------------------------------------------------------------
static void TestRemoveOldNodeLabels() {
    labeller::Labels node = {{"amd.com/gpu.cu-count", "104"}, {"amd.com/gpu.vram", "64G"}, {"beta.amd.com/gpu.cu-count", "104"},
                             {"beta.amd.com/gpu.cu-count.104", "1"}, {"beta.amd.com/gpu.family", "AI"}, {"beta.amd.com/gpu.family.AI", "1"},
                             {"amd.com/cpu", "true"}, {"dummyLabel1", "1"}};
    labeller::removeOldNodeLabels(node);
    CHECK((node == labeller::Labels{{"amd.com/cpu", "true"}, {"dummyLabel1", "1"}}));
    CHECK((labeller::createLabels("vram", {{"16G", 2}}) ==
           labeller::Labels{{"beta.amd.com/gpu.vram.16G", "2"}, {"beta.amd.com/gpu.vram", "16G"}, {"amd.com/gpu.vram", "16G"}}));
    CHECK((labeller::createLabels("vram", {{"16G", 2}, {"64G", 1}}) ==
           labeller::Labels{{"beta.amd.com/gpu.vram.16G", "2"}, {"beta.amd.com/gpu.vram.64G", "1"}, {"amd.com/gpu.vram.16G", "2"}, {"amd.com/gpu.vram.64G", "1"}}));
}

// ---- plugin flow on a sysroot (no reference test exists: needs a live node there) ---------------------------------------
static std::string g_sysroot;
static void TestPluginFlowOnSysroot() {
    auto oc = Context::Open("kfd:" + g_sysroot);
    CHECK(!oc.second);
    if (oc.second) return;
    Ctx ctx = oc.first;
    CHECK(bool(Context::Open("kfd:" + g_sysroot + "/nope").second));  // glog.Fatalf in the reference (amdgpu.go:150-152)
    auto gpus = amdgpu::GetAMDGPUs(ctx);
    CHECK(gpus.size() == 63 && amdgpu::IsHomogeneous(ctx));
    CHECK((amdgpu::UniquePartitionConfigCount(ctx) == std::map<std::string, int>{{"cpx_nps4", 63}}));
    auto rl = plugin::getResourceList(ctx, "single");
    CHECK(!rl.second && rl.first == std::vector<std::string>{"gpu"});
    CHECK(bool(plugin::getResourceList(ctx, "bogus").second));

    plugin::AMDGPULister lister(ctx);
    CHECK(lister.GetResourceNamespace() == "amd.com");
    auto p = lister.NewPlugin("gpu");
    CHECK(!p->Start() && !p->allocatorInitError && p->GetDevicePluginOptions().GetPreferredAllocationAvailable);

    std::vector<std::string> sent;
    int ticks = 2;  // two heartbeats, then the stop signal
    Error e = p->ListAndWatch([&](const std::string& wire) { sent.push_back(wire); }, [&] { return ticks-- > 0; }, B2DP_LW_NO_PROBE);
    CHECK(!e && sent.size() == 3 && sent[0] == sent[1] && sent[1] == sent[2] && p->last_stats.n_devices == 63);

    std::vector<std::string> ids;
    for (auto& kv : gpus) ids.push_back(kv.first);
    auto pa = p->GetPreferredAllocation({{ids, {}, 3}, {ids, {ids[5]}, 2}});
    CHECK(!pa.second && pa.first.size() == 2 && pa.first[0].size() == 3 && pa.first[1].size() == 2);
    CHECK(std::find(pa.first[1].begin(), pa.first[1].end(), ids[5]) != pa.first[1].end());
    auto bad = p->GetPreferredAllocation({{{ids[0], ids[1]}, {}, 3}});
    CHECK(bad.second.what() == "unable to get preferred allocation list. Error:available devices count less than allocation size");

    auto al = p->Allocate({{ids[0], "bogus"}, {}});
    CHECK(!al.second && al.first.size() == 2);
    CHECK(al.first[0].Devices.size() == 3 && al.first[0].Devices[0].HostPath == "/dev/kfd" && al.first[0].Devices[0].Permissions == "rw");
    CHECK(al.first[0].Devices[1].HostPath == "/dev/dri/card" + std::to_string(gpus[ids[0]].card));
    CHECK(al.first[0].Devices[2].HostPath == "/dev/dri/renderD" + std::to_string(gpus[ids[0]].renderD));
    CHECK(al.first[1].Devices.size() == 1);

    std::vector<exporter::PluginDevice> devs = {{"a", "", 0}, {"b", "", 0}, {"c", "", 0}};
    const std::map<std::string, std::string> hmap = {{"b", exporter::Unhealthy}, {"zz", exporter::Unhealthy}};
    exporter::PopulatePerGPUDHealth(devs, exporter::Healthy, &hmap);
    CHECK(devs[0].Health == "Healthy" && devs[1].Health == "Unhealthy" && devs[2].Health == "Healthy");
    exporter::PopulatePerGPUDHealth(devs, exporter::Unhealthy, nullptr);  // exporter absent: default for everyone
    CHECK(devs[0].Health == "Unhealthy" && devs[1].Health == "Unhealthy");

    auto labels = labeller::generateLabels(ctx, {"cu-count", "simd-count", "compute-memory-partition"});
    CHECK(labels.count("amd.com/gpu.cu-count") && labels["amd.com/gpu.compute-memory-partition"] == "cpx_nps4");
}

// ---- the same plugin flow on real B200s (`host_mirror_test --cuda <uri>`, run by the -m gpu suite) -------------------
static std::string g_uri;
static void TestPluginFlowOnCuda() {
    auto oc = Context::Open(g_uri);
    CHECK(!oc.second);
    if (oc.second) { fprintf(stderr, "open: %s\n", oc.second.what().c_str()); return; }
    Ctx ctx = oc.first;
    auto gpus = amdgpu::GetAMDGPUs(ctx);
    CHECK(!gpus.empty() && amdgpu::IsHomogeneous(ctx));
    plugin::AMDGPULister lister(ctx);
    auto p = lister.NewPlugin("gpu");
    CHECK(!p->Start());
    CHECK(p->allocatorInitError == (gpus.size() < 2));  // one GPU: no pair weights, kubelet default allocation
    std::vector<std::string> sent;
    int ticks = 3;
    Error e = p->ListAndWatch([&](const std::string& wire) { sent.push_back(wire); }, [&] { return ticks-- > 0; });
    CHECK(!e && sent.size() == 4 && sent[0] == sent[3]);                        // every heartbeat: all Healthy again
    CHECK(p->last_stats.n_devices == (int)gpus.size() && p->last_stats.n_unhealthy == 0 && p->last_stats.node_healthy);
    CHECK(p->last_stats.probe_gbs_min > 1000.f && p->last_stats.probe_bytes > 0);  // the HBM pass ran on every GPU
    std::vector<std::string> ids;
    for (auto& kv : gpus) ids.push_back(kv.first);
    auto al = p->Allocate({{ids[0]}});
    CHECK(!al.second && al.first[0].Devices.size() == 4 && al.first[0].Devices[0].HostPath == "/dev/nvidiactl");
    CHECK(al.first[0].Devices[3].HostPath == "/dev/nvidia" + std::to_string(gpus[ids[0]].card));
    if (gpus.size() >= 2) {
        auto pa = p->GetPreferredAllocation({{ids, {}, 2}});
        CHECK(!pa.second && pa.first[0].size() == 2);
    }
    auto labels = labeller::generateLabels(ctx, {"cu-count", "product-name"});
    CHECK(labels["amd.com/gpu.cu-count"] == "148");
    printf("cuda flow: %zu GPUs, min %.0f GB/s per GPU inside the heartbeat, cycle %.3f ms\n", gpus.size(),
           p->last_stats.probe_gbs_min, p->last_stats.ms_total);
}

int main(int argc, char** argv) {
    if (argc == 3 && std::string(argv[1]) == "--cuda") {
        g_uri = argv[2];
        RUN(TestPluginFlowOnCuda);
        printf("%s: %d checks, %d failed\n", g_failed ? "FAIL" : "PASS", g_checks, g_failed);
        return g_failed ? 1 : 0;
    }
    if (argc < 2) { fprintf(stderr, "usage: host_mirror_test <fixtures_dir> [<cpx sysroot>]\n"); return 2; }
    testdata = argv[1];
    RUN(TestParseTopologyProperties);
    RUN(TestParseDebugFSFirmwareInfo);
    RUN(TestRenderDevIdsFromTopology);
    RUN(TestCountGPUDevFromTopology);
    RUN(TestPairWeightCalculation);
    RUN(TestGroupPartitionsByDevId);
    RUN(TestGetSubsetsMethod);
    RUN(TestBestPolicyAllocator);
    RUN(TestRemoveOldNodeLabels);
    if (argc > 2) { g_sysroot = argv[2]; RUN(TestPluginFlowOnSysroot); }
    printf("%s: %d checks, %d failed\n", g_failed ? "FAIL" : "PASS", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
