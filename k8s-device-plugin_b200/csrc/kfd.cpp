// kfd.cpp -- sysfs / kfd-topology readers: the `kfd:` (parity) backend.
//
// Same results as internal/pkg/amdgpu/amdgpu.go and the helpers in
// internal/pkg/plugin/plugin.go, built differently: every properties file is read once
// and all keys are extracted in one pass over its lines (the reference re-opens and
// regex-scans a GPU node file 4x per enumeration).
#include <cstring>

#include "gosem.hpp"
#include "internal.hpp"

namespace b2dp {

// First-match-wins lookup of several `<key>\s(\d+)` patterns in one pass
// (ParseTopologyProperties semantics per key, amdgpu.go:442-463).
struct KeyHit {
    const char* key;
    bool found = false;
    go::NumErr err = go::NumErr::none;
    int64_t value = 0;
};
static void scan_first_match(std::string_view data, KeyHit* hits, int n_hits) {
    int remaining = n_hits;
    go::scan_lines(data, [&](std::string_view line) {
        for (int i = 0; i < n_hits; ++i) {
            if (hits[i].found) continue;
            std::string_view digits;
            if (go::match_key_digits(line, hits[i].key, digits)) {
                hits[i].found = true;
                hits[i].err = go::parse_int(digits, 0, 64, &hits[i].value);
                --remaining;
            }
        }
        return remaining > 0;
    });
}

int kfd_parse_property(const std::string& path, const char* key, int64_t* value) {
    std::string data;
    *value = 0;
    if (!go::read_attr(path, data)) return B2DP_E_IO;
    KeyHit h{key};
    scan_first_match(data, &h, 1);
    if (!h.found) return B2DP_E_NOTFOUND;
    *value = h.value;
    if (h.err == go::NumErr::syntax) return B2DP_E_SYNTAX;
    if (h.err == go::NumErr::range) return B2DP_E_RANGE;
    return B2DP_OK;
}

void kfd_topology_maps(const std::string& topo_root, TopoMaps& out) {
    out.dev_ids.clear();
    out.node_ids.clear();
    std::string data;
    for (const auto& node_file : go::glob_node_properties(topo_root)) {
        if (!go::read_attr(node_file, data)) continue;
        KeyHit h[3] = {{"drm_render_minor"}, {"location_id"}, {"domain"}};
        scan_first_match(data, h, 3);
        if (!h[0].found || h[0].err != go::NumErr::none) continue;  // amdgpu.go:118-122,513-517
        if (h[0].value <= 0) continue;                              // amdgpu.go:124,519
        const int minor = (int)h[0].value;
        // amdgpu.go:525-534: node id = directory name through Atoi
        int64_t node_id;
        if (go::atoi(go::base(go::dir(node_file)), &node_id) == go::NumErr::none) out.node_ids[minor] = (int)node_id;
        // amdgpu.go:128-142
        if (!h[1].found || h[1].err != go::NumErr::none) continue;
        if (!h[2].found || h[2].err != go::NumErr::none) continue;
        const int64_t loc = h[1].value, domain = h[2].value;
        char buf[64];
        // fmt.Sprintf("%04x:%02x:%02x:0", domain, bus, dev) with int64 operands
        snprintf(buf, sizeof buf, "%04llx:%02llx:%02llx:0", (unsigned long long)domain,
                 (unsigned long long)((loc >> 8) & 0xff), (unsigned long long)((loc >> 3) & 0x1f));
        out.dev_ids[minor] = buf;
    }
}

static bool read_trim_lower(const std::string& path, std::string& out) {
    std::string data;
    if (!go::read_attr(path, data)) return false;
    out = go::to_lower(go::trim_space(data));
    return true;
}

// The drm/* loop shared by both halves of GetAMDGPUs (amdgpu.go:200-214, 230-254).
// Returns B2DP_E_PANIC where Go's name[0:4] / name[0:7] slicing would panic.
struct Sticky { int card = 0, render_d = 128, node_id = 0; std::string dev_id; };

int kfd_enumerate(const std::string& sysroot, std::vector<Device>& out, std::string& err) {
    out.clear();
    const std::string root = sysroot.empty() ? "/" : sysroot;
    if (!go::exists(go::join(root, "sys/module/amdgpu/drivers/"))) {
        err = "amdgpu driver unavailable (reference: glog.Fatalf, exit code 2)";
        return B2DP_E_NODRIVER;
    }
    TopoMaps maps;
    kfd_topology_maps(go::join(root, "sys/class/kfd/kfd"), maps);

    std::map<std::string, Device> devices;
    Sticky st;  // declared outside the loops, exactly like amdgpu.go:157-159
    for (const auto& path : go::glob_pci_bdf(go::join(root, "sys/module/amdgpu/drivers/pci:amdgpu"))) {
        Device d;
        read_trim_lower(path + "/current_compute_partition", d.compute);
        read_trim_lower(path + "/current_memory_partition", d.memory);
        std::string numa_raw;
        if (!go::read_file(path + "/numa_node", numa_raw)) continue;  // amdgpu.go:192-195
        int64_t numa;
        if (go::atoi(go::trim_space(numa_raw), &numa) != go::NumErr::none) continue;  // amdgpu.go:187-191
        d.numa = (int)numa;
        for (const auto& dev_path : go::glob_all(path + "/drm")) {
            const std::string name = go::base(dev_path);
            if (name.size() < 4) { err = "drm entry '" + name + "': reference panics on name[0:4]"; return B2DP_E_PANIC; }
            if (name.compare(0, 4, "card") == 0) {
                st.card = (int)go::atoi_ignore_err(std::string_view(name).substr(4));
            } else {
                if (name.size() < 7) { err = "drm entry '" + name + "': reference panics on name[0:7]"; return B2DP_E_PANIC; }
                if (name.compare(0, 7, "renderD") == 0) {
                    st.render_d = (int)go::atoi_ignore_err(std::string_view(name).substr(7));
                    auto di = maps.dev_ids.find(st.render_d);
                    if (di != maps.dev_ids.end()) st.dev_id = di->second;
                    auto ni = maps.node_ids.find(st.render_d);
                    if (ni != maps.node_ids.end()) st.node_id = ni->second;
                }
            }
        }
        d.id = go::base(path);
        d.card = st.card; d.render_d = st.render_d; d.dev_id = st.dev_id; d.node_id = st.node_id;
        devices[d.id] = d;
    }

    for (const auto& path : go::glob_prefixed(go::join(root, "sys/devices/platform"), "amdgpu_xcp_")) {
        Device d;
        d.numa = -1;
        for (const auto& dev_path : go::glob_all(path + "/drm")) {
            const std::string name = go::base(dev_path);
            if (name.size() < 4) { err = "drm entry '" + name + "': reference panics on name[0:4]"; return B2DP_E_PANIC; }
            if (name.compare(0, 4, "card") == 0) {
                st.card = (int)go::atoi_ignore_err(std::string_view(name).substr(4));
            } else {
                if (name.size() < 7) { err = "drm entry '" + name + "': reference panics on name[0:7]"; return B2DP_E_PANIC; }
                if (name.compare(0, 7, "renderD") == 0) {
                    st.render_d = (int)go::atoi_ignore_err(std::string_view(name).substr(7));
                    auto di = maps.dev_ids.find(st.render_d);
                    if (di != maps.dev_ids.end()) st.dev_id = di->second;
                    // amdgpu.go:240-249 (Go: random map order, first hit; here: sorted by id)
                    for (const auto& kv : devices) {
                        const Device& o = kv.second;
                        if (o.dev_id == st.dev_id && !o.compute.empty() && !o.memory.empty()) {
                            d.compute = o.compute; d.memory = o.memory; d.numa = o.numa;
                            break;
                        }
                    }
                    auto ni = maps.node_ids.find(st.render_d);
                    if (ni != maps.node_ids.end()) st.node_id = ni->second;
                }
            }
        }
        if (maps.dev_ids.find(st.render_d) == maps.dev_ids.end()) continue;  // amdgpu.go:258-260
        if (d.numa == -1) continue;                                           // amdgpu.go:261-263
        d.id = go::base(path);
        d.card = st.card; d.render_d = st.render_d; d.dev_id = st.dev_id; d.node_id = st.node_id;
        devices[d.id] = d;
    }
    out.reserve(devices.size());
    for (auto& kv : devices) out.push_back(std::move(kv.second));
    return B2DP_OK;
}

std::map<std::string, int> partition_histogram(const std::vector<Device>& devs) {
    std::map<std::string, int> m;  // amdgpu.go:270-285
    for (const auto& d : devs)
        if (!d.compute.empty() && !d.memory.empty()) m[d.compute + "_" + d.memory]++;
    return m;
}

bool kfd_partition_supported(const std::string& sysroot, int which) {
    // amdgpu.go:295-328: first PCI dir only
    auto matches = go::glob_pci_bdf(go::join(sysroot.empty() ? "/" : sysroot, "sys/module/amdgpu/drivers/pci:amdgpu"));
    if (matches.empty()) return false;
    return go::exists(matches[0] + (which == 0 ? "/available_compute_partition" : "/available_memory_partition"));
}

int resource_list(const std::vector<Device>& devs, const char* strategy, std::vector<std::string>& out) {
    out.clear();
    const bool single = strategy && strcmp(strategy, "single") == 0;
    const bool mixed = strategy && strcmp(strategy, "mixed") == 0;
    if (!single && !mixed) return B2DP_E_INVAL;  // main.go:42-51
    auto counts = partition_histogram(devs);
    const bool homogeneous = counts.size() <= 1;
    if (devs.empty()) return B2DP_OK;            // main.go:59-61
    if (homogeneous) {
        if (single || counts.empty()) { out.push_back("gpu"); return B2DP_OK; }
        for (auto& kv : counts) if (kv.second > 0) out.push_back(kv.first);
        return B2DP_OK;
    }
    if (single) return B2DP_E_HETEROGENEOUS;     // main.go:78-80
    for (auto& kv : counts) if (kv.second > 0) out.push_back(kv.first);
    return B2DP_OK;
}

int kfd_count_gpu_dev(const std::string& topo_root) {
    int count = 0;  // plugin.go:123-159
    std::string data;
    for (const auto& f : go::glob_node_properties(topo_root)) {
        if (!go::read_attr(f, data)) continue;
        go::scan_lines(data, [&](std::string_view line) {
            std::string_view digits;
            if (!go::match_key_digits(line, "simd_count", digits)) return true;
            if (go::atoi_ignore_err(digits) > 0) { ++count; return false; }
            return true;
        });
    }
    return count;
}

bool kfd_simple_health_check(const std::string& topo_root) {
    std::string data;  // plugin.go:161-206
    bool found = false;
    go::for_each_node_properties(topo_root, [&](const std::string& f) {
        if (!go::read_attr(f, data)) return true;
        int64_t cpu_cores = 0, gfx = 0;
        auto starts = [](std::string_view l, const char* p) { return l.compare(0, strlen(p), p) == 0; };
        const bool scan_err = go::scan_lines(data, [&](std::string_view line) {
            if (starts(line, "cpu_cores_count")) {
                auto parts = go::fields(line);
                if (parts.size() == 2) cpu_cores = go::atoi_ignore_err(parts[1]);
            } else if (starts(line, "gfx_target_version")) {
                auto parts = go::fields(line);
                if (parts.size() == 2) gfx = go::atoi_ignore_err(parts[1]);
            }
            return true;
        });
        if (scan_err) return true;  // plugin.go:193-196
        if (cpu_cores == 0 && gfx > 0) { found = true; return false; }  // first hit wins
        return true;
    });
    return found;
}

}  // namespace b2dp

// =========================== C ABI: stateless readers ====================================
using namespace b2dp;

extern "C" int b2dp_parse_topology_property(const char* path, const char* key, int64_t* value) {
    if (!path || !key || !value || !*key) return B2DP_E_INVAL;
    return kfd_parse_property(path, key, value);
}

extern "C" int b2dp_dev_ids_from_topology(const char* topo_root, int32_t* render_minor, char (*dev_id)[24], int cap,
                                          int* n) {
    if (!topo_root || !n || cap < 0) return B2DP_E_INVAL;
    TopoMaps m;
    kfd_topology_maps(topo_root, m);
    *n = (int)m.dev_ids.size();
    if (*n > cap) return B2DP_E_NOSPC;
    if (*n && (!render_minor || !dev_id)) return B2DP_E_INVAL;
    int i = 0;
    for (auto& kv : m.dev_ids) { render_minor[i] = kv.first; copy_str(dev_id[i], 24, kv.second); ++i; }
    return B2DP_OK;
}

extern "C" int b2dp_node_ids_from_topology(const char* topo_root, int32_t* render_minor, int32_t* node_id, int cap,
                                           int* n) {
    if (!topo_root || !n || cap < 0) return B2DP_E_INVAL;
    TopoMaps m;
    kfd_topology_maps(topo_root, m);
    *n = (int)m.node_ids.size();
    if (*n > cap) return B2DP_E_NOSPC;
    if (*n && (!render_minor || !node_id)) return B2DP_E_INVAL;
    int i = 0;
    for (auto& kv : m.node_ids) { render_minor[i] = kv.first; node_id[i] = kv.second; ++i; }
    return B2DP_OK;
}

extern "C" int b2dp_count_gpu_dev_from_topology(const char* topo_root, int32_t* count) {
    if (!topo_root || !count) return B2DP_E_INVAL;
    *count = kfd_count_gpu_dev(topo_root);
    return B2DP_OK;
}

extern "C" int b2dp_simple_health_check(const char* topo_root, int32_t* healthy) {
    if (!topo_root || !healthy) return B2DP_E_INVAL;
    *healthy = kfd_simple_health_check(topo_root) ? 1 : 0;
    return B2DP_OK;
}

extern "C" int b2dp_parse_debugfs_firmware_info(const char* path, b2dp_fw_entry* out, int cap, int* n) {
    if (!path || !n || cap < 0) return B2DP_E_INVAL;
    // amdgpu.go:465: `(\w+) feature version: (\d+), firmware version: (0x[0-9a-fA-F]+)`, unanchored
    static const char kMid[] = " feature version: ";
    static const char kFw[] = ", firmware version: 0x";
    std::map<std::string, std::pair<uint32_t, uint32_t>> m;
    std::string data;
    if (go::read_file(path, data)) {
        auto is_word = [](char c) { return go::is_digit(c) || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || c == '_'; };
        go::scan_lines(data, [&](std::string_view line) {
            // leftmost match: try every occurrence of the middle literal that has >= 1 \w before it
            size_t from = 0;
            for (;;) {
                size_t p = line.find(kMid, from);
                if (p == std::string_view::npos) break;
                from = p + 1;
                if (p == 0 || !is_word(line[p - 1])) continue;
                size_t q = p + sizeof(kMid) - 1, e = q;
                while (e < line.size() && go::is_digit(line[e])) ++e;
                if (e == q) continue;
                if (line.compare(e, sizeof(kFw) - 1, kFw) != 0) continue;
                size_t hx = e + sizeof(kFw) - 1, he = hx;
                while (he < line.size() && go::is_hex(line[he])) ++he;
                if (he == hx) continue;
                // (\w+) is greedy and leftmost: extend the name to the left as far as \w goes
                size_t nb = p;
                while (nb > 0 && is_word(line[nb - 1])) --nb;
                int64_t feat, fw;
                go::parse_int(line.substr(q, e - q), 0, 32, &feat);           // errors ignored, amdgpu.go:479
                go::parse_int(line.substr(hx - 2, he - (hx - 2)), 0, 32, &fw);  // amdgpu.go:481
                m[std::string(line.substr(nb, p - nb))] = {(uint32_t)feat, (uint32_t)fw};
                break;
            }
            return true;
        });
    }
    *n = (int)m.size();
    if (*n > cap) return B2DP_E_NOSPC;
    if (*n && !out) return B2DP_E_INVAL;
    int i = 0;
    for (auto& kv : m) {
        memset(&out[i], 0, sizeof out[i]);
        copy_str(out[i].name, sizeof out[i].name, kv.first);
        out[i].feature = kv.second.first; out[i].firmware = kv.second.second;
        ++i;
    }
    return B2DP_OK;
}
