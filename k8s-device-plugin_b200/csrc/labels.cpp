// labels.cpp -- node-labeller label arithmetic (cmd/k8s-node-labeller/main.go:37-397).
//
// The label *scheme* (dual prefix, counter labels, value formatting, clean-up of old labels)
// is the reference's, key for key.  The label *sources* are per backend: the kfd backend
// reads the same sysfs/kfd files as the reference's generators; the cuda backend answers
// from CUDA/NVML queries (vram, SM count, product name, driver version, ...).
#include <cmath>
#include <cstring>

#include "gosem.hpp"
#include "internal.hpp"

namespace b2dp {

// main.go:38-39: amdPrefix = "amd.com", experimentalAMDPrefix = "beta.amd.com".  The domain is
// a process-wide setting (b2dp_set_vendor_domain) so a deployment can publish e.g. nvidia.com/gpu.*
// labels; the default keeps the reference's keys.
static std::mutex g_domain_mu;
static std::string g_domain = "amd.com";
static std::string domain() { std::lock_guard<std::mutex> g(g_domain_mu); return g_domain; }

// sorted; main.go:115-379 map keys
const char* const kGeneratorNames[12] = {
    "compute-memory-partition", "compute-partitioning-supported", "cu-count", "device-id", "driver-src-version",
    "driver-version", "family", "firmware", "memory-partitioning-supported", "product-name", "simd-count", "vram"};

static std::string label_prefix(const std::string& name, bool experimental) {  // main.go:76-85
    return (experimental ? "beta." + domain() : domain()) + "/gpu." + name;
}

void create_labels(const std::string& kind, const std::map<std::string, int>& entries,
                   std::map<std::string, std::string>& out) {  // main.go:87-108
    std::string pfx = label_prefix(kind, true);
    for (const auto& kv : entries) {
        out[pfx + "." + kv.first] = std::to_string(kv.second);
        if (entries.size() == 1) out[pfx] = kv.first;
    }
    pfx = label_prefix(kind, false);
    for (const auto& kv : entries) {
        if (entries.size() == 1) out[pfx] = kv.first;
        else out[pfx + "." + kv.first] = std::to_string(kv.second);
    }
}

// Extension generator (no reference counterpart, BASELINE.json "labeller's link labels"): the
// interconnect class of each GPU from the measured P2P matrix; same label scheme, and cleaned up
// together with the reference's twelve.
static const char kP2pLink[] = "p2p-link";

void remove_old_node_labels(std::map<std::string, std::string>& labels) {  // main.go:55-74
    std::vector<std::string> names(kGeneratorNames, kGeneratorNames + 12);
    names.push_back(kP2pLink);
    for (const auto& g : names) labels.erase(label_prefix(g, false));
    for (const auto& g : names) {
        const std::string k = label_prefix(g, true);
        auto it = labels.find(k);
        if (it != labels.end()) {
            const std::string counter = k + "." + it->second;
            labels.erase(it);
            labels.erase(counter);
        }
    }
}

static std::string vram_value(int64_t size_in_bytes) {  // main.go:268-272
    const int64_t tmp = size_in_bytes / (1024 * 1024);
    const long s = std::lround((double)tmp / 1024.0);  // math.Round: half away from zero
    return std::to_string(s) + "G";
}

static std::string replace_product(std::string s) {  // main.go:211: " "->"_", "(" and ")" removed
    std::string o;
    for (char c : s) {
        if (c == ' ') o.push_back('_');
        else if (c == '(' || c == ')') continue;
        else o.push_back(c);
    }
    return o;
}

// a version string as part of a label key: Kubernetes allows [A-Za-z0-9_.-] there
static std::string label_safe(const std::string& v) {
    std::string o;
    for (char ch : go::trim_space(v)) o.push_back(isalnum((unsigned char)ch) || ch == '.' || ch == '-' || ch == '_' ? ch : '_');
    return o;
}

static bool enabled(const std::string& csv, const char* name) {
    size_t pos = 0;
    const size_t len = strlen(name);
    while (pos <= csv.size()) {
        size_t e = csv.find(',', pos);
        if (e == std::string::npos) e = csv.size();
        if (e - pos == len && csv.compare(pos, len, name) == 0) return true;
        pos = e + 1;
    }
    return false;
}

int generate_labels(const std::vector<Device>& devs, const LabelSource& src, const std::string& csv,
                    std::map<std::string, std::string>& out) {
    const std::string root = src.sysroot.empty() ? "/" : src.sysroot;
    auto first_readable = [&](const char* leaf) {  // main.go:160-170: first card whose file reads
        std::string version, data;
        for (const auto& d : devs) {
            if (!go::read_file(go::join(root, "sys/class/drm/card" + std::to_string(d.card) + leaf), data)) continue;
            version = go::trim_space(data);
            break;
        }
        return version;
    };
    for (const char* g : kGeneratorNames) {
        if (!enabled(csv, g)) continue;
        const std::string name = g;
        std::map<std::string, int> counts;
        if (name == "firmware") {
            // main.go:116-144: counts["<block>.feat.<n>"] / counts["<block>.fw.<n>"] per card, experimental prefix only.
            // The reference asks libdrm (ioctl; no file form).  cuda backend: NVML's versioned blocks (VBIOS, InfoROM
            // image / OEM / ECC / power objects, GSP firmware) as "<block>.fw.<version>"; kfd backend: the per-card
            // side file b2dp_export_kfd_tree() writes ("<block> <version>" lines), absent on a real amdgpu sysfs.
            for (size_t i = 0; i < devs.size(); ++i) {
                if (src.native) {
                    if (i >= src.firmware.size()) continue;
                    for (const auto& bv : src.firmware[i]) counts[bv.first + ".fw." + label_safe(bv.second)]++;
                } else {
                    std::string data;
                    if (!go::read_file(go::join(root, "sys/class/drm/card" + std::to_string(devs[i].card) + "/device/b2dp_firmware"), data))
                        continue;
                    size_t pos = 0;
                    while (pos < data.size()) {
                        size_t e = data.find('\n', pos);
                        if (e == std::string::npos) e = data.size();
                        const std::string line = data.substr(pos, e - pos);
                        pos = e + 1;
                        const size_t sp = line.find(' ');
                        if (sp == std::string::npos || sp == 0 || sp + 1 >= line.size()) continue;
                        counts[line.substr(0, sp) + ".fw." + label_safe(line.substr(sp + 1))]++;
                    }
                }
            }
            const std::string pfx = label_prefix("firmware", true);
            for (auto& kv : counts) out[pfx + "." + kv.first] = std::to_string(kv.second);
        } else if (name == "family") {  // main.go:145-158; libdrm on the reference side, same two sources as above
            for (size_t i = 0; i < devs.size(); ++i) {
                std::string fam;
                if (src.native) { if (i < src.family.size()) fam = src.family[i]; }
                else {
                    std::string data;
                    if (go::read_file(go::join(root, "sys/class/drm/card" + std::to_string(devs[i].card) + "/device/b2dp_family"), data))
                        fam = go::trim_space(data);
                }
                if (!fam.empty()) counts[fam]++;
            }
            create_labels("family", counts, out);
        } else if (name == "driver-version") {  // main.go:159-174
            out[label_prefix(name, false)] =
                src.native ? src.driver_version : first_readable("/device/driver/module/version");
        } else if (name == "driver-src-version") {  // main.go:175-190
            out[label_prefix(name, false)] =
                src.native ? src.driver_src_version : first_readable("/device/driver/module/srcversion");
        } else if (name == "device-id") {  // main.go:191-209
            for (size_t i = 0; i < devs.size(); ++i) {
                std::string devid;
                if (src.native) {
                    if (i >= src.device_id.size()) continue;
                    devid = src.device_id[i];
                } else {
                    std::string data;
                    if (!go::read_file(go::join(root, "sys/class/drm/card" + std::to_string(devs[i].card) + "/device/device"), data))
                        continue;
                    devid = go::trim_space(data);
                }
                if (devid.size() < 2) return B2DP_E_PANIC;  // devid[0:2], main.go:200
                if (devid.compare(0, 2, "0x") == 0) devid = devid.substr(2);
                counts[devid]++;
            }
            create_labels("device-id", counts, out);
        } else if (name == "product-name") {  // main.go:210-238
            for (size_t i = 0; i < devs.size(); ++i) {
                std::string prod;
                if (src.native) {
                    if (i < src.product_name.size()) prod = replace_product(go::trim_space(src.product_name[i]));
                } else {
                    std::string data;
                    if (go::read_file(go::join(root, "sys/class/drm/card" + std::to_string(devs[i].card) + "/device/product_name"), data))
                        prod = replace_product(go::trim_space(data));
                }
                if (prod.empty()) continue;
                counts[prod]++;
            }
            create_labels("product-name", counts, out);
        } else if (name == "vram" || name == "simd-count" || name == "cu-count") {  // main.go:239-354
            if (src.native) {
                for (size_t i = 0; i < devs.size(); ++i) {
                    if (name == "vram") { if (i < src.vram_bytes.size()) counts[vram_value(src.vram_bytes[i])]++; }
                    else if (i < src.sm_count.size())
                        counts[std::to_string(name == "simd-count" ? src.sm_count[i] * 4 : src.sm_count[i])]++;
                }
            } else {
                const std::string kfd = go::join(root, "sys/class/kfd/kfd");
                auto files = go::glob_node_properties(kfd);
                if (files.empty()) continue;  // main.go:247-250: returns an empty map
                // read every node file once: render minor, simd_count, simd_per_cu
                struct NodeInfo { std::string node; int64_t minor; int rc_simd, rc_per; int64_t simd, per; };
                std::vector<NodeInfo> nodes;
                for (auto& f : files) {
                    NodeInfo ni{};
                    ni.node = go::base(go::dir(f));
                    kfd_parse_property(f, "drm_render_minor", &ni.minor);  // error ignored, main.go:256
                    ni.rc_simd = kfd_parse_property(f, "simd_count", &ni.simd);
                    ni.rc_per = kfd_parse_property(f, "simd_per_cu", &ni.per);
                    nodes.push_back(ni);
                }
                for (const auto& d : devs) {
                    for (const auto& ni : nodes) {
                        if ((int)ni.minor != d.render_d) continue;
                        std::string key;
                        if (name == "vram") {
                            int64_t vsize;
                            if (kfd_parse_property(kfd + "/topology/nodes/" + ni.node + "/mem_banks/0/properties",
                                                   "size_in_bytes", &vsize) != B2DP_OK) continue;
                            key = vram_value(vsize);
                        } else if (name == "simd-count") {
                            if (ni.rc_simd != B2DP_OK) continue;
                            key = std::to_string(ni.simd);
                        } else {
                            if (ni.rc_simd != B2DP_OK) continue;
                            if (ni.rc_per != B2DP_OK || ni.per == 0) continue;
                            key = std::to_string(ni.simd / ni.per);
                        }
                        counts[key]++;
                        break;
                    }
                }
            }
            create_labels(name, counts, out);
        } else if (name == "compute-memory-partition") {  // main.go:355-368
            auto hist = partition_histogram(devs);
            if (hist.size() <= 1)
                for (auto& kv : hist)
                    if (kv.second > 0) { out[label_prefix(name, false)] = kv.first; break; }
        } else if (name == "compute-partitioning-supported" || name == "memory-partitioning-supported") {
            const int which = name[0] == 'c' ? 0 : 1;  // main.go:369-378
            const bool v = src.native ? src.part_supported[which] : kfd_partition_supported(root, which);
            out[label_prefix(name, false)] = v ? "true" : "false";
        }
    }
    if (enabled(csv, kP2pLink) && src.native && !src.p2p_class.empty()) {
        std::map<std::string, int> counts;
        for (const auto& c : src.p2p_class) counts[c]++;
        create_labels(kP2pLink, counts, out);
    }
    return B2DP_OK;
}

}  // namespace b2dp

// ================================ C ABI ====================================================
using namespace b2dp;

static int emit_labels(const std::map<std::string, std::string>& m, b2dp_label* out, int cap, int* n) {
    *n = (int)m.size();
    if (*n > cap) return B2DP_E_NOSPC;
    if (*n && !out) return B2DP_E_INVAL;
    int i = 0;
    for (auto& kv : m) {
        memset(&out[i], 0, sizeof out[i]);
        copy_str(out[i].key, sizeof out[i].key, kv.first);
        copy_str(out[i].value, sizeof out[i].value, kv.second);
        ++i;
    }
    return B2DP_OK;
}

extern "C" int b2dp_create_labels(const char* kind, const b2dp_kv_count* entries, int n_entries, b2dp_label* out,
                                  int cap, int* n) {
    if (!kind || !n || n_entries < 0 || (n_entries && !entries) || cap < 0) return B2DP_E_INVAL;
    std::map<std::string, int> e;
    for (int i = 0; i < n_entries; ++i) e[entries[i].key] = entries[i].count;
    std::map<std::string, std::string> m;
    create_labels(kind, e, m);
    return emit_labels(m, out, cap, n);
}

extern "C" int b2dp_set_vendor_domain(const char* d) {
    if (!d || !*d || strlen(d) > 63 || strchr(d, '/')) return B2DP_E_INVAL;
    std::lock_guard<std::mutex> g(g_domain_mu);
    g_domain = d;
    return B2DP_OK;
}

extern "C" int b2dp_label_generator_names(char (*names)[64], int cap, int* n) {
    if (!n || cap < 0) return B2DP_E_INVAL;
    *n = 12;
    if (cap < 12) return B2DP_E_NOSPC;
    if (!names) return B2DP_E_INVAL;
    for (int i = 0; i < 12; ++i) copy_str(names[i], 64, kGeneratorNames[i]);
    return B2DP_OK;
}

extern "C" int b2dp_remove_old_node_labels(b2dp_label* labels, int n_in, int* n) {
    if (!n || n_in < 0 || (n_in && !labels)) return B2DP_E_INVAL;
    std::map<std::string, std::string> m;
    for (int i = 0; i < n_in; ++i) m[labels[i].key] = labels[i].value;
    remove_old_node_labels(m);
    // keep the caller's order for survivors
    int k = 0;
    for (int i = 0; i < n_in; ++i) {
        auto it = m.find(labels[i].key);
        if (it == m.end()) continue;
        if (k != i) labels[k] = labels[i];
        m.erase(it);  // duplicates of a key keep the first
        ++k;
    }
    *n = k;
    return B2DP_OK;
}

int b2dp_emit_labels_internal(const std::map<std::string, std::string>& m, b2dp_label* out, int cap, int* n) {
    return emit_labels(m, out, cap, n);
}
