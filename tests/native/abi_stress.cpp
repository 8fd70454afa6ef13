// Sanitizer harness over the public C ABI (include/b200dp.h), kfd: backend only.
//   abi_stress sweep <scratch_dir> <sysroot>...    every CPU-side entry point on each tree, incl. the
//                                                  B2DP_E_NOSPC / bad-argument paths (ASan + UBSan build)
//   abi_stress threads <sysroot> <T> <iters>       T threads hammer ONE shared context (enumerate,
//                                                  ListAndWatch, Start, GetPreferredAllocation, Allocate,
//                                                  labels) next to a native watch loop; every result must
//                                                  equal the single-threaded answer (TSan build)
// The reference has no sanitizer/race build at all (no `-race`, latent p.AMDGPUs race
// plugin.go:231 vs :375); this is the check that the library's "callable from any thread" contract holds.
#include <dirent.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200dp.h"

static int g_fail = 0;
#define EXPECT(c) do { if (!(c)) { fprintf(stderr, "EXPECT failed %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_fail; } } while (0)

static uint64_t fnv(uint64_t h, const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

struct Ids {
    std::vector<b2dp_device> devs;
    std::vector<const char*> ptr;
};

static int enumerate_all(b2dp_ctx* c, Ids& out) {
    int n = 0;
    b2dp_device one;
    int rc = b2dp_enumerate(c, &one, 0, &n);  // capacity 0: count only
    if (rc != B2DP_OK && rc != B2DP_E_NOSPC) return rc;
    out.devs.assign((size_t)n + 1, b2dp_device{});
    rc = b2dp_enumerate(c, out.devs.data(), (int)out.devs.size(), &n);
    if (rc != B2DP_OK) return rc;
    out.devs.resize((size_t)n);
    out.ptr.clear();
    for (auto& d : out.devs) out.ptr.push_back(d.id);
    return B2DP_OK;
}

// One deterministic pass over the context-level entry points; returns a digest of every output.
static uint64_t ctx_pass(b2dp_ctx* c, bool with_start) {
    uint64_t h = 1469598103934665603ull;
    Ids ids;
    int rc = enumerate_all(c, ids);
    h = fnv(h, &rc, sizeof rc);
    if (rc != B2DP_OK) return h;
    const int n = (int)ids.devs.size();
    for (auto& d : ids.devs) h = fnv(h, &d, sizeof d);

    b2dp_kv_count hist[64];
    int nh = 0;
    rc = b2dp_partition_histogram(c, hist, 64, &nh);
    h = fnv(h, &rc, sizeof rc);
    if (rc == B2DP_OK) for (int i = 0; i < nh; ++i) { h = fnv(h, hist[i].key, strlen(hist[i].key)); h = fnv(h, &hist[i].count, 4); }
    int32_t v = -1;
    rc = b2dp_is_homogeneous(c, &v); h = fnv(h, &rc, 4); h = fnv(h, &v, 4);
    for (int which = 0; which < 2; ++which) { rc = b2dp_partition_supported(c, which, &v); h = fnv(h, &rc, 4); h = fnv(h, &v, 4); }
    rc = b2dp_node_health(c, &v); h = fnv(h, &rc, 4); h = fnv(h, &v, 4);

    std::vector<std::string> resources;
    for (const char* strat : {"single", "mixed", "bogus"}) {
        char names[64][64];
        int nn = 0;
        rc = b2dp_resource_list(c, strat, names, 64, &nn);
        h = fnv(h, &rc, 4);
        if (rc == B2DP_OK) for (int i = 0; i < nn; ++i) { h = fnv(h, names[i], strlen(names[i])); resources.push_back(names[i]); }
    }
    resources.push_back("no_such_resource");

    std::vector<uint8_t> buf(1 << 16);
    for (auto& res : resources) {
        for (uint32_t flags : {B2DP_LW_INITIAL, B2DP_LW_HEARTBEAT | B2DP_LW_NO_PROBE, B2DP_LW_HEARTBEAT | B2DP_LW_EXTERNAL_SOURCE}) {
            b2dp_cycle_opts o{};
            o.flags = flags;
            std::vector<int32_t> src_h;
            if (flags & B2DP_LW_EXTERNAL_SOURCE) {  // every other device reported unhealthy by the "exporter"
                for (int i = 0; i < n; ++i) src_h.push_back(i & 1);
                o.src_ids = (const char (*)[64])nullptr;
                o.src_n = 0;
            }
            std::vector<char> src_ids((size_t)n * 64 + 64, 0);
            if (flags & B2DP_LW_EXTERNAL_SOURCE) {
                for (int i = 0; i < n; ++i) memcpy(&src_ids[(size_t)i * 64], ids.devs[(size_t)i].id, 64);
                o.src_ids = (const char (*)[64])src_ids.data();
                o.src_health = src_h.data();
                o.src_n = n;
            }
            size_t len = 0;
            b2dp_cycle_stats st{};
            rc = b2dp_list_and_watch(c, res.c_str(), &o, buf.data(), buf.size(), &len, &st);
            h = fnv(h, &rc, 4);
            if (rc == B2DP_OK) {
                h = fnv(h, buf.data(), len);
                h = fnv(h, &st.n_devices, 4); h = fnv(h, &st.n_unhealthy, 4); h = fnv(h, &st.homogeneous, 4);
                if (len > 1) {  // too-small buffer: the needed size comes back with B2DP_E_NOSPC
                    size_t need = 0;
                    const int rc2 = b2dp_list_and_watch(c, res.c_str(), &o, buf.data(), len - 1, &need, nullptr);
                    EXPECT(rc2 == B2DP_E_NOSPC && need == len);
                }
            }
        }
    }

    if (n) {
        std::vector<const char*> req(ids.ptr);
        req.push_back("unknown-device-id");
        b2dp_devspec specs[512];
        int ns = 0;
        rc = b2dp_device_specs(c, req.data(), (int)req.size(), specs, 512, &ns);
        h = fnv(h, &rc, 4);
        if (rc == B2DP_OK) for (int i = 0; i < ns; ++i) h = fnv(h, &specs[i], sizeof specs[i]);
        size_t len = 0;
        rc = b2dp_allocate_response(c, req.data(), (int)req.size(), buf.data(), buf.size(), &len);
        h = fnv(h, &rc, 4);
        if (rc == B2DP_OK) h = fnv(h, buf.data(), len);
    }

    if (with_start) (void)b2dp_start(c);   // not hashed: a pass with and without a re-Start must agree
    rc = b2dp_preferred_allocation_available(c, &v); h = fnv(h, &rc, 4); h = fnv(h, &v, 4);
    std::vector<char> out((size_t)(n + 2) * 64);
    for (int size = 0; size <= n + 1 && size <= 9; ++size) {
        int no = 0;
        rc = b2dp_preferred_allocation(c, ids.ptr.data(), n, nullptr, 0, size, (char (*)[64])out.data(), n + 2, &no);
        h = fnv(h, &rc, 4);
        if (rc == B2DP_OK) h = fnv(h, out.data(), (size_t)no * 64);
        if (n >= 2 && size >= 1) {  // must-include the last id; available minus the first
            const char* must[1] = {ids.ptr[(size_t)n - 1]};
            rc = b2dp_preferred_allocation(c, ids.ptr.data() + 1, n - 1, must, 1, size, (char (*)[64])out.data(), n + 2, &no);
            h = fnv(h, &rc, 4);
            if (rc == B2DP_OK) h = fnv(h, out.data(), (size_t)no * 64);
        }
    }

    char gens[16][64];
    int ng = 0;
    EXPECT(b2dp_label_generator_names(gens, 16, &ng) == B2DP_OK && ng == 12);
    std::string enabled;
    for (int i = 0; i < ng; ++i) { enabled += gens[i]; enabled += ","; }
    enabled += "p2p-link";
    std::vector<b2dp_label> labels(4096);
    int nl = 0;
    rc = b2dp_generate_labels(c, enabled.c_str(), labels.data(), (int)labels.size(), &nl);
    h = fnv(h, &rc, 4);
    if (rc == B2DP_OK) {
        for (int i = 0; i < nl; ++i) { h = fnv(h, labels[(size_t)i].key, strlen(labels[(size_t)i].key)); h = fnv(h, labels[(size_t)i].value, strlen(labels[(size_t)i].value)); }
        int kept = -1;
        // counter labels of a multi-valued generator ("amd.com/gpu.device-id.74a1") survive the clean-up,
        // exactly as in the reference (main.go:55-74 only deletes the bare keys and beta "<key>.<val>")
        std::vector<b2dp_label> tmp(labels.begin(), labels.begin() + nl);
        EXPECT(b2dp_remove_old_node_labels(tmp.data(), nl, &kept) == B2DP_OK && kept >= 0 && kept <= nl);
        for (int i = 0; i < kept; ++i) h = fnv(h, tmp[(size_t)i].key, strlen(tmp[(size_t)i].key));
        if (nl > 1) {
            int need = 0;
            EXPECT(b2dp_generate_labels(c, enabled.c_str(), labels.data(), nl - 1, &need) == B2DP_E_NOSPC && need == nl);
        }
    }
    (void)b2dp_last_error(c);
    return h;
}

static void stateless_pass(const std::string& sysroot) {
    const std::string topo = sysroot + "/sys/class/kfd/kfd";
    int32_t v = 0;
    (void)b2dp_count_gpu_dev_from_topology(topo.c_str(), &v);
    (void)b2dp_simple_health_check(topo.c_str(), &v);
    int32_t minors[512], nodes[512];
    char devids[512][24];
    int n = 0;
    int rc = b2dp_dev_ids_from_topology(topo.c_str(), minors, devids, 512, &n);
    if (rc == B2DP_OK && n > 1) { int need = 0; EXPECT(b2dp_dev_ids_from_topology(topo.c_str(), minors, devids, n - 1, &need) == B2DP_E_NOSPC && need == n); }
    rc = b2dp_node_ids_from_topology(topo.c_str(), minors, nodes, 512, &n);
    if (rc == B2DP_OK && n > 1) { int need = 0; EXPECT(b2dp_node_ids_from_topology(topo.c_str(), minors, nodes, n - 1, &need) == B2DP_E_NOSPC && need == n); }
    const std::string nodes_dir = topo + "/topology/nodes";
    if (DIR* d = opendir(nodes_dir.c_str())) {
        while (dirent* e = readdir(d)) {
            if (e->d_name[0] == '.') continue;
            const std::string f = nodes_dir + "/" + e->d_name + "/properties";
            for (const char* key : {"drm_render_minor", "location_id", "domain", "simd_count", "gfx_target_version", "unique_id", "no_such_key", ""}) {
                int64_t val = 0;
                (void)b2dp_parse_topology_property(f.c_str(), key, &val);
            }
        }
        closedir(d);
    }
    int64_t val = 0;
    EXPECT(b2dp_parse_topology_property((sysroot + "/does/not/exist").c_str(), "x", &val) == B2DP_E_IO);
    b2dp_fw_entry fw[64];
    (void)b2dp_parse_debugfs_firmware_info((sysroot + "/sys/kernel/debug/dri/0/amdgpu_firmware_info").c_str(), fw, 64, &n);
}

static void allocator_pass(b2dp_ctx* c, const std::string& sysroot) {
    Ids ids;
    if (enumerate_all(c, ids) != B2DP_OK) return;
    b2dp_allocator* a = nullptr;
    EXPECT(b2dp_allocator_new(&a) == B2DP_OK && a);
    const int n = (int)ids.devs.size();
    int rc = b2dp_allocator_init(a, ids.devs.data(), n, (sysroot + "/sys/class/kfd/kfd/topology/nodes").c_str());
    if (rc == B2DP_OK) {
        std::vector<b2dp_pair_weight> pw(8192);
        int np = 0, rows = 0;
        rc = b2dp_allocator_pair_weights(a, pw.data(), (int)pw.size(), &np, &rows);
        EXPECT(rc == B2DP_OK);
        if (np > 1) { int need = 0, r2 = 0; EXPECT(b2dp_allocator_pair_weights(a, pw.data(), np - 1, &need, &r2) == B2DP_E_NOSPC && need == np); }
        int32_t groups = 0;
        EXPECT(b2dp_allocator_group_count(a, &groups) == B2DP_OK);
        std::vector<char> out((size_t)(n + 1) * 64);
        for (int size = 1; size <= n && size <= 4; ++size) {
            int32_t nc = 0, best = 0;
            (void)b2dp_allocator_candidates(a, ids.ptr.data(), n, nullptr, 0, size, &nc, &best);
            int no = 0;
            (void)b2dp_allocator_allocate(a, ids.ptr.data(), n, nullptr, 0, size, (char (*)[64])out.data(), n + 1, &no);
        }
        // the same weights through the measured-links entry point
        std::vector<b2dp_link> links;
        for (int i = 0; i < np; ++i) links.push_back({pw[(size_t)i].node_from, pw[(size_t)i].node_to, 11});
        (void)b2dp_allocator_init_links(a, ids.devs.data(), n, links.data(), (int)links.size());
    }
    (void)b2dp_allocator_init(a, ids.devs.data(), 0, "");                       // empty device list
    (void)b2dp_allocator_init(a, ids.devs.data(), n, (sysroot + "/nope").c_str());  // no link files
    b2dp_allocator_free(a);
}

static int cmd_sweep(int argc, char** argv) {
    const std::string scratch = argv[2];
    EXPECT(b2dp_abi_version() == B2DP_ABI_VERSION);
    for (int code = 1; code > -40; --code) EXPECT(b2dp_strerror(code) != nullptr);
    b2dp_kv_count kv[3] = {{"a", 1}, {"b", 2}, {"a-very-long-key-that-is-still-legal-0123456789-0123456789-0123", 300}};
    b2dp_label lab[64];
    int nlab = 0;
    EXPECT(b2dp_create_labels("vram", kv, 3, lab, 64, &nlab) == B2DP_OK);
    { int need = 0; EXPECT(b2dp_create_labels("vram", kv, 3, lab, 1, &need) == B2DP_E_NOSPC && need == nlab); }
    char ids[4][64] = {"a", "b", "c", "d"}, src[2][64] = {"b", "zz"};
    int32_t sh[2] = {0, 0}, outh[4];
    EXPECT(b2dp_merge_health(ids, 4, 1, 1, src, sh, 2, outh) == B2DP_OK && outh[0] == 1 && outh[1] == 0);
    EXPECT(b2dp_merge_health(ids, 4, 0, 0, nullptr, nullptr, 0, outh) == B2DP_OK && outh[3] == 0);

    uint64_t digest = 0;
    for (int i = 3; i < argc; ++i) {
        const std::string sysroot = argv[i];
        stateless_pass(sysroot);
        b2dp_ctx* c = nullptr;
        const int rc = b2dp_open(("kfd:" + sysroot).c_str(), &c);
        if (rc != B2DP_OK) { EXPECT(rc == B2DP_E_NODRIVER && c == nullptr); continue; }
        (void)ctx_pass(c, false);                     // before Start(): "Init method must be called" paths
        const uint64_t h1 = ctx_pass(c, true);
        const uint64_t h2 = ctx_pass(c, true);        // re-Start replaces the policy; same answers
        EXPECT(h1 == h2);
        digest ^= h1;
        allocator_pass(c, sysroot);
        const std::string exp = scratch + "/export" + std::to_string(i);
        const int erc = b2dp_export_kfd_tree(c, exp.c_str());
        if (erc == B2DP_OK) {  // exported tree of a kfd context enumerates to the same device table
            b2dp_ctx* c2 = nullptr;
            if (b2dp_open(("kfd:" + exp).c_str(), &c2) == B2DP_OK) {
                Ids a, b;
                EXPECT(enumerate_all(c, a) == enumerate_all(c2, b));
                b2dp_close(c2);
            }
        }
        b2dp_close(c);
    }
    for (const char* uri : {"synthetic:8,mig=7", "synthetic:1", "synthetic:16,mig=2,compute=cpx,memory=nps4"}) {
        b2dp_ctx* sc = nullptr;
        EXPECT(b2dp_open(uri, &sc) == B2DP_OK && sc);
        if (!sc) continue;
        (void)ctx_pass(sc, false);
        const uint64_t h1 = ctx_pass(sc, true), h2 = ctx_pass(sc, true);
        EXPECT(h1 == h2);
        digest ^= h1;
        b2dp_close(sc);  // removes the generated tree
    }
    b2dp_ctx* c = nullptr;
    EXPECT(b2dp_open("synthetic:0", &c) == B2DP_E_INVAL && c == nullptr);
    EXPECT(b2dp_open("bogus:", &c) == B2DP_E_INVAL);
    EXPECT(b2dp_open("cuda:", &c) == B2DP_E_NOGPU);      // stubbed in this build
    EXPECT(b2dp_open("cuda:slots=1", &c) == B2DP_E_INVAL);
    printf("sweep ok: %d trees, digest %016llx, %d failures\n", argc - 3, (unsigned long long)digest, g_fail);
    return g_fail ? 1 : 0;
}

struct WatchCount { std::atomic<int> sends{0}; std::atomic<int> errors{0}; std::atomic<uint64_t> last{0}; };
static void on_send(void* user, int rc, const uint8_t* buf, size_t len, const b2dp_cycle_stats*) {
    auto* w = (WatchCount*)user;
    if (rc != B2DP_OK) { w->errors++; return; }
    w->last.store(fnv(1469598103934665603ull, buf, len));
    w->sends++;
}

static int cmd_threads(char** argv) {
    // argv[2]: a sysroot (opened as kfd:<sysroot>) or a full backend uri ("cuda:probe=off,...": the NVML-driven units
    // backend against whatever NVML B2DP_NVML_LIBRARY names)
    const std::string arg = argv[2];
    const std::string uri = arg.compare(0, 5, "cuda:") == 0 || arg.compare(0, 10, "synthetic:") == 0 ? arg : "kfd:" + arg;
    const int T = atoi(argv[3]), iters = atoi(argv[4]);
    b2dp_ctx* c = nullptr;
    if (b2dp_open(uri.c_str(), &c) != B2DP_OK) { fprintf(stderr, "open failed: %s\n", b2dp_last_error(nullptr)); return 2; }
    (void)b2dp_start(c);
    const uint64_t want = ctx_pass(c, true);
    WatchCount wc;
    b2dp_watch* w = nullptr;
    b2dp_cycle_opts wo{};
    wo.flags = B2DP_LW_NO_PROBE;
    EXPECT(b2dp_watch_start(c, "gpu", 1, &wo, on_send, &wc, &w) == B2DP_OK);
    std::atomic<int> bad{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            for (int i = 0; i < iters; ++i) {
                if (ctx_pass(c, (i + t) % 3 == 0) != want) bad++;   // Start() races with allocation on purpose
                if (w && i % 2 == 0) (void)b2dp_watch_beat(w);
                if (i % 4 == t % 4) {                               // private contexts come and go meanwhile
                    b2dp_ctx* p = nullptr;
                    if (b2dp_open(uri.c_str(), &p) == B2DP_OK) { if (ctx_pass(p, true) != want) bad++; b2dp_close(p); }
                }
            }
        });
    for (auto& x : th) x.join();
    if (w) b2dp_watch_stop(w);
    EXPECT(bad.load() == 0);
    EXPECT(wc.errors.load() == 0 && wc.sends.load() >= 1);
    b2dp_close(c);
    printf("threads ok: %d threads x %d iterations, %d watch sends, %d mismatching passes, %d failures\n", T, iters,
           wc.sends.load(), bad.load(), g_fail);
    return g_fail ? 1 : 0;
}

// probe=helpers against whatever B2DP_PROBE_HELPER names (tests/native/fake_probe_helper.cpp): T threads run probe
// fan-outs and heartbeat cycles on ONE context (the backend serialises fan-outs), children are spawned, fed and reaped.
static int cmd_helpers(char** argv) {
    const int T = atoi(argv[3]), iters = atoi(argv[4]);
    b2dp_ctx* c = nullptr;
    if (b2dp_open(argv[2], &c) != B2DP_OK) { fprintf(stderr, "open failed: %s\n", b2dp_last_error(nullptr)); return 2; }
    Ids ids;
    EXPECT(enumerate_all(c, ids) == B2DP_OK);
    const int n = (int)ids.devs.size();
    std::atomic<int> bad{0}, passes{0};
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] {
            std::vector<b2dp_probe_result> res((size_t)n + 1);
            std::vector<uint8_t> buf(1 << 14);
            for (int i = 0; i < iters; ++i) {
                int m = 0;
                if ((i + t) % 2 == 0) {
                    if (b2dp_probe_health(c, nullptr, res.data(), (int)res.size(), &m) != B2DP_OK || m != n) { bad++; continue; }
                    for (int k = 0; k < m; ++k)
                        if (!res[(size_t)k].healthy || res[(size_t)k].checksum != res[(size_t)k].expected_checksum || res[(size_t)k].device != k) bad++;
                } else {
                    b2dp_cycle_opts co{};
                    co.flags = B2DP_LW_HEARTBEAT;
                    size_t len = 0;
                    b2dp_cycle_stats st{};
                    if (b2dp_list_and_watch(c, "1g_23gb", &co, buf.data(), buf.size(), &len, &st) != B2DP_OK || st.n_devices != n || st.n_unhealthy) bad++;
                }
                passes++;
            }
        });
    for (auto& x : th) x.join();
    b2dp_probe_info pi{};
    EXPECT(b2dp_probe_describe(c, n - 1, &pi) == B2DP_OK && pi.via_helper == 1 && pi.usable == 1);
    EXPECT(bad.load() == 0);
    b2dp_close(c);
    printf("helpers ok: %d units, %d threads, %d passes, %d bad, %d failures\n", n, T, passes.load(), bad.load(), g_fail);
    return g_fail ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc == 5 && !strcmp(argv[1], "helpers")) return cmd_helpers(argv);
    if (argc >= 4 && !strcmp(argv[1], "sweep")) return cmd_sweep(argc, argv);
    if (argc == 5 && !strcmp(argv[1], "threads")) return cmd_threads(argv);
    fprintf(stderr, "usage: abi_stress sweep <scratch> <sysroot>... | threads <sysroot> <T> <iters>\n");
    return 2;
}
