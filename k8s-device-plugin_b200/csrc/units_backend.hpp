// units_backend.hpp -- the NVML-driven half of the `cuda:` backend: enumeration without a CUDA context, MIG devices,
// and probing through helper processes.  Included by cuda_backend.cu only.
//
// Replaces, for MIG-partitioned B200s, the platform-device loop of GetAMDGPUs (amdgpu.go:221-265) and the partition
// resource naming of cmd/k8s-device-plugin/main.go:53-91.  A MIG-enabled GPU is listed the way the reference lists an
// MI300 in CPX mode: the first instance under the GPU's own PCI id (the reference's PCI function IS partition 0), the
// others as amdgpu_xcp_<8*gpu + p> platform devices; every instance carries the parent's devID (so the allocator groups
// them per physical GPU, allocator/device.go:287-304), its compute / memory partition strings are the MIG profile
// ("<N>g" / "<M>gb", lower case like amdgpu.go:172,179), and b2dp_export_kfd_tree() writes exactly the tree on which
// the reference algorithm reproduces this table.
//
// Modes (URI option probe=):  helpers -- one b200dp_probe_helper child per listed unit, each with
// CUDA_VISIBLE_DEVICES=<unit UUID> (the only way to reach every MIG compute instance: CUDA shows a process one);
// off -- enumeration, allocation and labels only (what a node labeller needs: no context, no HBM ring).
// With mig=auto (default) a node that has any MIG-enabled GPU uses helpers for every unit, because CUDA hides
// the other GPUs from a process that sees a MIG instance.
#pragma once
#include <fcntl.h>
#include <poll.h>
#include <signal.h>
#include <spawn.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "gosem.hpp"
#include "helper_proto.hpp"
#include "internal.hpp"
#include "nvml_dyn.hpp"

extern char** environ;

namespace b2dp {

struct Unit {
    Device dev;                   // the enumerate record
    std::string uuid;             // GPU-... / MIG-...
    int gpu_pos = 0;              // position of the physical GPU in the BDF-sorted list
    int nvml_index = -1;          // NVML index of the physical GPU
    int mig_slot = -1;            // MIG device index under its parent (NVIDIA_VISIBLE_DEVICES "<gpu>:<slot>"), -1 = whole GPU
    unsigned gi = 0, ci = 0;
    int parent_minor = 0;         // /dev/nvidia<N>
    int cap_gi = -1, cap_ci = -1; // /dev/nvidia-caps/nvidia-cap<N>
    void* nvh = nullptr;          // NVML handle of the physical GPU
    std::string name, family, pci_device_id, vbios;
    std::vector<std::pair<std::string, std::string>> firmware;
    int64_t vram = 0, sms = 0;
    bool nvlink_up = false;       // the physical GPU has an active NVLink (declared links)
    bool mig_capable = false;
    // helper process
    pid_t pid = -1;
    int fd = -1;
    uint64_t seq = 0, stale = 0;  // stale: responses of timed-out requests still to be discarded
    bool broken = false;
    std::string broken_reason;
    int last_healthy = 1;
    uint64_t slot_bytes = 0;
    float gbs_cal = 0.f, gbs_ref = 0.f;
    std::string helper_uri;
    bool starting = false;        // a (re)started child has been sent HELLO and has not answered yet
    uint64_t hello_seq = 0;
    std::chrono::steady_clock::time_point start_t{};
};

struct UnitsBackend {
    std::vector<Unit> units;  // sorted by dev.id (canonical order)
    bool helpers = false;
    std::string helper_exe;
};

// ---- small helpers ------------------------------------------------------------------------------------------------
static inline std::string units_read_trim(const std::string& p) {
    std::string d;
    if (!go::read_file(p, d)) return "";
    return go::trim_space(d);
}

// /proc/driver/nvidia-caps/mig-minors: "gpu<minor>/gi<G>/access <cap>" and "gpu<minor>/gi<G>/ci<C>/access <cap>"
static inline void units_cap_minors(const std::string& sysroot, int gpu_minor, unsigned gi, unsigned ci, int* cap_gi, int* cap_ci) {
    std::string data;
    if (!go::read_file(go::join(sysroot, "proc/driver/nvidia-caps/mig-minors"), data)) return;
    char want_gi[64], want_ci[64];
    snprintf(want_gi, sizeof want_gi, "gpu%d/gi%u/access", gpu_minor, gi);
    snprintf(want_ci, sizeof want_ci, "gpu%d/gi%u/ci%u/access", gpu_minor, gi, ci);
    size_t pos = 0;
    while (pos < data.size()) {
        size_t e = data.find('\n', pos);
        if (e == std::string::npos) e = data.size();
        const std::string line = data.substr(pos, e - pos);
        pos = e + 1;
        const size_t sp = line.find(' ');
        if (sp == std::string::npos) continue;
        const std::string key = line.substr(0, sp);
        if (key == want_gi) *cap_gi = atoi(line.c_str() + sp + 1);
        else if (key == want_ci) *cap_ci = atoi(line.c_str() + sp + 1);
    }
}

static inline std::string family_of_cc(int major, int minor) {
    return major == 10 || major == 12 ? "Blackwell" : major == 9 ? "Hopper" : major == 8 ? (minor == 9 ? "Ada" : "Ampere")
           : "sm_" + std::to_string(major * 10 + minor);
}

// ---- helper process plumbing ----------------------------------------------------------------------------------------
static inline bool units_send(Unit& u, const HelperReq& q) {
    const char* c = reinterpret_cast<const char*>(&q);
    size_t n = sizeof q;
    while (n) {
        const ssize_t r = send(u.fd, c, n, MSG_NOSIGNAL);  // a dead child must not SIGPIPE the host
        if (r < 0) { if (errno == EINTR) continue; return false; }
        c += r; n -= (size_t)r;
    }
    return true;
}

// Read exactly n bytes before `deadline`; 1 ok, 0 timeout, -1 child gone.
static inline int units_recv(Unit& u, void* p, size_t n, std::chrono::steady_clock::time_point deadline) {
    char* c = static_cast<char*>(p);
    while (n) {
        const auto now = std::chrono::steady_clock::now();
        struct pollfd pf{u.fd, POLLIN, 0};
        const long long left = std::chrono::duration_cast<std::chrono::milliseconds>(deadline - now).count();
        const int pr = poll(&pf, 1, left > 0 ? (int)left : 0);  // a deadline already reached still gets one look
        if (pr < 0) { if (errno == EINTR) continue; return -1; }
        if (pr == 0) { if (std::chrono::steady_clock::now() >= deadline) return 0; continue; }
        const ssize_t r = recv(u.fd, c, n, 0);
        if (r == 0) return -1;
        if (r < 0) { if (errno == EINTR || errno == EAGAIN) continue; return -1; }
        c += r; n -= (size_t)r;
    }
    return 1;
}

static inline void units_kill(Unit& u) {
    if (u.fd >= 0) { close(u.fd); u.fd = -1; }
    if (u.pid > 0) {
        int st = 0;
        for (int i = 0; i < 200 && waitpid(u.pid, &st, WNOHANG) == 0; ++i) usleep(5000);  // stdin closed: it exits by itself
        if (waitpid(u.pid, &st, WNOHANG) == 0) { kill(u.pid, SIGKILL); waitpid(u.pid, &st, 0); }
        u.pid = -1;
    }
}

// One request/response exchange (not the fan-out): discards stale responses first.
static inline int units_call(Unit& u, HelperReq q, HelperRsp* r, int timeout_ms, std::vector<uint32_t>* payload = nullptr) {
    if (u.fd < 0) return B2DP_E_CUDA;
    q.magic = kHelperMagic;
    q.seq = ++u.seq;
    if (!units_send(u, q)) return B2DP_E_CUDA;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    for (;;) {
        const int rc = units_recv(u, r, sizeof *r, deadline);
        if (rc == 0) { u.stale++; return B2DP_E_TIMEOUT; }
        if (rc < 0 || r->magic != kHelperMagic) return B2DP_E_CUDA;
        if (r->seq == q.seq) break;
        if (u.stale) u.stale--;  // a late answer to a request that had timed out
    }
    if (payload && q.op == HOP_PEEK && r->rc == B2DP_OK) {
        payload->assign((size_t)q.b, 0);
        if (units_recv(u, payload->data(), (size_t)q.b * 4, deadline) != 1) return B2DP_E_CUDA;
    }
    return B2DP_OK;
}

static inline bool units_spawn(UnitsBackend* ub, Unit& u, std::string& err) {
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0, sv) != 0) { err = std::string("socketpair: ") + strerror(errno); return false; }
    posix_spawn_file_actions_t fa;
    posix_spawn_file_actions_init(&fa);
    posix_spawn_file_actions_adddup2(&fa, sv[1], 0);
    posix_spawn_file_actions_adddup2(&fa, sv[1], 1);
    std::vector<std::string> env_store;
    for (char** e = environ; e && *e; ++e)
        if (strncmp(*e, "CUDA_VISIBLE_DEVICES=", 21) != 0 && strncmp(*e, "B2DP_NVML_LIBRARY=", 18) != 0) env_store.push_back(*e);
    env_store.push_back("CUDA_VISIBLE_DEVICES=" + u.uuid);
    std::vector<char*> envp;
    for (auto& s : env_store) envp.push_back(const_cast<char*>(s.c_str()));
    envp.push_back(nullptr);
    char* argv[] = {const_cast<char*>(ub->helper_exe.c_str()), const_cast<char*>(u.helper_uri.c_str()), nullptr};
    pid_t pid = -1;
    const int rc = posix_spawn(&pid, ub->helper_exe.c_str(), &fa, nullptr, argv, envp.data());
    posix_spawn_file_actions_destroy(&fa);
    close(sv[1]);
    if (rc != 0) { close(sv[0]); err = "posix_spawn(" + ub->helper_exe + "): " + strerror(rc); return false; }
    u.pid = pid;
    u.fd = sv[0];
    u.seq = u.stale = 0;
    return true;
}

// (Re)start u's helper without waiting for it: spawn, send HELLO.  A fresh CUDA process needs seconds to come up;
// a heartbeat must not stall on that, so the answer is collected later (units_finish_start).
static inline bool units_begin_start(UnitsBackend* ub, Unit& u, std::string& err) {
    units_kill(u);
    u.starting = false;
    if (!units_spawn(ub, u, err)) return false;
    HelperReq q{};
    q.magic = kHelperMagic;
    q.op = HOP_HELLO;
    q.seq = u.hello_seq = ++u.seq;
    if (!units_send(u, q)) { err = "probe helper exited at start"; units_kill(u); return false; }
    u.starting = true;
    u.start_t = std::chrono::steady_clock::now();
    return true;
}

// Has the child answered HELLO?  1 ready (u describes what it holds), 0 not yet, -1 failed (child dropped, err says why).
static inline int units_finish_start(Unit& u, int wait_ms, std::string& err) {
    HelperRsp r{};
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(wait_ms);
    for (;;) {
        const int rc = units_recv(u, &r, sizeof r, deadline);
        if (rc == 0) {
            if (std::chrono::steady_clock::now() - u.start_t < std::chrono::seconds(120)) return 0;
            err = "probe helper did not answer within 120 s";
            units_kill(u); u.starting = false;
            return -1;
        }
        if (rc < 0 || r.magic != kHelperMagic) { err = "probe helper exited while starting"; units_kill(u); u.starting = false; return -1; }
        if (r.seq == u.hello_seq) break;
    }
    u.starting = false;
    if (r.rc != B2DP_OK) { err = std::string("probe helper: ") + r.text; units_kill(u); return -1; }
    if (r.text[0]) u.name = r.text;
    if (r.extra[0]) u.sms = (int64_t)r.extra[0];
    if (r.extra[1]) u.vram = (int64_t)r.extra[1];
    u.slot_bytes = r.extra[2];
    memcpy(&u.gbs_cal, &r.extra[3], sizeof(float));
    if (u.gbs_ref > 0) {  // a restarted child inherits the ceiling its predecessor was judged against
        HelperReq q{};
        q.op = HOP_SETREF;
        memcpy(&q.a, &u.gbs_ref, sizeof(float));
        HelperRsp r2{};
        units_call(u, q, &r2, 2000);
    }
    return 1;
}

// ---- open -------------------------------------------------------------------------------------------------------------
// Returns B2DP_OK and *out, or an error.  `want_helpers`: probe=helpers (or forced by MIG); false: probe=off.
static inline int units_open(const CudaConfig& cfg, Nvml& nv, bool want_helpers, std::unique_ptr<UnitsBackend>* out, std::string& err) {
    if (!nv.ok || !nv.device_count || !nv.handle_by_index || !nv.pci_info) { err = "NVML is not available (needed for probe=off/helpers and MIG)"; return B2DP_E_NOGPU; }
    unsigned count = 0;
    if (nv.device_count(&count) != 0 || count == 0) { err = "NVML reports no devices"; return B2DP_E_NOGPU; }
    auto ub = std::make_unique<UnitsBackend>();
    ub->helpers = want_helpers;

    int n_cpu = 0;  // CPU topology nodes precede GPU nodes in a kfd tree
    for (auto& p : go::glob_prefixed(go::join(cfg.sysroot, "sys/devices/system/node"), "node"))
        if (p.size() > 4 && go::is_digit(p.back())) ++n_cpu;
    if (n_cpu < 1) n_cpu = 1;

    struct Phys { std::string bdf, dev_id; void* h; unsigned nvml_index; };
    std::vector<Phys> phys;
    for (unsigned i = 0; i < count; ++i) {
        if (!cfg.devices.empty() && std::find(cfg.devices.begin(), cfg.devices.end(), (int)i) == cfg.devices.end()) continue;
        void* h = nullptr;
        NvmlPciInfo pi{};
        if (nv.handle_by_index(i, &h) != 0 || nv.pci_info(h, &pi) != 0) continue;
        char bdf[32], devid[32];
        snprintf(bdf, sizeof bdf, "%04x:%02x:%02x.0", pi.domain, pi.bus, pi.device);
        snprintf(devid, sizeof devid, "%04x:%02x:%02x:0", pi.domain, pi.bus, pi.device);  // amdgpu.go:141
        phys.push_back({bdf, devid, h, i});
    }
    if (phys.empty()) { err = "no NVML device selected"; return B2DP_E_NOGPU; }
    std::sort(phys.begin(), phys.end(), [](const Phys& a, const Phys& b) { return a.bdf < b.bdf; });

    int running = 0;  // card / renderD / node counters run over the listed units like the kfd tree's do
    for (size_t g = 0; g < phys.size(); ++g) {
        Unit base;
        base.gpu_pos = (int)g;
        base.nvml_index = (int)phys[g].nvml_index;
        base.nvh = phys[g].h;
        base.dev.dev_id = phys[g].dev_id;
        char buf[128] = {0};
        if (nv.name_of && nv.name_of(phys[g].h, buf, sizeof buf) == 0) base.name = buf;
        const std::string pci_dir = go::join(cfg.sysroot, "sys/bus/pci/devices/" + phys[g].bdf);
        base.pci_device_id = units_read_trim(pci_dir + "/device");
        if (base.pci_device_id.empty()) base.pci_device_id = "0x0000";
        int64_t nvv = 0;
        const std::string numa = units_read_trim(pci_dir + "/numa_node");
        base.dev.numa = (!numa.empty() && go::atoi(numa, &nvv) == go::NumErr::none) ? (int)nvv : 0;
        unsigned mn = 0;
        if (nv.minor_number && nv.minor_number(phys[g].h, &mn) == 0) base.parent_minor = (int)mn;
        if (nv.vbios && nv.vbios(phys[g].h, buf, sizeof buf) == 0) base.vbios = buf;
        if (!base.vbios.empty()) base.firmware.push_back({"vbios", base.vbios});
        if (nv.inforom_image && nv.inforom_image(phys[g].h, buf, sizeof buf) == 0) base.firmware.push_back({"inforom-img", buf});
        if (nv.inforom_object) {
            if (nv.inforom_object(phys[g].h, 0, buf, sizeof buf) == 0) base.firmware.push_back({"inforom-oem", buf});
            if (nv.inforom_object(phys[g].h, 1, buf, sizeof buf) == 0) base.firmware.push_back({"inforom-ecc", buf});
            if (nv.inforom_object(phys[g].h, 2, buf, sizeof buf) == 0) base.firmware.push_back({"inforom-pwr", buf});
        }
        if (nv.gsp_firmware && nv.gsp_firmware(phys[g].h, buf) == 0) base.firmware.push_back({"gsp", buf});
        int ccM = 10, ccm = 0;
        if (nv.cuda_cc) nv.cuda_cc(phys[g].h, &ccM, &ccm);
        base.family = family_of_cc(ccM, ccm);
        unsigned active = 0;
        if (nv.nvlink_state && nv.nvlink_state(phys[g].h, 0, &active) == 0 && active) base.nvlink_up = true;
        unsigned cur = 0, pend = 0;
        const bool mig_known = nv.mig_mode && nv.mig_mode(phys[g].h, &cur, &pend) == 0;
        base.mig_capable = mig_known;

        std::vector<Unit> list;
        if (mig_known && cur == 1 && cfg.mig_auto && nv.max_mig_count && nv.mig_handle) {
            unsigned maxc = 0;
            nv.max_mig_count(phys[g].h, &maxc);
            struct Inst { unsigned slot, gi, ci; void* h; };
            std::vector<Inst> insts;
            for (unsigned s = 0; s < maxc; ++s) {
                void* mh = nullptr;
                if (nv.mig_handle(phys[g].h, s, &mh) != 0) continue;  // NVML_ERROR_NOT_FOUND: empty slot
                Inst in{s, 0, 0, mh};
                if (nv.gi_id) nv.gi_id(mh, &in.gi);
                if (nv.ci_id) nv.ci_id(mh, &in.ci);
                insts.push_back(in);
            }
            std::sort(insts.begin(), insts.end(), [](const Inst& a, const Inst& b) { return a.gi != b.gi ? a.gi < b.gi : a.ci < b.ci; });
            for (size_t p = 0; p < insts.size(); ++p) {
                Unit u = base;
                u.mig_slot = (int)insts[p].slot;
                u.gi = insts[p].gi;
                u.ci = insts[p].ci;
                // the reference's CPX shape: partition 0 is the PCI function itself, the rest are platform devices
                u.dev.id = p == 0 ? phys[g].bdf : "amdgpu_xcp_" + std::to_string(g * 8 + p);
                if (nv.uuid_of && nv.uuid_of(insts[p].h, buf, sizeof buf) == 0) u.uuid = buf;
                NvmlDeviceAttributes at{};
                if (nv.attributes && nv.attributes(insts[p].h, &at) == 0) {
                    u.sms = at.multiprocessorCount;
                    u.vram = (int64_t)at.memorySizeMB << 20;
                    // MIG profile "<N>g.<M>gb" -> compute "<N>g", memory "<M>gb" (resource name "<N>g_<M>gb", main.go:62-89)
                    const unsigned long long gb = (at.memorySizeMB + 512) / 1024;
                    u.dev.compute = std::to_string(at.computeInstanceSliceCount ? at.computeInstanceSliceCount : at.gpuInstanceSliceCount) + "g";
                    u.dev.memory = std::to_string(gb) + "gb";
                }
                if (nv.name_of && nv.name_of(insts[p].h, buf, sizeof buf) == 0 && buf[0]) {
                    // "NVIDIA B200 MIG 1g.23gb": the profile in the name is authoritative for the memory size label
                    const char* m = strstr(buf, "MIG ");
                    unsigned ng = 0, ngb = 0;
                    if (m && sscanf(m + 4, "%ug.%ugb", &ng, &ngb) == 2) { u.dev.compute = std::to_string(ng) + "g"; u.dev.memory = std::to_string(ngb) + "gb"; }
                }
                units_cap_minors(cfg.sysroot, u.parent_minor, u.gi, u.ci, &u.cap_gi, &u.cap_ci);
                list.push_back(std::move(u));
            }
        }
        if (list.empty() && !(mig_known && cur == 1 && cfg.mig_auto)) {  // a whole GPU (a MIG-enabled GPU with no instance lists nothing)
            Unit u = base;
            u.dev.id = phys[g].bdf;
            if (nv.uuid_of && nv.uuid_of(phys[g].h, buf, sizeof buf) == 0) u.uuid = buf;
            NvmlMemory mem{};
            if (nv.memory_info && nv.memory_info(phys[g].h, &mem) == 0) u.vram = (int64_t)mem.total;
            unsigned cores = 0;
            if (nv.num_cores && nv.num_cores(phys[g].h, &cores) == 0) u.sms = cores / 128;
            list.push_back(std::move(u));
        }
        for (auto& u : list) {
            u.dev.card = running;
            u.dev.render_d = 128 + running;
            u.dev.node_id = n_cpu + running;
            ++running;
            ub->units.push_back(std::move(u));
        }
    }
    if (ub->units.empty()) { err = "no device to list (MIG enabled but no instance created?)"; return B2DP_E_NOGPU; }
    // whole GPUs keep the reference-facing meaning of `card`: the device minor (Allocate mounts /dev/nvidia<card>)
    bool any_mig = false;
    for (auto& u : ub->units) any_mig = any_mig || u.mig_slot >= 0;
    if (!any_mig) for (auto& u : ub->units) u.dev.card = u.parent_minor;
    // On a node that mixes MIG-partitioned and whole GPUs the whole ones carry their full profile ("7g" / "<M>gb"), the way
    // an MI300 left in SPX mode reports "spx": the reference's histogram (amdgpu.go:270-285) skips devices without both
    // partition strings, which would fold the whole GPUs into the partitions' resource.
    if (any_mig)
        for (auto& u : ub->units)
            if (u.mig_slot < 0) {
                unsigned maxc = 7;
                if (nv.max_mig_count) nv.max_mig_count(u.nvh, &maxc);
                u.dev.compute = std::to_string(maxc) + "g";
                u.dev.memory = std::to_string((u.vram + (1ll << 29)) >> 30) + "gb";
            }
    std::sort(ub->units.begin(), ub->units.end(), [](const Unit& a, const Unit& b) { return a.dev.id < b.dev.id; });

    if (want_helpers) {
        const char* ov = getenv("B2DP_PROBE_HELPER");
        if (ov && *ov) ub->helper_exe = ov;
        else {
            Dl_info di{};
            if (dladdr(reinterpret_cast<void*>(&units_read_trim), &di) && di.dli_fname) ub->helper_exe = go::dir(di.dli_fname) + "/b200dp_probe_helper";
        }
        if (ub->helper_exe.empty() || access(ub->helper_exe.c_str(), X_OK) != 0) {
            err = "probe helper executable not found: " + ub->helper_exe + " (build it, or set B2DP_PROBE_HELPER)";
            return B2DP_E_NOGPU;
        }
        size_t ok = 0;
        // all children are started before any is waited for: N fresh CUDA processes come up in parallel
        for (size_t i = 0; i < ub->units.size(); ++i) {
            Unit& u = ub->units[i];
            const uint64_t bytes = u.mig_slot >= 0 ? std::min<uint64_t>(cfg.mig_bytes, cfg.bytes) : cfg.bytes;
            u.helper_uri = "cuda:devices=0,probe=inproc,mig=off,bytes=" + std::to_string(bytes) + ",slots=" + std::to_string(cfg.slots) +
                           ",seed_index=" + std::to_string(i) + cfg.passthrough;
            const bool forced = std::find(cfg.break_devices.begin(), cfg.break_devices.end(), (int)i) != cfg.break_devices.end();
            std::string e2;
            if (forced || !units_begin_start(ub.get(), u, e2)) {
                u.broken = true;
                u.broken_reason = forced ? "setup failure injected (break=)" : e2;
            }
        }
        for (auto& u : ub->units) {
            if (!u.broken) {
                std::string e2;
                int st = 0;
                while ((st = units_finish_start(u, 1000, e2)) == 0) {}
                if (st > 0) { ++ok; continue; }
                u.broken = true;
                u.broken_reason = e2;
            }
            u.last_healthy = 0;
            err = u.broken_reason + " on " + u.dev.id;
            logf(2, "%s could not be set up and will be reported Unhealthy: %s", u.dev.id.c_str(), u.broken_reason.c_str());
        }
        if (!ok) { for (auto& u : ub->units) units_kill(u); return B2DP_E_CUDA; }
        // the ceiling: ref_gbs= or the best calibration among units of the same product and ring size
        for (auto& u : ub->units) {
            if (u.broken) continue;
            float ref = cfg.ref_gbs;
            if (ref <= 0)
                for (auto& o : ub->units)
                    if (!o.broken && o.name == u.name && o.slot_bytes == u.slot_bytes) ref = std::max(ref, o.gbs_cal);
            u.gbs_ref = ref;
            HelperReq q{};
            q.op = HOP_SETREF;
            memcpy(&q.a, &ref, sizeof ref);
            HelperRsp r{};
            units_call(u, q, &r, 5000);
        }
    }
    *out = std::move(ub);
    return B2DP_OK;
}

static inline void units_close(UnitsBackend* ub) {
    for (auto& u : ub->units) {
        if (u.fd >= 0) {
            HelperReq q{};
            q.magic = kHelperMagic;
            q.op = HOP_QUIT;
            units_send(u, q);
        }
    }
    for (auto& u : ub->units) units_kill(u);
}

// ---- probe fan-out over the helpers ---------------------------------------------------------------------------------
static inline int units_probe(UnitsBackend* ub, const b2dp_probe_opts* opts, std::vector<b2dp_probe_result>& out, std::string& err) {
    const size_t n = ub->units.size();
    const uint32_t timeout_ms = opts && opts->timeout_ms ? opts->timeout_ms : 5000;  // health.go:37
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    out.assign(n, b2dp_probe_result{});
    std::vector<char> state(n, 0);  // 0 waiting, 1 done, 2 failed
    std::vector<uint64_t> want(n, 0);
    for (size_t i = 0; i < n; ++i) {  // send everywhere before waiting anywhere: the N passes run concurrently
        Unit& u = ub->units[i];
        out[i].device = (int)i;
        out[i].first_bad_word = ~0ull;
        if (u.broken) {
            // A helper that died is restarted by the next heartbeat -- without stalling it: the child is spawned and
            // greeted, and whichever heartbeat finds its answer (this one, if it comes within 250 ms; a fresh CUDA
            // process takes seconds) puts the unit back in service.  Until then the unit is reported Unhealthy.
            std::string e2;
            bool ready = false;
            if (!u.helper_uri.empty() && u.broken_reason.find("injected") == std::string::npos) {
                if (u.starting || units_begin_start(ub, u, e2)) {
                    const int st = units_finish_start(u, (int)std::min<uint32_t>(timeout_ms, 250), e2);
                    ready = st > 0;
                    if (st == 0) e2 = "probe helper is starting";
                }
                if (!ready && !e2.empty()) u.broken_reason = e2;
            }
            if (!ready) { out[i].err = B2DP_E_CUDA; state[i] = 2; err = u.broken_reason + " on " + u.dev.id; continue; }
            logf(0, "%s: probe helper (pid %d) is back in service", u.dev.id.c_str(), (int)u.pid);
            u.broken = false;
        }
        HelperReq q{};
        q.magic = kHelperMagic;
        q.op = HOP_PROBE;
        q.seq = want[i] = ++u.seq;
        if (opts) q.opts = *opts;
        if (!units_send(u, q)) { u.broken = true; u.broken_reason = "probe helper exited"; out[i].err = B2DP_E_CUDA; state[i] = 2; err = u.broken_reason + " on " + u.dev.id; }
    }
    for (size_t i = 0; i < n; ++i) {
        if (state[i] != 0) continue;
        Unit& u = ub->units[i];
        HelperRsp r{};
        for (;;) {
            const int rc = units_recv(u, &r, sizeof r, deadline);
            if (rc == 0) { u.stale++; out[i].err = B2DP_E_TIMEOUT; state[i] = 2; break; }
            if (rc < 0 || r.magic != kHelperMagic) {
                logf(2, "%s: probe helper (pid %d) exited; reported Unhealthy, a replacement starts with the next heartbeat", u.dev.id.c_str(), (int)u.pid);
                u.broken = true; u.broken_reason = "probe helper exited"; units_kill(u);
                out[i].err = B2DP_E_CUDA; state[i] = 2; err = u.broken_reason + " on " + u.dev.id; break;
            }
            if (r.seq != want[i]) { if (u.stale) u.stale--; continue; }
            if (r.rc != B2DP_OK) { out[i].err = r.rc; state[i] = 2; err = std::string(r.text) + " on " + u.dev.id; break; }
            out[i] = r.res;
            out[i].device = (int)i;
            state[i] = 1;
            break;
        }
        if (state[i] == 2) { out[i].healthy = 0; u.last_healthy = 0; } else u.last_healthy = out[i].healthy;
    }
    return B2DP_OK;
}

// Declared link class between two units (no cross-process P2P measurement exists): instances of one GPU share the
// package; different GPUs are NVLink-class when NVML reports an NVLink P2P path (or, in MIG mode where P2P is
// disabled, active NVLinks on both parents), PCIe-class otherwise.
static inline int units_link_type(const Nvml& nv, const Unit& a, const Unit& b) {
    if (a.gpu_pos == b.gpu_pos) return 11;
    int status = -1;
    if (nv.p2p_status && nv.p2p_status(a.nvh, b.nvh, 2 /*NVML_P2P_CAPS_INDEX_NVLINK*/, &status) == 0 && status == 0) return 11;
    if (a.nvlink_up && b.nvlink_up) return 11;
    if (nv.p2p_status && nv.p2p_status(a.nvh, b.nvh, 0 /*NVML_P2P_CAPS_INDEX_READ*/, &status) == 0 && status == 0) return 2;
    return a.mig_slot >= 0 || b.mig_slot >= 0 ? 2 : 0;
}

}  // namespace b2dp
