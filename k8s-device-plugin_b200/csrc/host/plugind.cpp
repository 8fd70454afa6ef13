// b200dp_plugind -- the device-plugin daemon as ONE native binary over libb200dp.so: the flags and resource list
// of cmd/k8s-device-plugin/main.go:93-155, one v1beta1.DevicePlugin gRPC server per resource on
// <plugin_dir>/<namespace>_<name>, registration with the kubelet, the `-pulse` heartbeat ticker, and dpm's
// lifecycle (vendor/github.com/kubevirt/device-plugin-manager/pkg/dpm): plugin start retried 3 times 3 s apart
// (manager.go:16-20,205-219), re-serve + re-register when kubelet.sock is re-created (manager.go:73-84; polled
// once a second instead of fsnotify), clean stop on SIGINT/SIGQUIT/SIGTERM (manager.go:47-48,85-91).
// gRPC comes from host/h2grpc.hpp; every RPC body is one C-ABI call (include/b200dp.h).
//
//   b200dp_plugind -pulse=10 -resource_naming_strategy=single -backend=cuda:xid=1 [-plugin_dir DIR]
//   kill -USR1 <pid>     one heartbeat now
#include <signal.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../../include/b200dp.h"
#include "../pbwire.hpp"
#include "h2grpc.hpp"
#include "pbread.hpp"

namespace {

std::atomic<bool> g_stop{false};

struct Flags {
    int pulse = 0;                                   // main.go:109
    std::string strategy = "single";                 // main.go:110
    std::string backend = "cuda:";
    std::string plugin_dir = "/var/lib/kubelet/device-plugins/";  // pluginapi.DevicePluginPath, constants.go:26
    std::string ns = "amd.com";                      // plugin.go:406-408
    double start_retry_wait = 3.0;                   // dpm/manager.go:19
    int link_check = 0;
    std::string exporter_socket;                     // optional: serve metricssvc.MetricsService here (health.go:36)
    bool version = false;
};

// Go's flag package: -name=value, -name value, --name...; bools not needed here
bool parse_flags(int argc, char** argv, Flags& f, std::string& err) {
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a.rfind("--", 0) == 0) a = a.substr(1);
        if (a == "-version") { f.version = true; continue; }
        if (a.empty() || a[0] != '-') { err = "unexpected argument " + a; return false; }
        std::string name = a.substr(1), val;
        const size_t eq = name.find('=');
        const std::string bare = eq == std::string::npos ? name : name.substr(0, eq);
        // glog's flags (the reference's image runs `-logtostderr=true -stderrthreshold=INFO -v=5`, Dockerfile:33) are
        // accepted and ignored so existing DaemonSet args keep working; this daemon always logs to stderr
        const bool glog_bool = bare == "logtostderr" || bare == "alsologtostderr";
        const bool glog_val = bare == "stderrthreshold" || bare == "v" || bare == "log_dir" || bare == "vmodule" || bare == "log_backtrace_at";
        if (eq != std::string::npos) { val = name.substr(eq + 1); name = bare; }
        else if (glog_bool) val = "true";  // Go bool flags take no separate argument
        else if (i + 1 < argc) val = argv[++i];
        else { err = "flag needs an argument: -" + name; return false; }
        if (glog_bool || glog_val) continue;
        if (name == "pulse") f.pulse = atoi(val.c_str());
        else if (name == "resource_naming_strategy") f.strategy = val;
        else if (name == "backend") f.backend = val;
        else if (name == "plugin_dir") f.plugin_dir = val;
        else if (name == "resource_namespace") f.ns = val;
        else if (name == "start_retry_wait") f.start_retry_wait = atof(val.c_str());
        else if (name == "link_check") f.link_check = atoi(val.c_str());
        else if (name == "exporter_socket") f.exporter_socket = val;
        else { err = "flag provided but not defined: -" + name; return false; }
    }
    if (!f.plugin_dir.empty() && f.plugin_dir.back() != '/') f.plugin_dir += '/';
    return true;
}

void logf(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
void logf(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}

// One resource's plugin: plugin.go:41-48 AMDGPUPlugin + dpm/plugin.go devicePlugin
class Plugin {
public:
    Plugin(b2dp_ctx* ctx, const Flags& fl, std::string name) : ctx_(ctx), fl_(fl), name_(std::move(name)) {
        socket_ = fl_.plugin_dir + fl_.ns + "_" + name_;  // dpm/plugin.go:51-59
    }
    ~Plugin() { stop(); }

    bool start(std::string& err) {
        // plugin.go:82-91 Start(): allocator init; failure degrades to the kubelet's default allocation
        const int rc = b2dp_start(ctx_);
        alloc_ok_ = rc == B2DP_OK;
        if (!alloc_ok_) logf("allocator init failed (%s). Falling back to kubelet default allocation", b2dp_strerror(rc));
        server_ = std::make_unique<h2::GrpcServer>();
        server_->add_unary("/v1beta1.DevicePlugin/GetDevicePluginOptions", [this](const std::string&, std::string& out) {
            out = options();
            return h2::Status{};
        });
        server_->add_unary("/v1beta1.DevicePlugin/PreStartContainer", [](const std::string&, std::string& out) {
            out.clear();  // plugin.go:222-224
            return h2::Status{};
        });
        server_->add_unary("/v1beta1.DevicePlugin/GetPreferredAllocation",
                           [this](const std::string& in, std::string& out) { return preferred_allocation(in, out); });
        server_->add_unary("/v1beta1.DevicePlugin/Allocate", [this](const std::string& in, std::string& out) { return allocate(in, out); });
        server_->add_server_streaming("/v1beta1.DevicePlugin/ListAndWatch",
                                      [this](const std::string&, std::shared_ptr<h2::ServerStream> s) { list_and_watch(std::move(s)); });
        if (!server_->listen_unix(socket_, err)) { server_.reset(); return false; }
        return true;
    }

    // dpm/plugin.go:127-162 register()
    bool register_with_kubelet(std::string& err) {
        std::string opts, req;
        if (alloc_ok_) { b2dp::pb::tag(opts, 2, 0); b2dp::pb::varint(opts, 1); }  // get_preferred_allocation_available
        b2dp::pb::string_field(req, 1, "v1beta1");                                   // pluginapi.Version
        b2dp::pb::string_field(req, 2, fl_.ns + "_" + name_);                       // endpoint = socket basename
        b2dp::pb::string_field(req, 3, fl_.ns + "/" + name_);                       // resource name
        b2dp::pb::bytes_field(req, 4, opts);
        std::string resp;
        const h2::Status st = h2::unary_call_unix(fl_.plugin_dir + "kubelet.sock", "/v1beta1.Registration/Register", req, resp);
        if (st.code != h2::GRPC_OK) { err = "Register: grpc status " + std::to_string(st.code) + " " + st.message; return false; }
        return true;
    }

    // main.go:129-137: one tick of the shared heartbeat; like the reference's single unbuffered channel a tick
    // wakes ONE ListAndWatch stream (returns false if this plugin has none, so the caller can offer it elsewhere).
    bool beat() {
        std::lock_guard<std::mutex> l(mu_);
        reap_locked();
        if (watches_.empty()) return false;
        rr_ = (rr_ + 1) % watches_.size();
        b2dp_watch_beat(watches_[rr_]->w);
        return true;
    }

    void stop() {
        std::vector<std::unique_ptr<Watch>> ws;
        { std::lock_guard<std::mutex> l(mu_); ws.swap(watches_); }
        for (auto& w : ws) { b2dp_watch_stop(w->w); w->stream->finish({}); }  // returning from ListAndWatch ends the stream
        if (server_) { server_->stop(); server_.reset(); }
    }
    const std::string& name() const { return name_; }

private:
    struct Watch { b2dp_watch* w = nullptr; std::shared_ptr<h2::ServerStream> stream; std::atomic<bool> failed{false}; };

    std::string options() const {  // plugin.go:210-217
        std::string o;
        if (alloc_ok_) { b2dp::pb::tag(o, 2, 0); b2dp::pb::varint(o, 1); }
        return o;
    }

    // plugin.go:229-330: the library's watch loop produces the initial list and one response per heartbeat; this
    // side only forwards bytes to the stream.
    void list_and_watch(std::shared_ptr<h2::ServerStream> s) {
        auto w = std::make_unique<Watch>();
        w->stream = std::move(s);
        b2dp_cycle_opts o{};
        if (fl_.link_check) o.flags |= B2DP_LW_LINK_CHECK;
        Watch* raw = w.get();
        const int rc = b2dp_watch_start(ctx_, name_.c_str(), 0, &o, &Plugin::on_cycle, raw, &raw->w);
        if (rc != B2DP_OK) { raw->stream->finish({h2::GRPC_UNKNOWN, b2dp_strerror(rc)}); return; }
        std::lock_guard<std::mutex> l(mu_);
        reap_locked();
        watches_.push_back(std::move(w));
    }
    static void on_cycle(void* user, int rc, const uint8_t* buf, size_t len, const b2dp_cycle_stats*) {
        Watch* w = (Watch*)user;
        if (rc != B2DP_OK) { logf("ListAndWatch cycle failed: %s", b2dp_strerror(rc)); return; }
        if (!w->stream->send(std::string((const char*)buf, len))) w->failed = true;
    }
    void reap_locked() {  // streams the kubelet cancelled (restart, re-registration)
        for (size_t i = 0; i < watches_.size();) {
            if (watches_[i]->failed.load() || watches_[i]->stream->cancelled()) {
                std::unique_ptr<Watch> w = std::move(watches_[i]);
                watches_.erase(watches_.begin() + (long)i);
                b2dp_watch_stop(w->w);
            } else ++i;
        }
    }

    static bool read_strings(std::string_view msg, int want_field, std::vector<std::string>& out, int varint_field = 0, int64_t* v = nullptr) {
        pbread::Reader r(msg);
        int field, wire;
        while (!r.done()) {
            if (!r.tag(field, wire)) return false;
            if (wire == 2 && field == want_field) { std::string_view b; if (!r.bytes(b)) return false; out.emplace_back(b); }
            else if (wire == 0 && field == varint_field && v) { uint64_t x; if (!r.varint(x)) return false; *v = (int64_t)x; }
            else if (!r.skip(wire)) return false;
        }
        return true;
    }

    // plugin.go:337-351
    h2::Status preferred_allocation(const std::string& in, std::string& out) {
        pbread::Reader r(in);
        int field, wire;
        while (!r.done()) {
            if (!r.tag(field, wire)) return {h2::GRPC_INTERNAL, "bad PreferredAllocationRequest"};
            if (!(field == 1 && wire == 2)) { if (!r.skip(wire)) return {h2::GRPC_INTERNAL, "bad PreferredAllocationRequest"}; continue; }
            std::string_view creq;
            if (!r.bytes(creq)) return {h2::GRPC_INTERNAL, "bad PreferredAllocationRequest"};
            // ContainerPreferredAllocationRequest{available_deviceIDs=1, must_include_deviceIDs=2, allocation_size=3}
            std::vector<std::string> avail, must;
            int64_t size = 0;
            if (!read_strings(creq, 1, avail, 3, &size) || !read_strings(creq, 2, must)) return {h2::GRPC_INTERNAL, "bad ContainerPreferredAllocationRequest"};
            std::vector<const char*> pa, pm;
            for (auto& s : avail) pa.push_back(s.c_str());
            for (auto& s : must) pm.push_back(s.c_str());
            std::vector<char> ids((avail.size() + 1) * 64);
            int n = 0;
            const int rc = b2dp_preferred_allocation(ctx_, pa.data(), (int)pa.size(), pm.data(), (int)pm.size(), (int)(int32_t)size,
                                                     (char (*)[64])ids.data(), (int)avail.size() + 1, &n);
            if (rc != B2DP_OK)  // plugin.go:341-344: a Go error becomes status Unknown with its text
                return {h2::GRPC_UNKNOWN, std::string("unable to get preferred allocation list. Error:") + b2dp_strerror(rc)};
            std::string cresp;
            for (int i = 0; i < n; ++i) b2dp::pb::bytes_field(cresp, 1, std::string(&ids[(size_t)i * 64]));
            b2dp::pb::bytes_field(out, 1, cresp);
        }
        return {};
    }

    // plugin.go:356-393
    h2::Status allocate(const std::string& in, std::string& out) {
        pbread::Reader r(in);
        int field, wire;
        while (!r.done()) {
            if (!r.tag(field, wire)) return {h2::GRPC_INTERNAL, "bad AllocateRequest"};
            if (!(field == 1 && wire == 2)) { if (!r.skip(wire)) return {h2::GRPC_INTERNAL, "bad AllocateRequest"}; continue; }
            std::string_view creq;
            if (!r.bytes(creq)) return {h2::GRPC_INTERNAL, "bad AllocateRequest"};
            std::vector<std::string> ids;
            if (!read_strings(creq, 1, ids)) return {h2::GRPC_INTERNAL, "bad ContainerAllocateRequest"};
            std::vector<const char*> p;
            for (auto& s : ids) p.push_back(s.c_str());
            std::string buf(4096, '\0');
            size_t len = 0;
            int rc = b2dp_allocate_response(ctx_, p.data(), (int)p.size(), (uint8_t*)&buf[0], buf.size(), &len);
            if (rc == B2DP_E_NOSPC) { buf.assign(len, '\0'); rc = b2dp_allocate_response(ctx_, p.data(), (int)p.size(), (uint8_t*)&buf[0], buf.size(), &len); }
            if (rc != B2DP_OK) return {h2::GRPC_UNKNOWN, b2dp_strerror(rc)};
            b2dp::pb::bytes_field(out, 1, buf.substr(0, len));
        }
        return {};
    }

    b2dp_ctx* ctx_;
    const Flags& fl_;
    std::string name_, socket_;
    bool alloc_ok_ = false;
    std::unique_ptr<h2::GrpcServer> server_;
    std::mutex mu_;
    std::vector<std::unique_ptr<Watch>> watches_;
    size_t rr_ = 0;
};

// metricssvc.MetricsService/{List,GetGPUState} (internal/pkg/exporter/metricssvc/metricssvc.pb.go:95-110,284-291;
// metricssvc_grpc.pb.go:45-46) answered from the HBM probe, so that an UNMODIFIED reference plugin reads B200
// verdicts through its own exporter client (health.go:42-82): GPUState{ID=1, UUID=2, Health=3 "healthy"|"unhealthy"
// (health.go:75), AssociatedWorkload=4, Device=5 = the kubelet device id (health.go:98)}.  cuda: backend only.
class ExporterService {
public:
    ExporterService(b2dp_ctx* ctx, std::string path) : ctx_(ctx), path_(std::move(path)) {}
    bool start(std::string& err) {
        const size_t slash = path_.rfind('/');
        if (slash != std::string::npos && slash > 0) ::mkdir(path_.substr(0, slash).c_str(), 0755);
        server_.add_unary("/metricssvc.MetricsService/List", [this](const std::string&, std::string& out) { return states({}, out); });
        server_.add_unary("/metricssvc.MetricsService/GetGPUState", [this](const std::string& in, std::string& out) {
            std::vector<std::string> want;  // GPUGetRequest{ID=1 repeated string}
            pbread::Reader r(in);
            int field, wire;
            while (!r.done()) {
                if (!r.tag(field, wire)) return h2::Status{h2::GRPC_INTERNAL, "bad GPUGetRequest"};
                std::string_view b;
                if (field == 1 && wire == 2) { if (!r.bytes(b)) return h2::Status{h2::GRPC_INTERNAL, "bad GPUGetRequest"}; want.emplace_back(b); }
                else if (!r.skip(wire)) return h2::Status{h2::GRPC_INTERNAL, "bad GPUGetRequest"};
            }
            return states(want, out);
        });
        return server_.listen_unix(path_, err);
    }
    void stop() { server_.stop(); }

private:
    h2::Status states(const std::vector<std::string>& want, std::string& out) {
        std::vector<b2dp_device> devs(64);
        int n = 0;
        int rc = b2dp_enumerate(ctx_, devs.data(), (int)devs.size(), &n);
        if (rc == B2DP_E_NOSPC) { devs.resize((size_t)n); rc = b2dp_enumerate(ctx_, devs.data(), n, &n); }
        if (rc != B2DP_OK) return {h2::GRPC_UNKNOWN, b2dp_strerror(rc)};
        std::vector<b2dp_probe_result> res((size_t)n + 1);
        int m = 0;
        rc = b2dp_probe_health(ctx_, nullptr, res.data(), (int)res.size(), &m);
        if (rc != B2DP_OK) return {h2::GRPC_UNKNOWN, std::string(b2dp_strerror(rc)) + ": " + b2dp_last_error(ctx_)};
        for (int i = 0; i < m; ++i) {
            const int d = res[(size_t)i].device;
            if (d < 0 || d >= n) continue;
            const std::string id = std::to_string(d);
            if (!want.empty() && std::find(want.begin(), want.end(), id) == want.end()) continue;
            std::string g;
            b2dp::pb::string_field(g, 1, id);
            b2dp::pb::string_field(g, 2, devs[(size_t)d].id);
            b2dp::pb::string_field(g, 3, res[(size_t)i].healthy ? "healthy" : "unhealthy");
            b2dp::pb::string_field(g, 5, devs[(size_t)d].id);
            b2dp::pb::bytes_field(out, 1, g);
        }
        return {};
    }
    b2dp_ctx* ctx_;
    std::string path_;
    h2::GrpcServer server_;
};

bool sock_identity(const std::string& path, ino_t& ino, long long& ctime_ns) {
    struct stat st;
    if (::stat(path.c_str(), &st) != 0) return false;
    ino = st.st_ino;
    ctime_ns = (long long)st.st_ctim.tv_sec * 1000000000LL + st.st_ctim.tv_nsec;
    return true;
}

// Sleeps unless a stop signal arrives (signals are blocked and collected with sigtimedwait).
void sleep_interruptible(double seconds) {
    sigset_t stop;
    sigemptyset(&stop);
    for (int s : {SIGINT, SIGTERM, SIGQUIT}) sigaddset(&stop, s);
    const auto until = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
    while (!g_stop.load() && std::chrono::steady_clock::now() < until) {
        struct timespec ts{0, 20 * 1000 * 1000};
        if (sigtimedwait(&stop, nullptr, &ts) > 0) g_stop = true;
    }
}

}  // namespace

int main(int argc, char** argv) {
    Flags fl;
    std::string err;
    if (!parse_flags(argc, argv, fl, err)) { logf("%s", err.c_str()); return 2; }
    if (fl.version) {  // the reference stamps main.gitDescribe at link time (Makefile -X main.gitDescribe)
        printf("b200dp_plugind: libb200dp ABI %d (header %d), built %s\n", b2dp_abi_version(), B2DP_ABI_VERSION, __DATE__);
        return 0;
    }
    if (fl.strategy != "single" && fl.strategy != "mixed") {  // main.go:42-51,113-117
        logf("invalid resource naming strategy: %s", fl.strategy.c_str());
        return 1;
    }
    if (b2dp_abi_version() != B2DP_ABI_VERSION) { logf("libb200dp ABI %d, built against %d", b2dp_abi_version(), B2DP_ABI_VERSION); return 1; }
    // Signals are taken synchronously by this thread (sigtimedwait below): block them before any other thread
    // exists so every thread inherits the mask.  SIGINT/SIGQUIT/SIGTERM stop the daemon (plugin.go:83-84,
    // dpm/manager.go:47-48); SIGUSR1 is an operator's "heartbeat now".
    sigset_t sigs;
    sigemptyset(&sigs);
    for (int s : {SIGINT, SIGTERM, SIGQUIT, SIGUSR1}) sigaddset(&sigs, s);
    pthread_sigmask(SIG_BLOCK, &sigs, nullptr);
    signal(SIGPIPE, SIG_IGN);

    b2dp_ctx* ctx = nullptr;
    int rc = b2dp_open(fl.backend.c_str(), &ctx);
    if (rc != B2DP_OK) { logf("open %s: %s (%s)", fl.backend.c_str(), b2dp_strerror(rc), b2dp_last_error(nullptr)); return 1; }
    char names[64][64];
    int n_res = 0;
    rc = b2dp_resource_list(ctx, fl.strategy.c_str(), names, 64, &n_res);  // main.go:141-146
    if (rc != B2DP_OK) { logf("Error occured: %s", b2dp_strerror(rc)); b2dp_close(ctx); return 1; }

    std::unique_ptr<ExporterService> exporter;
    if (!fl.exporter_socket.empty()) {
        exporter = std::make_unique<ExporterService>(ctx, fl.exporter_socket);
        if (!exporter->start(err)) { logf("exporter socket: %s", err.c_str()); b2dp_close(ctx); return 1; }
        logf("metricssvc.MetricsService on %s", fl.exporter_socket.c_str());
    }
    std::vector<std::unique_ptr<Plugin>> plugins;
    auto start_all = [&] {  // dpm handleNewPlugins + startPlugin
        plugins.clear();
        for (int i = 0; i < n_res && !g_stop.load(); ++i) {
            std::string last;
            for (int attempt = 0; attempt < 3 && !g_stop.load(); ++attempt) {  // dpm/manager.go:16-20,205-219
                auto p = std::make_unique<Plugin>(ctx, fl, names[i]);
                if (p->start(last) && p->register_with_kubelet(last)) {
                    logf("%s/%s: serving on %s%s_%s, registered", fl.ns.c_str(), names[i], fl.plugin_dir.c_str(), fl.ns.c_str(), names[i]);
                    plugins.push_back(std::move(p));
                    last.clear();
                    break;
                }
                p.reset();
                sleep_interruptible(fl.start_retry_wait);
            }
            if (!last.empty()) logf("Failed to start plugin %s: %s", names[i], last.c_str());
        }
    };
    start_all();

    const std::string kubelet_sock = fl.plugin_dir + "kubelet.sock";
    ino_t seen_ino = 0;
    long long seen_ctime = 0;
    bool have_seen = sock_identity(kubelet_sock, seen_ino, seen_ctime);
    auto next_beat = std::chrono::steady_clock::now() + std::chrono::seconds(fl.pulse > 0 ? fl.pulse : 0);
    auto next_check = std::chrono::steady_clock::now() + std::chrono::seconds(1);
    size_t beat_rr = 0;
    auto heartbeat = [&] {  // main.go:129-137: `l.Heartbeat <- true` reaches one receiver
        for (size_t k = 0; k < plugins.size(); ++k) {
            beat_rr = (beat_rr + 1) % plugins.size();
            if (plugins[beat_rr]->beat()) break;
        }
    };
    while (!g_stop.load()) {
        struct timespec ts{0, 20 * 1000 * 1000};
        const int sig = sigtimedwait(&sigs, nullptr, &ts);
        if (sig == SIGUSR1) heartbeat();
        else if (sig > 0) { g_stop = true; break; }
        const auto now = std::chrono::steady_clock::now();
        if (fl.pulse > 0 && now >= next_beat) {
            next_beat += std::chrono::seconds(fl.pulse);
            heartbeat();
        }
        if (now >= next_check) {  // dpm/manager.go:73-84: the kubelet restarted -> serve again and re-register
            next_check = now + std::chrono::seconds(1);
            ino_t ino = 0;
            long long ct = 0;
            if (sock_identity(kubelet_sock, ino, ct)) {
                if (have_seen && (ino != seen_ino || ct != seen_ctime)) { logf("kubelet.sock re-created: restarting plugins"); start_all(); }
                else if (!have_seen && plugins.empty()) start_all();
                seen_ino = ino; seen_ctime = ct; have_seen = true;
            }
        }
    }
    logf("Received signal, exiting");
    plugins.clear();
    if (exporter) exporter->stop();
    b2dp_close(ctx);
    return 0;
}
