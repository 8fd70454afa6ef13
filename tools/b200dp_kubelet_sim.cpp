// b200dp_kubelet_sim -- a kubelet stand-in written on the same native gRPC (csrc/host/h2grpc.hpp): serves
// v1beta1.Registration/Register on <dir>/kubelet.sock, starts b200dp_plugind, opens ListAndWatch on the socket the
// daemon registered, then times SIGUSR1 ("heartbeat now") -> next ListAndWatchResponse received, N times.  With no
// Python and no third-party gRPC on either side this is the floor of the heartbeat -> kubelet latency.
//   b200dp_kubelet_sim <path to b200dp_plugind> <backend uri> [iterations]
#include <signal.h>
#include <sys/wait.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../k8s-device-plugin_b200/csrc/host/h2grpc.hpp"
#include "../k8s-device-plugin_b200/csrc/host/pbread.hpp"

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: b200dp_kubelet_sim <b200dp_plugind> <backend uri> [iterations]\n"); return 2; }
    const int iters = argc > 3 ? atoi(argv[3]) : 200;
    char tmpl[] = "/tmp/b2k_XXXXXX";
    if (!mkdtemp(tmpl)) { perror("mkdtemp"); return 1; }
    const std::string dir = std::string(tmpl) + "/";
    std::mutex mu;
    std::condition_variable cv;
    std::string endpoint, resource;
    h2::GrpcServer kubelet;
    kubelet.add_unary("/v1beta1.Registration/Register", [&](const std::string& req, std::string& out) {
        pbread::Reader r(req);  // RegisterRequest{version=1, endpoint=2, resource_name=3, options=4}
        int field, wire;
        std::string ep, rn;
        while (!r.done() && r.tag(field, wire)) {
            std::string_view b;
            if (wire == 2 && r.bytes(b)) { if (field == 2) ep = std::string(b); else if (field == 3) rn = std::string(b); }
            else if (wire != 2 && !r.skip(wire)) break;
        }
        { std::lock_guard<std::mutex> l(mu); endpoint = ep; resource = rn; }
        cv.notify_all();
        out.clear();
        return h2::Status{};
    });
    std::string err;
    if (!kubelet.listen_unix(dir + "kubelet.sock", err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }

    const pid_t pid = fork();
    if (pid == 0) {
        const std::string b = std::string("-backend=") + argv[2];
        execl(argv[1], argv[1], "-pulse=0", b.c_str(), "-plugin_dir", dir.c_str(), (char*)nullptr);
        _exit(127);
    }
    int rc = 1;
    {
        std::unique_lock<std::mutex> l(mu);
        if (!cv.wait_for(l, std::chrono::seconds(120), [&] { return !endpoint.empty(); })) { fprintf(stderr, "the daemon did not register\n"); goto done; }
    }
    {
        h2::ClientStream s;
        h2::Status st = s.open(dir + endpoint, "/v1beta1.DevicePlugin/ListAndWatch", "");
        std::string msg;
        if (st.code != 0 || !s.next(msg, 60000)) { fprintf(stderr, "ListAndWatch: %s %s\n", st.message.c_str(), s.status().message.c_str()); goto done; }
        const size_t first_len = msg.size();
        std::vector<double> lat;
        for (int i = -5; i < iters; ++i) {
            const double t0 = now_ms();
            kill(pid, SIGUSR1);
            if (!s.next(msg, 10000)) { fprintf(stderr, "stream ended: %s\n", s.status().message.c_str()); goto done; }
            if (i >= 0) lat.push_back(now_ms() - t0);
        }
        std::sort(lat.begin(), lat.end());
        printf("{\"resource\": \"%s\", \"iterations\": %d, \"response_bytes\": %zu, \"heartbeat_to_kubelet_ms_median\": %.4f, "
               "\"heartbeat_to_kubelet_ms_p99\": %.4f, \"heartbeat_to_kubelet_ms_min\": %.4f}\n",
               resource.c_str(), iters, first_len, lat[lat.size() / 2], lat[std::min(lat.size() - 1, lat.size() * 99 / 100)], lat[0]);
        s.close();
        rc = 0;
    }
done:
    kill(pid, SIGTERM);
    int wst = 0;
    waitpid(pid, &wst, 0);
    kubelet.stop();
    ::rmdir((dir + ".").c_str());
    ::rmdir(tmpl);
    return rc;
}
