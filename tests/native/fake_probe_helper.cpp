// Test-only stand-in for b200dp_probe_helper: speaks csrc/helper_proto.hpp on stdin/stdout and fabricates the answers a
// healthy unit would give (seed schedule of its enumeration index, the host closed-form checksum, a plausible rate), so
// that the PARENT side of probe=helpers -- spawning one child per (MIG) unit, the fan-out, deadlines, stale answers,
// dead children and their restart -- can be exercised on a box without a GPU.  Selected with B2DP_PROBE_HELPER.
//
// Behaviour knobs (environment, all optional):
//   FAKE_HELPER_SLOW_UNIT=<i> FAKE_HELPER_SLOW_MS=<ms>   unit i answers every PROBE after sleeping ms
//   FAKE_HELPER_DIE_UNIT=<i>  FAKE_HELPER_DIE_AFTER=<n>  unit i exits instead of answering its (n+1)-th PROBE -- once: a
//                                                        restarted child (FAKE_HELPER_MARK file exists) lives on
//   FAKE_HELPER_RESTART_HELLO_MS=<ms>                    a restarted child takes this long to answer HELLO
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../k8s-device-plugin_b200/csrc/helper_proto.hpp"
#include "../../k8s-device-plugin_b200/csrc/pattern_math.hpp"

using namespace b2dp;

static bool read_all(int fd, void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n) { const ssize_t r = read(fd, c, n); if (r == 0) return false; if (r < 0) { if (errno == EINTR) continue; return false; } c += r; n -= (size_t)r; }
    return true;
}
static bool write_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) { const ssize_t r = write(fd, c, n); if (r < 0) { if (errno == EINTR) continue; return false; } c += r; n -= (size_t)r; }
    return true;
}
static long long uri_num(const std::string& uri, const char* key, long long dflt) {
    const size_t p = uri.find(std::string(key) + "=");
    return p == std::string::npos ? dflt : atoll(uri.c_str() + p + strlen(key) + 1);
}
static int env_int(const char* k, int d) { const char* v = getenv(k); return v && *v ? atoi(v) : d; }

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    const std::string uri = argv[1];
    const int unit = (int)uri_num(uri, "seed_index", 0);
    const unsigned long long bytes = (unsigned long long)uri_num(uri, "bytes", 1 << 20);
    const char* vis = getenv("CUDA_VISIBLE_DEVICES");  // the parent must have named this unit's UUID
    uint32_t seed = 0x5EED0000u | (uint32_t)(unit & 0xffff);
    float gbs_ref = 900.f;
    bool fault = false;
    unsigned long long fault_word = 0;
    int probes = 0;
    const char* mark = getenv("FAKE_HELPER_MARK");
    const bool restarted = mark && access(mark, F_OK) == 0;
    for (;;) {
        HelperReq q{};
        if (!read_all(0, &q, sizeof q) || q.magic != kHelperMagic || q.op == HOP_QUIT) break;
        HelperRsp r{};
        r.magic = kHelperMagic;
        r.seq = q.seq;
        switch (q.op) {
            case HOP_HELLO:
                if (restarted) usleep(1000u * (unsigned)env_int("FAKE_HELPER_RESTART_HELLO_MS", 0));  // a slow-starting (CUDA-like) restart
                snprintf(r.text, sizeof r.text, "FAKE B200 %s", vis ? vis : "?");
                r.extra[0] = 18; r.extra[1] = 23ull << 30; r.extra[2] = bytes;
                memcpy(&r.extra[3], &gbs_ref, sizeof gbs_ref);
                break;
            case HOP_PROBE: {
                if (unit == env_int("FAKE_HELPER_DIE_UNIT", -1) && !restarted && probes >= env_int("FAKE_HELPER_DIE_AFTER", 1 << 30)) {
                    if (mark) { FILE* f = fopen(mark, "w"); if (f) fclose(f); }
                    _exit(7);
                }
                if (unit == env_int("FAKE_HELPER_SLOW_UNIT", -1)) usleep(1000u * (unsigned)env_int("FAKE_HELPER_SLOW_MS", 0));
                ++probes;
                b2dp_probe_result& o = r.res;
                o.seed = seed;
                o.bytes = 2 * bytes;
                o.expected_checksum = expected_checksum_host(bytes / 4, seed);
                o.checksum = o.expected_checksum ^ (fault ? 1ull : 0ull);
                o.mismatches = fault ? 1 : 0;
                o.first_bad_word = fault ? fault_word : ~0ull;
                o.ms_device = 1.0f;
                o.gbs = 850.f;
                o.gbs_ref = gbs_ref;
                o.frac = o.gbs / gbs_ref;
                o.min_gbs_applied = q.opts.min_gbs > 0 ? q.opts.min_gbs : 0.8f * gbs_ref;
                const bool fast = o.gbs >= o.min_gbs_applied;
                if (!fast) o.flags |= B2DP_RES_SLOW;
                o.healthy = (!fault && fast) ? 1 : 0;
                fault = false;
                seed = seed * 1664525u + 1013904223u;
                break;
            }
            case HOP_INJECT: fault = true; fault_word = q.a; break;
            case HOP_RESET: fault = false; break;
            case HOP_SETREF: memcpy(&gbs_ref, &q.a, sizeof gbs_ref); break;
            default: r.rc = B2DP_E_INVAL;
        }
        if (!write_all(1, &r, sizeof r)) break;
    }
    return 0;
}
