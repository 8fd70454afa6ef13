// b200dp_probe_helper -- one child process per probed unit (helper_proto.hpp).
//
//   CUDA_VISIBLE_DEVICES=<GPU-uuid | MIG-uuid>  b200dp_probe_helper "cuda:devices=0,bytes=...,seed_index=<i>,..."
//
// The parent (libb200dp in probe=helpers mode, or on any node with a MIG-enabled GPU) owns enumeration, health merge,
// allocation and labels; this process owns ONE CUDA context on ONE device and runs the ordinary in-process probe on
// it through the public C ABI.  It exits when its stdin closes (parent gone) or on HOP_QUIT.
#include <unistd.h>

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../helper_proto.hpp"

using namespace b2dp;

static bool read_all(int fd, void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n) {
        const ssize_t r = read(fd, c, n);
        if (r == 0) return false;
        if (r < 0) { if (errno == EINTR) continue; return false; }
        c += r; n -= (size_t)r;
    }
    return true;
}
static bool write_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        const ssize_t r = write(fd, c, n);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        c += r; n -= (size_t)r;
    }
    return true;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s <cuda: backend uri>   (speaks helper_proto.hpp on stdin/stdout)\n", argv[0]); return 2; }
    b2dp_ctx* ctx = nullptr;
    const int open_rc = b2dp_open(argv[1], &ctx);
    char open_err[128] = {0};
    if (open_rc != B2DP_OK) snprintf(open_err, sizeof open_err, "%s", b2dp_last_error(nullptr));
    std::vector<uint32_t> words;
    for (;;) {
        HelperReq q{};
        if (!read_all(0, &q, sizeof q) || q.magic != kHelperMagic) break;
        HelperRsp r{};
        r.magic = kHelperMagic;
        r.seq = q.seq;
        size_t payload = 0;
        if (q.op == HOP_QUIT) break;
        if (open_rc != B2DP_OK) {  // the unit could not be set up: every request says why
            r.rc = open_rc;
            memcpy(r.text, open_err, sizeof open_err);
        } else switch (q.op) {
            case HOP_HELLO: {
                b2dp_probe_info pi{};
                r.rc = b2dp_probe_describe(ctx, 0, &pi);
                snprintf(r.text, sizeof r.text, "%s", pi.name);
                r.extra[0] = (uint64_t)pi.sm_count;
                r.extra[1] = pi.total_memory;
                r.extra[2] = pi.slot_bytes;
                memcpy(&r.extra[3], &pi.gbs_cal, sizeof(float));
                break;
            }
            case HOP_PROBE: {
                int n = 0;
                r.rc = b2dp_probe_health(ctx, &q.opts, &r.res, 1, &n);
                break;
            }
            case HOP_INJECT: r.rc = b2dp_probe_inject_fault(ctx, 0, q.a, (uint32_t)q.b); break;
            case HOP_RESET: r.rc = b2dp_probe_reset(ctx, 0); break;
            case HOP_SETREF: { float f; memcpy(&f, &q.a, sizeof f); r.rc = b2dp_probe_set_ref(ctx, 0, f); break; }
            case HOP_PEEK:
                if (q.b > (1u << 20)) { r.rc = B2DP_E_INVAL; break; }
                words.assign((size_t)q.b, 0);
                r.rc = b2dp_probe_peek(ctx, 0, q.a, words.data(), q.b);
                if (r.rc == B2DP_OK) payload = (size_t)q.b * 4;
                break;
            default: r.rc = B2DP_E_INVAL;
        }
        if (r.rc != B2DP_OK && open_rc == B2DP_OK) snprintf(r.text, sizeof r.text, "%s", b2dp_last_error(ctx));
        if (!write_all(1, &r, sizeof r)) break;
        if (payload && !write_all(1, words.data(), payload)) break;
    }
    if (ctx) b2dp_close(ctx);
    return 0;
}
