// helper_proto.hpp -- wire format between libb200dp (parent) and b200dp_probe_helper (one child per probed unit).
//
// Why helper processes: CUDA exposes at most ONE MIG compute instance to a process (and hides every other GPU while it
// does), so the in-process fan-out cannot reach the instances of a MIG-partitioned B200.  The parent therefore never
// touches CUDA in this mode; each unit (MIG instance, or a whole GPU with probe=helpers) gets a child started with
// CUDA_VISIBLE_DEVICES=<its UUID> that opens the ordinary in-process backend on "device 0" and answers fixed-size
// requests over a unix socketpair (stdin/stdout of the child).
#pragma once
#include <cstdint>

#include "../../include/b200dp.h"

namespace b2dp {

constexpr uint32_t kHelperMagic = 0xB2D90003u;  // changes with the layouts below

enum HelperOp : uint32_t {
    HOP_HELLO = 0,   // -> text = product name, extra = {SM count, total memory bytes, ring slot bytes, calibrated ceiling as float bits}
    HOP_PROBE = 1,   // opts -> res
    HOP_INJECT = 2,  // a = word index, b = mask
    HOP_RESET = 3,
    HOP_PEEK = 4,    // a = word index, b = n words (<= 1 Mi); the response is followed by b*4 bytes
    HOP_SETREF = 5,  // a = float bits of gbs_ref
    HOP_QUIT = 6,
};

struct HelperReq {
    uint32_t magic;
    uint32_t op;
    uint64_t seq;
    uint64_t a, b;
    b2dp_probe_opts opts;
};

struct HelperRsp {
    uint32_t magic;
    int32_t rc;  // B2DP_OK or a negative B2DP_E_* from the child's own ABI call
    uint64_t seq;
    uint64_t extra[4];
    b2dp_probe_result res;
    char text[128];  // HELLO: product name; on failure: b2dp_last_error()
};

}  // namespace b2dp
