// pbread.hpp -- minimal protobuf wire reader for the three request messages the kubelet sends
// (PreferredAllocationRequest, AllocateRequest, PreStartContainerRequest; api.proto:120-223).
#pragma once
#include <cstdint>
#include <string>
#include <string_view>

namespace pbread {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    explicit Reader(std::string_view s) : p((const uint8_t*)s.data()), end((const uint8_t*)s.data() + s.size()) {}
    bool done() const { return p >= end; }
    bool varint(uint64_t& v) {
        v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) return false;
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return true;
        }
        return false;
    }
    bool tag(int& field, int& wire) {
        uint64_t t;
        if (!varint(t)) return false;
        field = (int)(t >> 3);
        wire = (int)(t & 7);
        return field > 0;
    }
    bool bytes(std::string_view& out) {
        uint64_t n;
        if (!varint(n) || n > (uint64_t)(end - p)) return false;
        out = std::string_view((const char*)p, (size_t)n);
        p += n;
        return true;
    }
    bool skip(int wire) {
        uint64_t v;
        std::string_view b;
        switch (wire) {
        case 0: return varint(v);
        case 1: if (end - p < 8) return false; p += 8; return true;
        case 2: return bytes(b);
        case 5: if (end - p < 4) return false; p += 4; return true;
        default: return false;
        }
    }
};

}  // namespace pbread
