// RFC 7541 Appendix C.4 (requests with Huffman coding, one connection => one dynamic table) and C.6 (responses
// with Huffman coding and evictions at SETTINGS_HEADER_TABLE_SIZE = 256 are not used: we never lower the table)
// through csrc/host/h2grpc.hpp's decoder, plus encoder -> decoder round trips and gRPC framing helpers.
#include <cstdio>
#include <string>
#include <vector>

#include "../../k8s-device-plugin_b200/csrc/host/h2grpc.hpp"

static int failed = 0;
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++failed; } } while (0)

static std::string unhex(const char* s) {
    std::string o;
    int hi = -1;
    for (; *s; ++s) {
        int v = *s >= '0' && *s <= '9' ? *s - '0' : *s >= 'a' && *s <= 'f' ? *s - 'a' + 10 : -1;
        if (v < 0) continue;
        if (hi < 0) hi = v; else { o.push_back((char)(hi * 16 + v)); hi = -1; }
    }
    return o;
}
static bool eq(const h2::Headers& h, std::initializer_list<std::pair<const char*, const char*>> want) {
    if (h.size() != want.size()) return false;
    size_t i = 0;
    for (auto& w : want) { if (h[i].name != w.first || h[i].value != w.second) return false; ++i; }
    return true;
}

int main() {
    h2::HpackDecoder d;
    h2::Headers h;
    std::string b = unhex("8286 8441 8cf1 e3c2 e5f2 3a6b a0ab 90f4 ff");  // C.4.1
    CHECK(d.decode((const uint8_t*)b.data(), b.size(), h));
    CHECK(eq(h, {{":method", "GET"}, {":scheme", "http"}, {":path", "/"}, {":authority", "www.example.com"}}));
    h.clear();
    b = unhex("8286 84be 5886 a8eb 1064 9cbf");  // C.4.2: :authority now comes from the dynamic table (index 62)
    CHECK(d.decode((const uint8_t*)b.data(), b.size(), h));
    CHECK(eq(h, {{":method", "GET"}, {":scheme", "http"}, {":path", "/"}, {":authority", "www.example.com"}, {"cache-control", "no-cache"}}));
    h.clear();
    b = unhex("8287 85bf 4088 25a8 49e9 5ba9 7d7f 8925 a849 e95b b8e8 b4bf");  // C.4.3
    CHECK(d.decode((const uint8_t*)b.data(), b.size(), h));
    CHECK(eq(h, {{":method", "GET"}, {":scheme", "https"}, {":path", "/index.html"}, {":authority", "www.example.com"}, {"custom-key", "custom-value"}}));

    // C.2.1 literal with indexing, no Huffman; C.2.2 literal without indexing; C.2.3 never indexed
    h2::HpackDecoder d2;
    h.clear();
    b = unhex("400a 6375 7374 6f6d 2d6b 6579 0d63 7573 746f 6d2d 6865 6164 6572");
    CHECK(d2.decode((const uint8_t*)b.data(), b.size(), h) && eq(h, {{"custom-key", "custom-header"}}));
    h.clear();
    b = unhex("040c 2f73 616d 706c 652f 7061 7468");
    CHECK(d2.decode((const uint8_t*)b.data(), b.size(), h) && eq(h, {{":path", "/sample/path"}}));
    h.clear();
    b = unhex("1008 7061 7373 776f 7264 0673 6563 7265 74");
    CHECK(d2.decode((const uint8_t*)b.data(), b.size(), h) && eq(h, {{"password", "secret"}}));
    h.clear();
    b = unhex("be");  // the entry C.2.1 added
    CHECK(d2.decode((const uint8_t*)b.data(), b.size(), h) && eq(h, {{"custom-key", "custom-header"}}));

    // C.1 integer representation: 10 in a 5-bit prefix, 1337 in a 5-bit prefix, 42 in an 8-bit prefix
    { std::string o; h2::hpack_put_int(o, 0, 5, 10); CHECK(o == unhex("0a")); }
    { std::string o; h2::hpack_put_int(o, 0, 5, 1337); CHECK(o == unhex("1f9a0a")); }
    { const std::string s = unhex("1f9a0a"); const uint8_t* p = (const uint8_t*)s.data(); uint64_t v = 0;
      CHECK(h2::HpackDecoder::read_int(p, p + s.size(), 5, v) && v == 1337); }

    // errors: index 0, index past both tables, truncated string, EOS / bad padding in a Huffman string
    for (const char* bad : {"80", "ff ff ff ff ff ff ff ff ff ff ff ff", "c0", "0005 6162", "0081 00 00", "0082 ffff ffff 00"}) {
        h2::HpackDecoder dx;
        h2::Headers hx;
        const std::string s = unhex(bad);
        CHECK(!dx.decode((const uint8_t*)s.data(), s.size(), hx));
    }

    // our encoder -> our decoder, including a 300-byte value (multi-byte length)
    h2::Headers out = {{":status", "200"}, {"content-type", "application/grpc"}, {"grpc-message", std::string(300, 'x')}}, back;
    std::string enc;
    h2::hpack_encode(enc, out);
    h2::HpackDecoder d3;
    CHECK(d3.decode((const uint8_t*)enc.data(), enc.size(), back) && back.size() == 3 && back[2].value == out[2].value && back[1].name == "content-type");

    // gRPC helpers
    std::vector<std::string> msgs;
    CHECK(h2::grpc_unframe(h2::grpc_frame("abc") + h2::grpc_frame(""), msgs) && msgs.size() == 2 && msgs[0] == "abc" && msgs[1].empty());
    msgs.clear();
    CHECK(!h2::grpc_unframe(std::string("\x01\x00\x00\x00\x01x", 6), msgs));  // compressed flag
    CHECK(!h2::grpc_unframe(std::string("\x00\x00\x00\x00\x05x", 6), msgs));  // truncated
    CHECK(h2::percent_encode("a b%\n\xc3\xa9") == "a b%25%0A%C3%A9" && h2::percent_decode("a b%25%0A%C3%A9") == "a b%\n\xc3\xa9");
    // flow-control bookkeeping: WINDOW_UPDATE for streams that are not open (closed long ago, or never opened by a
    // confused or hostile peer) must not create entries nothing ever erases; open streams are tracked until forgotten
    {
        int sv[2];
        CHECK(socketpair(AF_UNIX, SOCK_STREAM, 0, sv) == 0);
        h2::Conn c(sv[0]);
        h2::Frame f;
        f.type = h2::F_WINDOW_UPDATE; f.flags = 0;
        f.payload = std::string("\x00\x00\x10\x00", 4);
        for (uint32_t sid = 1; sid < 20001; sid += 2) { f.stream = sid; CHECK(c.handle_control(f)); }
        CHECK(c.tracked_streams() == 0);
        c.open_stream(7);
        f.stream = 7;
        CHECK(c.handle_control(f) && c.tracked_streams() == 1);
        f.stream = 0;                       // connection-level credit is always taken
        CHECK(c.handle_control(f) && c.tracked_streams() == 1);
        c.forget_stream(7);
        f.stream = 7;
        CHECK(c.handle_control(f) && c.tracked_streams() == 0);
        f.payload = "abc";                  // malformed: connection error
        CHECK(!c.handle_control(f));
        // a peer that never reads: the send times out (SO_SNDTIMEO set by Conn) instead of blocking for ever -- only
        // checked for being configured, a 10 s stall has no place in a unit test
        struct timeval tv{};
        socklen_t len = sizeof tv;
        CHECK(getsockopt(sv[0], SOL_SOCKET, SO_SNDTIMEO, &tv, &len) == 0 && tv.tv_sec == 10);
        close(sv[1]);
    }
    printf("%s\n", failed ? "FAIL" : "PASS");
    return failed ? 1 : 0;
}
