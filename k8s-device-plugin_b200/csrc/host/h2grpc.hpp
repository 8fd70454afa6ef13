// h2grpc.hpp -- the small part of gRPC-over-HTTP/2 that the kubelet device-plugin contract needs, so that the
// plugin daemon (plugind.cpp) is one native binary: a unix-socket gRPC server (unary + server-streaming methods)
// for v1beta1.DevicePlugin and a unary client call for v1beta1.Registration/Register.  The reference gets this
// from google.golang.org/grpc through dpm (vendor/github.com/kubevirt/device-plugin-manager/pkg/dpm/plugin.go:
// 93-123 serve, :127-162 register); no gRPC C++ library exists in the build image.
//
// Implemented from the RFCs: HTTP/2 framing, SETTINGS/PING/WINDOW_UPDATE/RST_STREAM/GOAWAY handling and both
// directions of flow control (RFC 9113); HPACK decoding with the dynamic table and Huffman strings, encoding as
// literals without indexing (RFC 7541); gRPC's length-prefixed messages, `grpc-status` / `grpc-message` trailers
// and trailers-only errors (gRPC over HTTP/2 protocol).  No TLS (the kubelet sockets are plaintext), no
// compression, no client streaming.
#pragma once
#include <poll.h>
#include <sys/socket.h>
#include <sys/time.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace h2 {

struct Header { std::string name, value; };
using Headers = std::vector<Header>;

// ---- HPACK (RFC 7541) ------------------------------------------------------------------------------------
struct HuffSym { uint32_t code; uint8_t len; };
static const HuffSym kHuffman[257] = {
#include "../hpack_huffman.inc"
};

class HuffmanTree {
public:
    HuffmanTree() {
        nodes_.push_back(Node{});
        for (int s = 0; s < 257; ++s) {
            int cur = 0;
            for (int b = kHuffman[s].len - 1; b >= 0; --b) {
                const int bit = (kHuffman[s].code >> b) & 1;
                if (nodes_[(size_t)cur].child[bit] < 0) { nodes_[(size_t)cur].child[bit] = (int)nodes_.size(); nodes_.push_back(Node{}); }
                cur = nodes_[(size_t)cur].child[bit];
            }
            nodes_[(size_t)cur].sym = s;
        }
    }
    bool decode(const uint8_t* p, size_t n, std::string& out) const {
        int cur = 0, depth = 0;
        bool all_ones = true;
        for (size_t i = 0; i < n; ++i)
            for (int b = 7; b >= 0; --b) {
                const int bit = (p[i] >> b) & 1;
                cur = nodes_[(size_t)cur].child[bit];
                if (cur < 0) return false;
                ++depth;
                all_ones = all_ones && bit;
                if (nodes_[(size_t)cur].sym >= 0) {
                    if (nodes_[(size_t)cur].sym == 256) return false;  // EOS inside a string is an error (RFC 7541 5.2)
                    out.push_back((char)nodes_[(size_t)cur].sym);
                    cur = 0; depth = 0; all_ones = true;
                }
            }
        return depth < 8 && all_ones;  // padding: fewer than 8 bits, all ones
    }

private:
    struct Node { int child[2] = {-1, -1}; int sym = -1; };
    std::vector<Node> nodes_;
};
inline const HuffmanTree& huffman() { static const HuffmanTree t; return t; }

static const Header kStaticTable[61] = {  // RFC 7541 Appendix A
    {":authority", ""}, {":method", "GET"}, {":method", "POST"}, {":path", "/"}, {":path", "/index.html"},
    {":scheme", "http"}, {":scheme", "https"}, {":status", "200"}, {":status", "204"}, {":status", "206"},
    {":status", "304"}, {":status", "400"}, {":status", "404"}, {":status", "500"}, {"accept-charset", ""},
    {"accept-encoding", "gzip, deflate"}, {"accept-language", ""}, {"accept-ranges", ""}, {"accept", ""},
    {"access-control-allow-origin", ""}, {"age", ""}, {"allow", ""}, {"authorization", ""}, {"cache-control", ""},
    {"content-disposition", ""}, {"content-encoding", ""}, {"content-language", ""}, {"content-length", ""},
    {"content-location", ""}, {"content-range", ""}, {"content-type", ""}, {"cookie", ""}, {"date", ""}, {"etag", ""},
    {"expect", ""}, {"expires", ""}, {"from", ""}, {"host", ""}, {"if-match", ""}, {"if-modified-since", ""},
    {"if-none-match", ""}, {"if-range", ""}, {"if-unmodified-since", ""}, {"last-modified", ""}, {"link", ""},
    {"location", ""}, {"max-forwards", ""}, {"proxy-authenticate", ""}, {"proxy-authorization", ""}, {"range", ""},
    {"referer", ""}, {"refresh", ""}, {"retry-after", ""}, {"server", ""}, {"set-cookie", ""},
    {"strict-transport-security", ""}, {"transfer-encoding", ""}, {"user-agent", ""}, {"vary", ""}, {"via", ""},
    {"www-authenticate", ""}};

class HpackDecoder {
public:
    // Decodes one complete header block.  false = COMPRESSION_ERROR (the connection must be torn down).
    bool decode(const uint8_t* p, size_t n, Headers& out) {
        const uint8_t* end = p + n;
        while (p < end) {
            const uint8_t b = *p;
            uint64_t idx = 0;
            if (b & 0x80) {  // indexed header field
                if (!read_int(p, end, 7, idx) || idx == 0) return false;
                Header h;
                if (!lookup(idx, h)) return false;
                out.push_back(std::move(h));
            } else if ((b & 0xc0) == 0x40) {  // literal with incremental indexing
                Header h;
                if (!read_int(p, end, 6, idx) || !literal(p, end, idx, h)) return false;
                add(h);
                out.push_back(std::move(h));
            } else if ((b & 0xe0) == 0x20) {  // dynamic table size update
                if (!read_int(p, end, 5, idx) || idx > settings_max_) return false;
                max_size_ = (size_t)idx;
                evict();
            } else {  // literal without indexing (0000) / never indexed (0001)
                Header h;
                if (!read_int(p, end, 4, idx) || !literal(p, end, idx, h)) return false;
                out.push_back(std::move(h));
            }
        }
        return true;
    }

    static bool read_int(const uint8_t*& p, const uint8_t* end, int prefix_bits, uint64_t& v) {
        if (p >= end) return false;
        const uint64_t max = (1u << prefix_bits) - 1;
        v = *p++ & max;
        if (v < max) return true;
        for (int shift = 0;; shift += 7) {
            if (p >= end || shift > 56) return false;
            const uint8_t b = *p++;
            v += (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return true;
        }
    }
    static bool read_string(const uint8_t*& p, const uint8_t* end, std::string& s) {
        if (p >= end) return false;
        const bool huff = *p & 0x80;
        uint64_t len = 0;
        if (!read_int(p, end, 7, len) || len > (uint64_t)(end - p)) return false;
        s.clear();
        if (huff) { if (!huffman().decode(p, (size_t)len, s)) return false; }
        else s.assign((const char*)p, (size_t)len);
        p += len;
        return true;
    }

private:
    bool lookup(uint64_t idx, Header& h) const {
        if (idx >= 1 && idx <= 61) { h = kStaticTable[idx - 1]; return true; }
        const uint64_t d = idx - 62;
        if (d >= dyn_.size()) return false;
        h = dyn_[(size_t)d];
        return true;
    }
    bool literal(const uint8_t*& p, const uint8_t* end, uint64_t name_idx, Header& h) {
        if (name_idx) {
            Header n;
            if (!lookup(name_idx, n)) return false;
            h.name = n.name;
        } else if (!read_string(p, end, h.name)) return false;
        return read_string(p, end, h.value);
    }
    void add(const Header& h) {
        const size_t sz = h.name.size() + h.value.size() + 32;
        dyn_.push_front(h);
        dyn_size_ += sz;
        evict();
    }
    void evict() {
        while (dyn_size_ > max_size_ && !dyn_.empty()) {
            dyn_size_ -= dyn_.back().name.size() + dyn_.back().value.size() + 32;
            dyn_.pop_back();
        }
    }
    std::deque<Header> dyn_;
    size_t dyn_size_ = 0, max_size_ = 4096;
    const size_t settings_max_ = 4096;  // we never advertise another SETTINGS_HEADER_TABLE_SIZE
};

inline void hpack_put_int(std::string& o, uint8_t first, int prefix_bits, uint64_t v) {
    const uint64_t max = (1u << prefix_bits) - 1;
    if (v < max) { o.push_back((char)(first | v)); return; }
    o.push_back((char)(first | max));
    v -= max;
    while (v >= 128) { o.push_back((char)((v & 0x7f) | 0x80)); v >>= 7; }
    o.push_back((char)v);
}
// literal header field without indexing, new name, no Huffman: valid for any decoder state
inline void hpack_encode(std::string& o, const Headers& hs) {
    for (const auto& h : hs) {
        o.push_back(0x00);
        hpack_put_int(o, 0x00, 7, h.name.size());
        o += h.name;
        hpack_put_int(o, 0x00, 7, h.value.size());
        o += h.value;
    }
}

// ---- frames (RFC 9113) --------------------------------------------------------------------------------------------
enum : uint8_t { F_DATA = 0, F_HEADERS = 1, F_PRIORITY = 2, F_RST_STREAM = 3, F_SETTINGS = 4, F_PUSH_PROMISE = 5, F_PING = 6,
                 F_GOAWAY = 7, F_WINDOW_UPDATE = 8, F_CONTINUATION = 9 };
enum : uint8_t { FL_END_STREAM = 0x1, FL_ACK = 0x1, FL_END_HEADERS = 0x4, FL_PADDED = 0x8, FL_PRIORITY = 0x20 };
static const char kPreface[] = "PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n";

struct Frame { uint8_t type = 0, flags = 0; uint32_t stream = 0; std::string payload; };

struct Status { int code = 0; std::string message; };  // grpc-status / grpc-message
enum { GRPC_OK = 0, GRPC_CANCELLED = 1, GRPC_UNKNOWN = 2, GRPC_DEADLINE = 4, GRPC_UNIMPLEMENTED = 12, GRPC_INTERNAL = 13, GRPC_UNAVAILABLE = 14 };

inline std::string percent_encode(const std::string& s) {  // grpc-message
    static const char* hex = "0123456789ABCDEF";
    std::string o;
    for (unsigned char c : s) {
        if (c >= 0x20 && c <= 0x7e && c != '%') o.push_back((char)c);
        else { o.push_back('%'); o.push_back(hex[c >> 4]); o.push_back(hex[c & 15]); }
    }
    return o;
}
inline std::string percent_decode(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == '%' && i + 2 < s.size() + 0 && isxdigit((unsigned char)s[i + 1]) && isxdigit((unsigned char)s[i + 2])) {
            o.push_back((char)strtol(s.substr(i + 1, 2).c_str(), nullptr, 16));
            i += 2;
        } else o.push_back(s[i]);
    }
    return o;
}
inline std::string grpc_frame(const std::string& msg) {  // 1-byte compressed flag + 4-byte big-endian length
    std::string o;
    o.push_back(0);
    const uint32_t n = (uint32_t)msg.size();
    o.push_back((char)(n >> 24)); o.push_back((char)(n >> 16)); o.push_back((char)(n >> 8)); o.push_back((char)n);
    o += msg;
    return o;
}
// Splits a request/response body into messages.  false: truncated or compressed (unsupported).
inline bool grpc_unframe(const std::string& body, std::vector<std::string>& msgs) {
    size_t p = 0;
    while (p < body.size()) {
        if (body.size() - p < 5 || body[p] != 0) return false;
        const uint32_t n = ((uint32_t)(uint8_t)body[p + 1] << 24) | ((uint32_t)(uint8_t)body[p + 2] << 16) |
                           ((uint32_t)(uint8_t)body[p + 3] << 8) | (uint32_t)(uint8_t)body[p + 4];
        if (body.size() - p - 5 < n) return false;
        msgs.emplace_back(body.substr(p + 5, n));
        p += 5 + (size_t)n;
    }
    return true;
}

// One HTTP/2 connection over a connected stream socket.  Reads happen on ONE thread (the owner's loop);
// writes may come from any thread and are serialised here, DATA subject to the peer's flow-control windows.
class Conn {
public:
    explicit Conn(int fd) : fd_(fd) {
        // a peer that stops reading must not wedge a writer (which holds wmu_, and with it the reader's SETTINGS / PING
        // acknowledgements) for ever: a send that makes no progress for 10 s fails and the connection is dropped
        struct timeval tv{10, 0};
        ::setsockopt(fd, SOL_SOCKET, SO_SNDTIMEO, &tv, sizeof tv);
    }
    ~Conn() { close_fd(); }
    Conn(const Conn&) = delete;
    Conn& operator=(const Conn&) = delete;

    void close_fd() {
        int fd = fd_.exchange(-1);
        if (fd >= 0) { ::shutdown(fd, SHUT_RDWR); ::close(fd); }
        { std::lock_guard<std::mutex> l(wmu_); dead_ = true; }
        wcv_.notify_all();
    }
    void shutdown_io() {  // wakes a blocked reader without racing on the descriptor number
        const int fd = fd_.load();
        if (fd >= 0) ::shutdown(fd, SHUT_RDWR);
        { std::lock_guard<std::mutex> l(wmu_); dead_ = true; }
        wcv_.notify_all();
    }
    bool dead() { std::lock_guard<std::mutex> l(wmu_); return dead_; }

    bool read_full(void* buf, size_t n, int timeout_ms = -1) {
        uint8_t* p = (uint8_t*)buf;
        while (n) {
            const int fd = fd_.load();
            if (fd < 0) return false;
            if (timeout_ms >= 0) {
                struct pollfd pf{fd, POLLIN, 0};
                const int pr = ::poll(&pf, 1, timeout_ms);
                if (pr <= 0) return false;
            }
            const ssize_t r = ::recv(fd, p, n, 0);
            if (r == 0) return false;
            if (r < 0) { if (errno == EINTR) continue; return false; }
            p += r; n -= (size_t)r;
        }
        return true;
    }
    bool read_frame(Frame& f, int timeout_ms = -1) {
        uint8_t h[9];
        if (!read_full(h, 9, timeout_ms)) return false;
        const uint32_t len = ((uint32_t)h[0] << 16) | ((uint32_t)h[1] << 8) | h[2];
        if (len > (1u << 20)) return false;
        f.type = h[3]; f.flags = h[4];
        f.stream = (((uint32_t)h[5] << 24) | ((uint32_t)h[6] << 16) | ((uint32_t)h[7] << 8) | h[8]) & 0x7fffffffu;
        f.payload.resize(len);
        return len == 0 || read_full(&f.payload[0], len, timeout_ms);
    }
    bool write_raw(const void* buf, size_t n) {  // caller holds wmu_
        const uint8_t* p = (const uint8_t*)buf;
        while (n) {
            const int fd = fd_.load();
            if (fd < 0) return false;
            const ssize_t r = ::send(fd, p, n, MSG_NOSIGNAL);
            if (r < 0) { if (errno == EINTR) continue; return false; }
            p += r; n -= (size_t)r;
        }
        return true;
    }
    bool write_frame(uint8_t type, uint8_t flags, uint32_t stream, const std::string& payload) {
        std::lock_guard<std::mutex> l(wmu_);
        return write_frame_locked(type, flags, stream, payload.data(), payload.size());
    }
    bool write_preface() { std::lock_guard<std::mutex> l(wmu_); return write_raw(kPreface, 24); }

    // Headers in one HEADERS frame (blocks here are far below the frame size limit).
    bool send_headers(uint32_t stream, const Headers& hs, bool end_stream) {
        std::string block;
        hpack_encode(block, hs);
        return write_frame(F_HEADERS, (uint8_t)(FL_END_HEADERS | (end_stream ? FL_END_STREAM : 0)), stream, block);
    }
    // DATA, split to the peer's frame size, waiting for flow-control credit (connection and stream).
    bool send_data(uint32_t stream, const std::string& data, bool end_stream, int timeout_ms = 10000) {
        size_t off = 0;
        std::unique_lock<std::mutex> l(wmu_);
        do {
            const size_t want = std::min<size_t>(data.size() - off, peer_max_frame_);
            if (want > 0) {
                const auto ok = [&] { return dead_ || (conn_window_ > 0 && stream_window_locked(stream) > 0); };
                if (!wcv_.wait_for(l, std::chrono::milliseconds(timeout_ms), ok) || dead_) return false;
            }
            const int64_t sw = stream_window_locked(stream);  // looked up again: the map may have changed while waiting
            const size_t n = (size_t)std::min<int64_t>((int64_t)want, std::min(conn_window_, std::max<int64_t>(sw, 0)));
            const bool last = off + n == data.size();
            if (!write_frame_locked(F_DATA, (uint8_t)(last && end_stream ? FL_END_STREAM : 0), stream, data.data() + off, n)) return false;
            conn_window_ -= (int64_t)n;
            stream_window_locked(stream) -= (int64_t)n;
            off += n;
        } while (off < data.size());
        return true;
    }
    // A stream this endpoint is serving / has opened: only those have a send window that WINDOW_UPDATE may move.
    void open_stream(uint32_t stream) { std::lock_guard<std::mutex> l(wmu_); stream_window_.emplace(stream, peer_initial_window_); }
    void forget_stream(uint32_t stream) { std::lock_guard<std::mutex> l(wmu_); stream_window_.erase(stream); }
    size_t tracked_streams() { std::lock_guard<std::mutex> l(wmu_); return stream_window_.size(); }

    // Bookkeeping for frames every endpoint must answer.  Returns false on a connection error.
    bool handle_control(const Frame& f) {
        switch (f.type) {
        case F_SETTINGS:
            if (f.flags & FL_ACK) return true;
            if (f.payload.size() % 6) return false;
            {
                std::lock_guard<std::mutex> l(wmu_);
                for (size_t i = 0; i + 6 <= f.payload.size(); i += 6) {
                    const uint16_t id = (uint16_t)(((uint8_t)f.payload[i] << 8) | (uint8_t)f.payload[i + 1]);
                    const uint32_t v = ((uint32_t)(uint8_t)f.payload[i + 2] << 24) | ((uint32_t)(uint8_t)f.payload[i + 3] << 16) |
                                       ((uint32_t)(uint8_t)f.payload[i + 4] << 8) | (uint32_t)(uint8_t)f.payload[i + 5];
                    if (id == 0x4) {  // SETTINGS_INITIAL_WINDOW_SIZE: applies to every stream, as a delta
                        for (auto& kv : stream_window_) kv.second += (int64_t)v - peer_initial_window_;
                        peer_initial_window_ = (int64_t)v;
                    } else if (id == 0x5 && v >= 16384 && v <= 16777215) peer_max_frame_ = std::min<size_t>(v, 1u << 20);
                }
                if (!write_frame_locked(F_SETTINGS, FL_ACK, 0, nullptr, 0)) return false;
            }
            wcv_.notify_all();
            return true;
        case F_PING:
            if (f.flags & FL_ACK) return true;
            if (f.payload.size() != 8) return false;
            return write_frame(F_PING, FL_ACK, 0, f.payload);
        case F_WINDOW_UPDATE: {
            if (f.payload.size() != 4) return false;
            const uint32_t inc = (((uint32_t)(uint8_t)f.payload[0] << 24) | ((uint32_t)(uint8_t)f.payload[1] << 16) |
                                  ((uint32_t)(uint8_t)f.payload[2] << 8) | (uint32_t)(uint8_t)f.payload[3]) & 0x7fffffffu;
            {
                std::lock_guard<std::mutex> l(wmu_);
                if (f.stream == 0) conn_window_ += inc;
                else {
                    // RFC 9113 5.1: WINDOW_UPDATE may arrive for a stream that is already closed (or was never opened by
                    // a confused peer) and is then ignored -- it must not create bookkeeping that nothing ever erases
                    auto it = stream_window_.find(f.stream);
                    if (it != stream_window_.end()) it->second += inc;
                }
            }
            wcv_.notify_all();
            return true;
        }
        default:
            return true;
        }
    }
    // Give the peer its credit back for `n` received DATA bytes.
    bool replenish(uint32_t stream, size_t n, bool stream_open) {
        if (!n) return true;
        std::string inc;
        inc.push_back((char)((n >> 24) & 0x7f)); inc.push_back((char)(n >> 16)); inc.push_back((char)(n >> 8)); inc.push_back((char)n);
        if (!write_frame(F_WINDOW_UPDATE, 0, 0, inc)) return false;
        return !stream_open || write_frame(F_WINDOW_UPDATE, 0, stream, inc);
    }
    // Strips padding / priority from a HEADERS or DATA payload.
    static bool strip(const Frame& f, std::string& body) {
        size_t off = 0, pad = 0;
        if (f.flags & FL_PADDED) { if (f.payload.empty()) return false; pad = (uint8_t)f.payload[0]; off = 1; }
        if (f.type == F_HEADERS && (f.flags & FL_PRIORITY)) off += 5;
        if (off + pad > f.payload.size()) return false;
        body.assign(f.payload, off, f.payload.size() - off - pad);
        return true;
    }
    HpackDecoder hpack;

private:
    bool write_frame_locked(uint8_t type, uint8_t flags, uint32_t stream, const char* p, size_t n) {
        if (dead_) return false;
        uint8_t h[9] = {(uint8_t)(n >> 16), (uint8_t)(n >> 8), (uint8_t)n, type, flags,
                        (uint8_t)(stream >> 24), (uint8_t)(stream >> 16), (uint8_t)(stream >> 8), (uint8_t)stream};
        if (!write_raw(h, 9) || (n && !write_raw(p, n))) { dead_ = true; return false; }
        return true;
    }
    int64_t& stream_window_locked(uint32_t stream) {
        auto it = stream_window_.find(stream);
        if (it == stream_window_.end()) it = stream_window_.emplace(stream, peer_initial_window_).first;
        return it->second;
    }
    std::atomic<int> fd_;
    std::mutex wmu_;
    std::condition_variable wcv_;
    bool dead_ = false;
    int64_t conn_window_ = 65535, peer_initial_window_ = 65535;
    size_t peer_max_frame_ = 16384;
    std::map<uint32_t, int64_t> stream_window_;
};

inline std::string find_header(const Headers& hs, const char* name) {
    for (auto& h : hs) if (h.name == name) return h.value;
    return "";
}

// ---- server -----------------------------------------------------------------------------------------------------------
// A response stream of a server-streaming call; usable from any thread until finish().
class ServerStream {
public:
    ServerStream(std::shared_ptr<Conn> c, uint32_t id) : conn_(std::move(c)), id_(id) {}
    bool send(const std::string& msg) {
        std::lock_guard<std::mutex> l(mu_);
        if (done_ || cancelled_.load()) return false;
        if (!headers_sent_) {
            if (!conn_->send_headers(id_, {{":status", "200"}, {"content-type", "application/grpc"}}, false)) { cancelled_ = true; return false; }
            headers_sent_ = true;
        }
        if (!conn_->send_data(id_, grpc_frame(msg), false)) { cancelled_ = true; return false; }
        return true;
    }
    void finish(const Status& st) {
        std::lock_guard<std::mutex> l(mu_);
        if (done_) return;
        done_ = true;
        if (cancelled_.load()) return;
        Headers t;
        if (!headers_sent_) { t.push_back({":status", "200"}); t.push_back({"content-type", "application/grpc"}); }
        t.push_back({"grpc-status", std::to_string(st.code)});
        if (!st.message.empty()) t.push_back({"grpc-message", percent_encode(st.message)});
        conn_->send_headers(id_, t, true);
        conn_->forget_stream(id_);
    }
    void cancel() { cancelled_ = true; }  // RST_STREAM from the peer, or the connection went away
    bool cancelled() const { return cancelled_.load() || conn_->dead(); }

private:
    std::shared_ptr<Conn> conn_;
    uint32_t id_;
    std::mutex mu_;
    bool headers_sent_ = false, done_ = false;
    std::atomic<bool> cancelled_{false};
};

class GrpcServer {
public:
    using Unary = std::function<Status(const std::string& request, std::string& response)>;
    using Streaming = std::function<void(const std::string& request, std::shared_ptr<ServerStream>)>;  // must not block
    void add_unary(const std::string& path, Unary fn) { unary_[path] = std::move(fn); }
    void add_server_streaming(const std::string& path, Streaming fn) { streaming_[path] = std::move(fn); }

    // dpm/plugin.go:93-123: remove a stale socket, listen, serve in the background.
    bool listen_unix(const std::string& path, std::string& err) {
        struct sockaddr_un a{};
        if (path.size() >= sizeof a.sun_path) { err = "socket path too long"; return false; }
        ::unlink(path.c_str());
        lfd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        if (lfd_ < 0) { err = std::string("socket: ") + strerror(errno); return false; }
        a.sun_family = AF_UNIX;
        memcpy(a.sun_path, path.c_str(), path.size() + 1);
        if (::bind(lfd_, (struct sockaddr*)&a, sizeof a) != 0 || ::listen(lfd_, 16) != 0) {
            err = std::string("bind/listen ") + path + ": " + strerror(errno);
            ::close(lfd_); lfd_ = -1;
            return false;
        }
        path_ = path;
        acceptor_ = std::thread([this] { accept_loop(); });
        return true;
    }
    void stop() {
        if (stopping_.exchange(true)) return;
        if (lfd_ >= 0) ::shutdown(lfd_, SHUT_RDWR);
        if (acceptor_.joinable()) acceptor_.join();
        if (lfd_ >= 0) { ::close(lfd_); lfd_ = -1; }
        std::vector<std::shared_ptr<Peer>> peers;
        { std::lock_guard<std::mutex> l(mu_); peers.swap(peers_); }
        for (auto& p : peers) {
            std::string goaway(8, '\0');  // last stream id 0 is conservative; error NO_ERROR
            p->conn->write_frame(F_GOAWAY, 0, 0, goaway);
            p->conn->shutdown_io();
        }
        for (auto& p : peers) if (p->th.joinable()) p->th.join();
        if (!path_.empty()) ::unlink(path_.c_str());
    }
    ~GrpcServer() { stop(); }
    const std::string& path() const { return path_; }

private:
    struct Req { Headers headers; std::string block, body; bool headers_done = false; std::shared_ptr<ServerStream> stream; };
    static constexpr size_t kMaxConnections = 64;
    static constexpr size_t kMaxOpenRequests = 256, kMaxHeaderBlock = 64 * 1024, kMaxMessageBytes = 4 * 1024 * 1024;
    struct Peer { std::shared_ptr<Conn> conn; std::thread th; std::atomic<bool> done{false}; };

    void accept_loop() {
        while (!stopping_.load()) {
            const int fd = ::accept4(lfd_, nullptr, nullptr, SOCK_CLOEXEC);
            if (fd < 0) { if (errno == EINTR) continue; break; }
            auto p = std::make_shared<Peer>();
            p->conn = std::make_shared<Conn>(fd);
            {
                std::lock_guard<std::mutex> l(mu_);
                if (stopping_.load()) { p->conn->close_fd(); break; }
                for (auto it = peers_.begin(); it != peers_.end();) {  // reap connections that have ended
                    if ((*it)->done.load()) { if ((*it)->th.joinable()) (*it)->th.join(); it = peers_.erase(it); }
                    else ++it;
                }
                if (peers_.size() >= kMaxConnections) { p->conn->close_fd(); continue; }  // the kubelet needs one
                peers_.push_back(p);
                p->th = std::thread([this, p] { serve(p->conn); p->done = true; });
            }
        }
    }

    // One connection: this thread reads frames; requests run on a per-connection worker so that the reader keeps
    // answering PING / WINDOW_UPDATE while a handler waits for flow-control credit.
    void serve(std::shared_ptr<Conn> c) {
        char pre[24];
        if (!c->read_full(pre, 24, 10000) || memcmp(pre, kPreface, 24) != 0) { c->shutdown_io(); return; }
        if (!c->write_frame(F_SETTINGS, 0, 0, "")) return;
        std::mutex qmu;
        std::condition_variable qcv;
        std::deque<std::function<void()>> q;
        bool quit = false;
        std::thread worker([&] {
            for (;;) {
                std::function<void()> fn;
                {
                    std::unique_lock<std::mutex> l(qmu);
                    qcv.wait(l, [&] { return quit || !q.empty(); });
                    if (q.empty()) return;
                    fn = std::move(q.front());
                    q.pop_front();
                }
                fn();
            }
        });
        std::map<uint32_t, Req> reqs;
        std::map<uint32_t, std::shared_ptr<ServerStream>> open_streams;
        uint32_t continuing = 0;
        Frame f;
        while (c->read_frame(f)) {
            if (continuing && (f.type != F_CONTINUATION || f.stream != continuing)) break;  // protocol error
            bool ok = true;
            switch (f.type) {
            case F_HEADERS: case F_CONTINUATION: {
                if (f.stream == 0) { ok = false; break; }
                if (reqs.size() >= kMaxOpenRequests && !reqs.count(f.stream)) { ok = false; break; }  // a peer hoarding streams
                if (!reqs.count(f.stream)) c->open_stream(f.stream);
                Req& r = reqs[f.stream];
                std::string frag;
                if (f.type == F_HEADERS) { if (!Conn::strip(f, frag)) { ok = false; break; } }
                else frag = f.payload;
                r.block += frag;
                if (r.block.size() > kMaxHeaderBlock) { ok = false; break; }
                const bool end_stream = f.type == F_HEADERS && (f.flags & FL_END_STREAM);
                if (f.flags & FL_END_HEADERS) {
                    continuing = 0;
                    if (!c->hpack.decode((const uint8_t*)r.block.data(), r.block.size(), r.headers)) { ok = false; break; }
                    r.block.clear();
                    r.headers_done = true;
                } else continuing = f.stream;
                if (end_stream) { if (!r.headers_done) { ok = false; break; } dispatch(c, f.stream, reqs, open_streams, qmu, qcv, q); }
                break;
            }
            case F_DATA: {
                auto it = reqs.find(f.stream);
                std::string body;
                if (!Conn::strip(f, body)) { ok = false; break; }
                const bool end = f.flags & FL_END_STREAM;
                if (it != reqs.end()) {
                    it->second.body += body;
                    if (it->second.body.size() > kMaxMessageBytes) { ok = false; break; }  // gRPC's default receive limit
                }
                ok = c->replenish(f.stream, f.payload.size(), it != reqs.end() && !end);
                if (ok && end && it != reqs.end()) dispatch(c, f.stream, reqs, open_streams, qmu, qcv, q);
                break;
            }
            case F_RST_STREAM: {
                reqs.erase(f.stream);
                auto it = open_streams.find(f.stream);
                if (it != open_streams.end()) { it->second->cancel(); open_streams.erase(it); }
                c->forget_stream(f.stream);
                break;
            }
            case F_GOAWAY:
                ok = false;
                break;
            case F_PUSH_PROMISE:
                ok = false;
                break;
            default:
                ok = c->handle_control(f);
            }
            if (!ok) break;
        }
        c->shutdown_io();
        for (auto& kv : open_streams) kv.second->cancel();
        { std::lock_guard<std::mutex> l(qmu); quit = true; }
        qcv.notify_all();
        worker.join();
    }

    void dispatch(const std::shared_ptr<Conn>& c, uint32_t sid, std::map<uint32_t, Req>& reqs,
                  std::map<uint32_t, std::shared_ptr<ServerStream>>& open_streams, std::mutex& qmu, std::condition_variable& qcv,
                  std::deque<std::function<void()>>& q) {
        Req r = std::move(reqs[sid]);
        reqs.erase(sid);
        const std::string path = find_header(r.headers, ":path");
        auto stream = std::make_shared<ServerStream>(c, sid);
        std::vector<std::string> msgs;
        const bool framed = grpc_unframe(r.body, msgs) && msgs.size() == 1;
        auto u = unary_.find(path);
        auto s = streaming_.find(path);
        std::function<void()> job;
        if (find_header(r.headers, ":method") != "POST" || (u == unary_.end() && s == streaming_.end()))
            job = [stream, path] { stream->finish({GRPC_UNIMPLEMENTED, "unknown method " + path}); };
        else if (!framed)
            job = [stream] { stream->finish({GRPC_INTERNAL, "malformed or compressed request message"}); };
        else if (u != unary_.end()) {
            const Unary* fn = &u->second;
            job = [stream, fn, req = std::move(msgs[0])] {
                std::string resp;
                const Status st = (*fn)(req, resp);
                if (st.code == GRPC_OK) stream->send(resp);
                stream->finish(st);
            };
        } else {
            open_streams[sid] = stream;
            const Streaming* fn = &s->second;
            job = [stream, fn, req = std::move(msgs[0])] { (*fn)(req, stream); };
        }
        { std::lock_guard<std::mutex> l(qmu); q.push_back(std::move(job)); }
        qcv.notify_one();
    }

    std::map<std::string, Unary> unary_;
    std::map<std::string, Streaming> streaming_;
    int lfd_ = -1;
    std::string path_;
    std::thread acceptor_;
    std::atomic<bool> stopping_{false};
    std::mutex mu_;
    std::vector<std::shared_ptr<Peer>> peers_;
};

// ---- client: one unary call on its own connection (dpm/plugin.go:127-162 dials the kubelet per registration) --------
inline Status unary_call_unix(const std::string& sock_path, const std::string& method_path, const std::string& request,
                              std::string& response, int timeout_ms = 5000) {
    struct sockaddr_un a{};
    if (sock_path.size() >= sizeof a.sun_path) return {GRPC_UNAVAILABLE, "socket path too long"};
    const int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return {GRPC_UNAVAILABLE, std::string("socket: ") + strerror(errno)};
    a.sun_family = AF_UNIX;
    memcpy(a.sun_path, sock_path.c_str(), sock_path.size() + 1);
    if (::connect(fd, (struct sockaddr*)&a, sizeof a) != 0) {
        const std::string e = std::string("connect ") + sock_path + ": " + strerror(errno);
        ::close(fd);
        return {GRPC_UNAVAILABLE, e};
    }
    Conn c(fd);
    if (!c.write_preface() || !c.write_frame(F_SETTINGS, 0, 0, "")) return {GRPC_UNAVAILABLE, "write failed"};
    const uint32_t sid = 1;
    c.open_stream(sid);
    if (!c.send_headers(sid, {{":method", "POST"}, {":scheme", "http"}, {":path", method_path}, {":authority", "localhost"},
                              {"content-type", "application/grpc"}, {"te", "trailers"}, {"user-agent", "b200dp-plugind"}}, false))
        return {GRPC_UNAVAILABLE, "write failed"};
    // the request is far below the initial 65,535-byte windows, so no credit is needed before the first read
    if (!c.send_data(sid, grpc_frame(request), true, timeout_ms)) return {GRPC_UNAVAILABLE, "write failed"};
    Headers hs;
    std::string block, body;
    const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
    Frame f;
    bool ended = false;
    while (!ended) {
        const auto left = std::chrono::duration_cast<std::chrono::milliseconds>(deadline - std::chrono::steady_clock::now()).count();
        if (left <= 0) return {GRPC_DEADLINE, "deadline exceeded"};
        if (!c.read_frame(f, (int)left)) return {GRPC_UNAVAILABLE, "connection closed before the response completed"};
        if ((f.type == F_HEADERS || f.type == F_CONTINUATION) && f.stream == sid) {
            std::string frag;
            if (f.type == F_HEADERS) { if (!Conn::strip(f, frag)) return {GRPC_INTERNAL, "bad HEADERS"}; }
            else frag = f.payload;
            block += frag;
            if (f.flags & FL_END_HEADERS) {
                if (!c.hpack.decode((const uint8_t*)block.data(), block.size(), hs)) return {GRPC_INTERNAL, "HPACK error"};
                block.clear();
            }
            if (f.type == F_HEADERS && (f.flags & FL_END_STREAM)) ended = true;
        } else if (f.type == F_DATA && f.stream == sid) {
            std::string d;
            if (!Conn::strip(f, d)) return {GRPC_INTERNAL, "bad DATA"};
            body += d;
            c.replenish(sid, f.payload.size(), !(f.flags & FL_END_STREAM));
            if (f.flags & FL_END_STREAM) ended = true;
        } else if (f.type == F_RST_STREAM && f.stream == sid) return {GRPC_UNAVAILABLE, "stream reset by the server"};
        else if (f.type == F_GOAWAY) return {GRPC_UNAVAILABLE, "server sent GOAWAY"};
        else if (!c.handle_control(f)) return {GRPC_INTERNAL, "HTTP/2 protocol error"};
    }
    if (find_header(hs, ":status") != "200") return {GRPC_UNKNOWN, "HTTP status " + find_header(hs, ":status")};
    const std::string gs = find_header(hs, "grpc-status");
    if (gs.empty()) return {GRPC_UNKNOWN, "no grpc-status in the response"};
    Status st{atoi(gs.c_str()), percent_decode(find_header(hs, "grpc-message"))};
    if (st.code == GRPC_OK) {
        std::vector<std::string> msgs;
        if (!grpc_unframe(body, msgs) || msgs.size() != 1) return {GRPC_INTERNAL, "malformed response message"};
        response = msgs[0];
    }
    return st;
}

// ---- client: a server-streaming call (what the kubelet does with ListAndWatch); single-threaded reader -----------
class ClientStream {
public:
    ~ClientStream() { close(); }
    Status open(const std::string& sock_path, const std::string& method_path, const std::string& request, int timeout_ms = 5000) {
        struct sockaddr_un a{};
        if (sock_path.size() >= sizeof a.sun_path) return {GRPC_UNAVAILABLE, "socket path too long"};
        const int fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
        if (fd < 0) return {GRPC_UNAVAILABLE, std::string("socket: ") + strerror(errno)};
        a.sun_family = AF_UNIX;
        memcpy(a.sun_path, sock_path.c_str(), sock_path.size() + 1);
        if (::connect(fd, (struct sockaddr*)&a, sizeof a) != 0) {
            const std::string e = std::string("connect ") + sock_path + ": " + strerror(errno);
            ::close(fd);
            return {GRPC_UNAVAILABLE, e};
        }
        conn_ = std::make_unique<Conn>(fd);
        conn_->open_stream(sid_);
        if (!conn_->write_preface() || !conn_->write_frame(F_SETTINGS, 0, 0, "") ||
            !conn_->send_headers(sid_, {{":method", "POST"}, {":scheme", "http"}, {":path", method_path}, {":authority", "localhost"},
                                        {"content-type", "application/grpc"}, {"te", "trailers"}}, false) ||
            !conn_->send_data(sid_, grpc_frame(request), true, timeout_ms))
            return {GRPC_UNAVAILABLE, "write failed"};
        return Status{};
    }
    // Next message of the stream.  false: the stream ended (status() says how) or timed out.
    bool next(std::string& msg, int timeout_ms) {
        const auto deadline = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms);
        for (;;) {
            if (pop(msg)) return true;
            if (ended_ || !conn_) return false;
            const auto left = std::chrono::duration_cast<std::chrono::milliseconds>(deadline - std::chrono::steady_clock::now()).count();
            Frame f;
            if (left <= 0 || !conn_->read_frame(f, (int)left)) { if (left > 0) { ended_ = true; status_ = {GRPC_UNAVAILABLE, "connection closed"}; } return false; }
            if ((f.type == F_HEADERS || f.type == F_CONTINUATION) && f.stream == sid_) {
                std::string frag;
                if (f.type == F_HEADERS) { if (!Conn::strip(f, frag)) return fail("bad HEADERS"); }
                else frag = f.payload;
                block_ += frag;
                if (f.flags & FL_END_HEADERS) {
                    if (!conn_->hpack.decode((const uint8_t*)block_.data(), block_.size(), headers_)) return fail("HPACK error");
                    block_.clear();
                }
                if (f.type == F_HEADERS && (f.flags & FL_END_STREAM)) {
                    ended_ = true;
                    const std::string gs = find_header(headers_, "grpc-status");
                    status_ = {gs.empty() ? GRPC_UNKNOWN : atoi(gs.c_str()), percent_decode(find_header(headers_, "grpc-message"))};
                }
            } else if (f.type == F_DATA && f.stream == sid_) {
                std::string d;
                if (!Conn::strip(f, d)) return fail("bad DATA");
                body_ += d;
                conn_->replenish(sid_, f.payload.size(), !(f.flags & FL_END_STREAM));
                if (f.flags & FL_END_STREAM) { ended_ = true; status_ = {GRPC_UNKNOWN, "stream ended without trailers"}; }
            } else if (f.type == F_RST_STREAM && f.stream == sid_) { ended_ = true; status_ = {GRPC_UNAVAILABLE, "stream reset"}; }
            else if (f.type == F_GOAWAY) { ended_ = true; status_ = {GRPC_UNAVAILABLE, "GOAWAY"}; }
            else if (!conn_->handle_control(f)) return fail("HTTP/2 protocol error");
        }
    }
    void close() {
        if (conn_) {
            if (!ended_) conn_->write_frame(F_RST_STREAM, 0, sid_, std::string("\x00\x00\x00\x08", 4));  // CANCEL
            conn_.reset();
        }
    }
    const Status& status() const { return status_; }

private:
    bool pop(std::string& msg) {
        if (body_.size() < 5 || body_[0] != 0) return false;
        const uint32_t n = ((uint32_t)(uint8_t)body_[1] << 24) | ((uint32_t)(uint8_t)body_[2] << 16) | ((uint32_t)(uint8_t)body_[3] << 8) | (uint32_t)(uint8_t)body_[4];
        if (body_.size() - 5 < n) return false;
        msg.assign(body_, 5, n);
        body_.erase(0, 5 + (size_t)n);
        return true;
    }
    bool fail(const char* why) { ended_ = true; status_ = {GRPC_INTERNAL, why}; return false; }
    std::unique_ptr<Conn> conn_;
    const uint32_t sid_ = 1;
    std::string block_, body_;
    Headers headers_;
    bool ended_ = false;
    Status status_;
};

}  // namespace h2
