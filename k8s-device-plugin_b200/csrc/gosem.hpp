// gosem.hpp -- the Go standard-library behaviours the reference's results depend on,
// implemented natively (no regex engine, single pass over each file).
//
//   filepath.Glob   : per-directory lexical order (Readdirnames + sort.Strings), `*` matches
//                     dot files, I/O errors ignored
//   bufio.Scanner   : lines split on '\n', one trailing '\r' dropped, a token >= 64 KiB stops
//                     the scan (ErrTooLong)
//   regexp `<key>\s(\d+)` (unanchored FindStringSubmatch): leftmost occurrence of <key>
//                     followed by ONE of [\t\n\f\r ] and >= 1 ASCII digit; capture = maximal
//                     digit run
//   strconv.ParseInt(s, 0, bits) / Atoi
#pragma once
#include <dirent.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>

namespace b2dp {
namespace go {

// ---- files ---------------------------------------------------------------------------
// os.ReadFile: read until read() returns 0.  seq_file-backed files (debugfs amdgpu_firmware_info, /proc), pipes and
// the like may hand out short reads before EOF, so a short read alone never ends the loop.
inline bool read_file(const std::string& path, std::string& out) {
    int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    out.clear();
    char buf[4096];
    for (;;) {
        ssize_t r = ::read(fd, buf, sizeof buf);
        if (r < 0) { if (errno == EINTR) continue; ::close(fd); return false; }
        if (r == 0) break;
        out.append(buf, (size_t)r);
    }
    ::close(fd);
    return true;
}
// The hot parse path only (kfd topology `properties` files: sysfs attributes and their captured copies): an attribute
// is at most one page and sysfs hands it out in a single read, a regular file only returns short at EOF -- so there a
// short read IS the end and the read() that would return 0 is saved (one syscall less per file, ~4000 files per
// pair-weight init on the CPX tree).  Same result as read_file on such files; never used for debugfs, /proc or pipes.
inline bool read_attr(const std::string& path, std::string& out) {
    int fd = ::open(path.c_str(), O_RDONLY | O_CLOEXEC);
    if (fd < 0) return false;
    out.clear();
    char buf[4096];
    for (;;) {
        ssize_t r = ::read(fd, buf, sizeof buf);
        if (r < 0) { if (errno == EINTR) continue; ::close(fd); return false; }
        if (r == 0) break;
        out.append(buf, (size_t)r);
        if ((size_t)r < sizeof buf) break;
    }
    ::close(fd);
    return true;
}
inline bool exists(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool lexists(const std::string& p) { struct stat st; return ::lstat(p.c_str(), &st) == 0; }
inline bool is_dir(const std::string& p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

inline std::string join(const std::string& a, const std::string& b) {
    if (a.empty()) return b;
    if (a.back() == '/') return a + b;
    return a + "/" + b;
}
inline std::string base(const std::string& p) {
    size_t e = p.size();
    while (e > 1 && p[e - 1] == '/') --e;
    size_t s = p.rfind('/', e - 1);
    return s == std::string::npos ? p.substr(0, e) : p.substr(s + 1, e - s - 1);
}
inline std::string dir(const std::string& p) {
    size_t s = p.rfind('/');
    if (s == std::string::npos) return ".";
    if (s == 0) return "/";
    return p.substr(0, s);
}

// Sorted names of a directory (filepath.Glob's Readdirnames + sort.Strings).
inline bool list_dir_sorted(const std::string& d, std::vector<std::string>& names) {
    names.clear();
    DIR* dp = ::opendir(d.c_str());
    if (!dp) return false;
    while (struct dirent* e = ::readdir(dp)) {
        const char* n = e->d_name;
        if (n[0] == '.' && (n[1] == 0 || (n[1] == '.' && n[2] == 0))) continue;
        names.emplace_back(n);
    }
    ::closedir(dp);
    std::sort(names.begin(), names.end());
    return true;
}

// The only glob shapes the reference uses on this path.
inline bool is_hex(char c) { return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'f') || (c >= 'A' && c <= 'F'); }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// `dir/*`            -> every entry
inline std::vector<std::string> glob_all(const std::string& d) {
    std::vector<std::string> names, out;
    if (!is_dir(d) || !list_dir_sorted(d, names)) return out;
    for (auto& n : names) out.push_back(join(d, n));
    return out;
}
// `dir/[0-9]*`       -> entries starting with a digit (device.go:160-161,229)
inline std::vector<std::string> glob_digit_prefixed(const std::string& d) {
    std::vector<std::string> names, out;
    if (!is_dir(d) || !list_dir_sorted(d, names)) return out;
    for (auto& n : names) if (!n.empty() && is_digit(n[0])) out.push_back(join(d, n));
    return out;
}
// `dir/<prefix>*`
inline std::vector<std::string> glob_prefixed(const std::string& d, const std::string& prefix) {
    std::vector<std::string> names, out;
    if (!is_dir(d) || !list_dir_sorted(d, names)) return out;
    for (auto& n : names) if (n.compare(0, prefix.size(), prefix) == 0) out.push_back(join(d, n));
    return out;
}
// `dir/[0-9a-fA-F]{4}:*` (amdgpu.go:155)
inline std::vector<std::string> glob_pci_bdf(const std::string& d) {
    std::vector<std::string> names, out;
    if (!is_dir(d) || !list_dir_sorted(d, names)) return out;
    for (auto& n : names)
        if (n.size() >= 5 && is_hex(n[0]) && is_hex(n[1]) && is_hex(n[2]) && is_hex(n[3]) && n[4] == ':')
            out.push_back(join(d, n));
    return out;
}
// `root/topology/nodes/*/properties` (amdgpu.go:111,506; plugin.go:132,162)
inline std::vector<std::string> glob_node_properties(const std::string& topo_root) {
    std::vector<std::string> out;
    for (auto& nd : glob_all(topo_root + "/topology/nodes")) {
        // Go: glob(dir, "properties") needs dir to be a directory and the entry to exist
        if (!is_dir(nd)) continue;
        std::string p = nd + "/properties";
        if (lexists(p)) out.push_back(std::move(p));
    }
    return out;
}

// Same match set and order, visited lazily: fn(path) returns false to stop early (the health
// check returns at the first GPU node, so only the files actually consumed are stat'ed).
template <class F>
inline void for_each_node_properties(const std::string& topo_root, F&& fn) {
    const std::string nodes = topo_root + "/topology/nodes";
    std::vector<std::string> names;
    if (!is_dir(nodes) || !list_dir_sorted(nodes, names)) return;
    for (const auto& n : names) {
        const std::string nd = nodes + "/" + n;
        if (!is_dir(nd)) continue;
        const std::string p = nd + "/properties";
        if (!lexists(p)) continue;
        if (!fn(p)) return;
    }
}

// ---- bufio.Scanner -------------------------------------------------------------------
constexpr size_t kMaxScanToken = 64 * 1024;

// Calls fn(line) for every token; fn returns false to stop early.  Returns true if the scan
// ended with scanner.Err() != nil (token too long).
template <class F>
inline bool scan_lines(std::string_view data, F&& fn) {
    size_t pos = 0, n = data.size();
    while (pos < n) {
        size_t nl = data.find('\n', pos);
        std::string_view line;
        if (nl == std::string_view::npos) { line = data.substr(pos); pos = n; }
        else { line = data.substr(pos, nl - pos); pos = nl + 1; }
        if (line.size() >= kMaxScanToken) return true;
        if (!line.empty() && line.back() == '\r') line.remove_suffix(1);
        if (!fn(line)) return false;
    }
    return false;
}

// ---- regexp `<key>\s(\d+)` -------------------------------------------------------------
inline bool re_space(char c) { return c == '\t' || c == '\n' || c == '\f' || c == '\r' || c == ' '; }

// Returns true and sets `digits` to the capture if the line matches.
inline bool match_key_digits(std::string_view line, std::string_view key, std::string_view& digits) {
    size_t from = 0;
    while (from + key.size() <= line.size()) {
        size_t p = line.find(key, from);
        if (p == std::string_view::npos) return false;
        size_t q = p + key.size();
        if (q + 1 < line.size() && re_space(line[q]) && is_digit(line[q + 1])) {
            size_t e = q + 1;
            while (e < line.size() && is_digit(line[e])) ++e;
            digits = line.substr(q + 1, e - (q + 1));
            return true;
        }
        from = p + 1;
    }
    return false;
}

// ---- strconv ---------------------------------------------------------------------------
enum class NumErr { none, syntax, range };

inline int lower(int c) { return (c >= 'A' && c <= 'Z') ? (c | 0x20) : c; }

// strconv.underscoreOK
inline bool underscore_ok(std::string_view s) {
    char saw = '^';
    size_t i = 0;
    if (!s.empty() && (s[0] == '-' || s[0] == '+')) s.remove_prefix(1);
    bool hex = false;
    if (s.size() >= 2 && s[0] == '0' && (lower(s[1]) == 'b' || lower(s[1]) == 'o' || lower(s[1]) == 'x')) {
        i = 2; saw = '0'; hex = lower(s[1]) == 'x';
    }
    for (; i < s.size(); ++i) {
        char c = s[i];
        if (is_digit(c) || (hex && lower(c) >= 'a' && lower(c) <= 'f')) { saw = '0'; continue; }
        if (c == '_') { if (saw != '0') return false; saw = '_'; continue; }
        if (saw == '_') return false;
        saw = '!';
    }
    return saw != '_';
}

// strconv.ParseUint(s, base, bits); on range error *out = max.
inline NumErr parse_uint(std::string_view s, int base, int bits, uint64_t* out) {
    *out = 0;
    if (s.empty()) return NumErr::syntax;
    const bool base0 = base == 0;
    std::string_view s0 = s;
    if (base0) {
        base = 10;
        if (s[0] == '0') {
            if (s.size() >= 3 && lower(s[1]) == 'b') { base = 2; s.remove_prefix(2); }
            else if (s.size() >= 3 && lower(s[1]) == 'o') { base = 8; s.remove_prefix(2); }
            else if (s.size() >= 3 && lower(s[1]) == 'x') { base = 16; s.remove_prefix(2); }
            else { base = 8; s.remove_prefix(1); }
        }
    } else if (base < 2 || base > 36) return NumErr::syntax;
    const uint64_t maxval = bits == 64 ? ~0ull : ((1ull << bits) - 1);
    const uint64_t cutoff = ~0ull / (uint64_t)base + 1;
    uint64_t n = 0;
    bool underscores = false, over = false;
    for (char c : s) {
        int d;
        if (c == '_' && base0) { underscores = true; continue; }
        if (is_digit(c)) d = c - '0';
        else if (lower(c) >= 'a' && lower(c) <= 'z') d = lower(c) - 'a' + 10;
        else return NumErr::syntax;
        if (d >= base) return NumErr::syntax;
        if (over) continue;
        if (n >= cutoff) { over = true; continue; }
        n *= (uint64_t)base;
        uint64_t n1 = n + (uint64_t)d;
        if (n1 < n || n1 > maxval) { over = true; continue; }
        n = n1;
    }
    if (underscores && !underscore_ok(s0)) return NumErr::syntax;
    if (over) { *out = maxval; return NumErr::range; }
    *out = n;
    return NumErr::none;
}

// strconv.ParseInt(s, base, bits); *out is the value Go returns next to the error.
inline NumErr parse_int(std::string_view s, int base, int bits, int64_t* out) {
    *out = 0;
    if (s.empty()) return NumErr::syntax;
    if (bits == 0) bits = 64;
    bool neg = false;
    std::string_view body = s;
    if (s[0] == '+') body.remove_prefix(1);
    else if (s[0] == '-') { neg = true; body.remove_prefix(1); }
    uint64_t un;
    NumErr e = parse_uint(body, base, bits, &un);
    const uint64_t cutoff = 1ull << (bits - 1);
    if (e == NumErr::syntax) return e;
    if (e == NumErr::range || (!neg && un >= cutoff) || (neg && un > cutoff)) {
        *out = neg ? -(int64_t)(cutoff - 1) - 1 : (int64_t)(cutoff - 1);
        return NumErr::range;
    }
    *out = neg ? (int64_t)(0 - un) : (int64_t)un;
    return NumErr::none;
}
inline NumErr atoi(std::string_view s, int64_t* out) { return parse_int(s, 10, 64, out); }
// `v, _ := strconv.Atoi(s)`
inline int64_t atoi_ignore_err(std::string_view s) { int64_t v; atoi(s, &v); return v; }

// strings.TrimSpace + strings.ToLower on ASCII data
inline std::string trim_space(std::string_view s) {
    auto sp = [](unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
    while (!s.empty() && sp((unsigned char)s.front())) s.remove_prefix(1);
    while (!s.empty() && sp((unsigned char)s.back())) s.remove_suffix(1);
    return std::string(s);
}
inline std::string to_lower(std::string s) {
    for (auto& c : s) c = (char)lower((unsigned char)c);
    return s;
}
// strings.Fields on ASCII data
inline std::vector<std::string_view> fields(std::string_view s) {
    std::vector<std::string_view> out;
    auto sp = [](unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); };
    size_t i = 0;
    while (i < s.size()) {
        while (i < s.size() && sp((unsigned char)s[i])) ++i;
        size_t b = i;
        while (i < s.size() && !sp((unsigned char)s[i])) ++i;
        if (i > b) out.push_back(s.substr(b, i - b));
    }
    return out;
}

}  // namespace go
}  // namespace b2dp
