"""Go standard-library semantics the reference relies on (oracle; test infrastructure only).

Restated from the Go 1.22 stdlib behaviour, as used at the cited reference call sites.
"""
import os
import re

INT64_MAX = (1 << 63) - 1
INT64_MIN = -(1 << 63)


class GoPanic(Exception):
    """The reference would panic here (nil dereference / slice out of range)."""


class GoFatal(Exception):
    """The reference calls glog.Fatalf here (process exit)."""


class ParseError(Exception):
    """strconv.NumError; .value is what Go returns alongside the error."""

    def __init__(self, kind, value=0):
        super().__init__(kind)
        self.kind = kind  # "syntax" | "range"
        self.value = value


# --- regexp (RE2): the reference's patterns use only literals, \s, \d, \w, (), +, [].
# Python's \s also matches \v and (for str) Unicode spaces, so translate to explicit
# ASCII classes and match on bytes.  All reference uses are unanchored FindStringSubmatch.
def compile_re2(pattern: str):
    p = pattern.replace(r"\s", r"[\t\n\f\r ]").replace(r"\d", r"[0-9]").replace(r"\w", r"[0-9A-Za-z_]")
    return re.compile(p.encode("ascii"))


def scanner_lines(data: bytes):
    """bufio.Scanner with ScanLines: split on \\n, drop one trailing \\r, no final empty token.

    A line longer than bufio.MaxScanTokenSize (64 KiB) makes Scan() return false
    (ErrTooLong): iteration simply stops, as the reference never checks scanner.Err()
    on these paths (amdgpu.go:450-459, device.go:116-130)."""
    max_tok = 64 * 1024
    pos = 0
    n = len(data)
    while pos < n:
        nl = data.find(b"\n", pos)
        if nl < 0:
            line = data[pos:]
            pos = n
        else:
            line = data[pos:nl]
            pos = nl + 1
        if len(line) >= max_tok:
            return
        if line.endswith(b"\r"):
            line = line[:-1]
        yield line


def scan(data: bytes):
    """(lines, err): like draining a bufio.Scanner and then asking scanner.Err() != nil."""
    lines = list(scanner_lines(data))
    err = any(len(l) >= 64 * 1024 for l in data.split(b"\n"))
    return lines, err


def _lower(c):
    return c | 0x20 if 0x41 <= c <= 0x5A else c


def parse_uint_digits(s: bytes, base: int, bits: int):
    """strconv.ParseUint core for a byte string without sign. Returns value or raises."""
    if not s:
        raise ParseError("syntax", 0)
    base0 = base == 0
    i = 0
    if base0:
        base = 10
        if s[0:1] == b"0":
            if len(s) >= 3 and _lower(s[1]) == ord("b"):
                base, i = 2, 2
            elif len(s) >= 3 and _lower(s[1]) == ord("o"):
                base, i = 8, 2
            elif len(s) >= 3 and _lower(s[1]) == ord("x"):
                base, i = 16, 2
            else:
                base, i = 8, 1
    maxval = (1 << bits) - 1
    n = 0
    underscores = False
    rng = False
    for c in s[i:]:
        if c == ord("_") and base0:
            underscores = True
            continue
        if ord("0") <= c <= ord("9"):
            d = c - ord("0")
        elif ord("a") <= _lower(c) <= ord("z"):
            d = _lower(c) - ord("a") + 10
        else:
            raise ParseError("syntax", 0)
        if d >= base:
            raise ParseError("syntax", 0)
        n = n * base + d
        if n > maxval:
            rng = True
    if underscores and not _underscore_ok(s):
        raise ParseError("syntax", 0)
    if rng:
        raise ParseError("range", maxval)
    return n


def _underscore_ok(s: bytes) -> bool:
    # strconv.underscoreOK, restated
    i = 0
    saw = ord("^")
    hexp = False
    if len(s) >= 2 and s[0:1] == b"0" and _lower(s[1]) in (ord("b"), ord("o"), ord("x")):
        i = 2
        saw = ord("0")
        hexp = _lower(s[1]) == ord("x")
    while i < len(s):
        c = s[i]
        if ord("0") <= c <= ord("9") or (hexp and ord("a") <= _lower(c) <= ord("f")):
            saw = ord("0")
        elif c == ord("_"):
            if saw != ord("0"):
                return False
            saw = ord("_")
        else:
            if saw == ord("_"):
                return False
            saw = ord("!")
        i += 1
    return saw != ord("_")


def parse_int(s: bytes, base: int, bits: int) -> int:
    """strconv.ParseInt(s, base, bits). Raises ParseError(kind, value_go_returns)."""
    if isinstance(s, str):
        s = s.encode()
    if not s:
        raise ParseError("syntax", 0)
    neg = False
    body = s
    if s[0:1] == b"+":
        body = s[1:]
    elif s[0:1] == b"-":
        neg = True
        body = s[1:]
    try:
        un = parse_uint_digits(body, base, 64 if bits == 0 else bits)
    except ParseError as e:
        if e.kind == "range":
            cutoff = 1 << ((64 if bits == 0 else bits) - 1)
            raise ParseError("range", -cutoff if neg else cutoff - 1)
        raise
    b = 64 if bits == 0 else bits
    cutoff = 1 << (b - 1)
    if not neg and un >= cutoff:
        raise ParseError("range", cutoff - 1)
    if neg and un > cutoff:
        raise ParseError("range", -cutoff)
    return -un if neg else un


def atoi(s) -> int:
    """strconv.Atoi: base 10, int is 64-bit on linux/amd64. Raises ParseError."""
    if isinstance(s, str):
        s = s.encode()
    return parse_int(s, 10, 64)


def atoi_ignore_err(s) -> int:
    """`v, _ := strconv.Atoi(s)`: the value Go assigns when the error is discarded."""
    try:
        return atoi(s)
    except ParseError as e:
        return e.value


def to_uint32(v: int) -> int:
    return v & 0xFFFFFFFF


# --- filepath.Glob -----------------------------------------------------------------
def _match_segment(pat: str, name: str) -> bool:
    """path.Match for one path element; supports *, ?, [..] (ranges, ^ negation), \\ escapes.
    Unlike the shell, `*` matches a leading dot."""
    return _match(pat, 0, name, 0)


def _match(p, pi, s, si):
    while pi < len(p):
        c = p[pi]
        if c == "*":
            # collapse stars
            while pi < len(p) and p[pi] == "*":
                pi += 1
            if pi == len(p):
                return True
            for k in range(si, len(s) + 1):
                if _match(p, pi, s, k):
                    return True
            return False
        if si >= len(s):
            return False
        if c == "?":
            pi += 1
            si += 1
        elif c == "[":
            pi += 1
            neg = False
            if pi < len(p) and p[pi] == "^":
                neg = True
                pi += 1
            ok = False
            first = True
            while pi < len(p) and (p[pi] != "]" or first):
                first = False
                lo = p[pi]
                if lo == "\\":
                    pi += 1
                    lo = p[pi]
                pi += 1
                hi = lo
                if pi + 1 < len(p) and p[pi] == "-" and p[pi + 1] != "]":
                    hi = p[pi + 1]
                    if hi == "\\":
                        hi = p[pi + 2]
                        pi += 1
                    pi += 2
                if lo <= s[si] <= hi:
                    ok = True
            pi += 1  # skip ]
            if ok == neg:
                return False
            si += 1
        else:
            if c == "\\":
                pi += 1
                c = p[pi]
            if s[si] != c:
                return False
            pi += 1
            si += 1
    return si == len(s)


def _has_meta(s: str) -> bool:
    return any(ch in s for ch in "*?[\\")


def glob(pattern: str):
    """filepath.Glob: matches come back in lexical order per directory level
    (Readdirnames + sort.Strings); I/O errors are ignored."""
    if not _has_meta(pattern):
        return [pattern] if os.path.lexists(pattern) else []
    d, f = os.path.split(pattern)
    d = d if d else "."
    if d != "/" and d.endswith("/"):
        d = d.rstrip("/")
    dirs = glob(d) if _has_meta(d) else [d]
    out = []
    for dd in dirs:
        if not os.path.isdir(dd):
            continue
        try:
            names = sorted(os.listdir(dd))
        except OSError:
            continue
        if _has_meta(f):
            for n in names:
                if _match_segment(f, n):
                    out.append(os.path.join(dd, n))
        else:
            if f in names:
                out.append(os.path.join(dd, f))
    return out


def fields(line: bytes):
    """strings.Fields on ASCII data: split around runs of white space."""
    return line.split()
