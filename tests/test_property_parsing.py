"""Property-based parity of the two text readers every other result rests on -- the
`<key>\\s(\\d+)` first-match reader (amdgpu.go:442-463) and the three-key last-match link reader
(allocator/device.go:107-133) -- between the C++ implementation behind the C ABI and the oracle's
restatement of the Go semantics (unanchored RE2 match, single \\s, bufio line splitting, ParseInt
base 0 with octal / overflow behaviour).  hypothesis, derandomised; CPU only."""
import os

from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import gosem

KEYS = ["drm_render_minor", "location_id", "domain", "type", "node_from", "node_to", "simd_count"]
FRAG = st.sampled_from(KEYS + ["x", "heap_", "_id", " ", "  ", "\t", "\r", "0", "00", "7", "08", "11", "128", "4294967296",
                              "99999999999999999999", "-", "+", "0x1f", "_", "\x0b", "é".encode().decode("latin1")])
LINE = st.lists(FRAG, min_size=0, max_size=6).map("".join)
FILE = st.lists(LINE, min_size=0, max_size=8).map(lambda ls: "\n".join(ls)) | \
    st.lists(LINE, min_size=1, max_size=5).map(lambda ls: "\r\n".join(ls) + "\n")


@settings(max_examples=400, derandomize=True, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=FILE, key=st.sampled_from(KEYS))
def test_first_match_reader(pkg, tmp_path, text, key):
    p = os.path.join(str(tmp_path), "properties")
    with open(p, "wb") as f:
        f.write(text.encode("latin1"))
    want_v, want_e = oamd.ParseTopologyProperties(p, gosem.compile_re2(key + r"\s(\d+)"))
    got_v, got_e = pkg.amdgpu.ParseTopologyProperties(p, key)
    assert got_v == want_v, (text, key)
    assert (got_e is None) == (want_e is None), (text, key, got_e, want_e)
    if want_e is not None:
        kind = getattr(want_e, "kind", "notfound")
        assert got_e.code == {"syntax": pkg._native.E_SYNTAX, "range": pkg._native.E_RANGE,
                              "notfound": pkg._native.E_NOTFOUND}[kind]


@settings(max_examples=200, derandomize=True, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(text=FILE)
def test_link_reader_through_pair_weights(pkg, tmp_path, text):
    """A single link file with arbitrary content: the pair-weight table the allocator derives from
    it equals the oracle's (last match wins; a bad 32-bit number skips the file)."""
    root = os.path.join(str(tmp_path), "nodes")
    for nid in (2, 3):
        os.makedirs(os.path.join(root, str(nid)), exist_ok=True)
        with open(os.path.join(root, str(nid), "properties"), "w") as f:
            f.write("drm_render_minor %d\n" % (126 + nid))
    os.makedirs(os.path.join(root, "2", "io_links", "0"), exist_ok=True)
    with open(os.path.join(root, "2", "io_links", "0", "properties"), "wb") as f:
        f.write(text.encode("latin1"))
    os.makedirs(os.path.join(root, "3", "io_links", "0"), exist_ok=True)
    with open(os.path.join(root, "3", "io_links", "0", "properties"), "w") as f:
        f.write("type 11\nnode_from 3\nnode_to 2\n")
    odevs = [oalloc.Device(Id="test1", NodeId=2, DevId="0"), oalloc.Device(Id="test2", NodeId=3, DevId="1", NumaNode=1)]
    want = {}
    oalloc.fetchAllPairWeights(odevs, want, root)
    pol = pkg.allocator.NewBestEffortPolicy()
    err = pol.Init([pkg.allocator.Device(Id=d.Id, NodeId=d.NodeId, DevId=d.DevId, NumaNode=d.NumaNode) for d in odevs], root)
    assert err is None and pol.pair_weights() == want, text


def test_read_to_eof_not_to_the_first_short_read(pkg, tmp_path):
    """seq_file-backed files (debugfs amdgpu_firmware_info, /proc) may hand out short reads before EOF; Go's
    os.ReadFile keeps reading until EOF and so must the reader (amdgpu.go:467-490).  A FIFO fed in small delayed
    chunks stands in for such a file."""
    import ctypes as C
    import os
    import threading
    import time
    N = pkg._native
    lines = ["%s feature version: %d, firmware version: 0x%08x" % (name, i, 0x1000 + i)
             for i, name in enumerate(["VCE", "UVD", "MC", "ME", "PFP", "CE", "RLC", "MEC", "SMC", "SDMA0"])]
    fifo = str(tmp_path / "amdgpu_firmware_info")
    os.mkfifo(fifo)

    def feed():
        with open(fifo, "w") as f:
            for ln in lines:
                f.write(ln + "\n")
                f.flush()
                time.sleep(0.01)
    t = threading.Thread(target=feed)
    t.start()
    out = (N.FwEntry * 64)()
    n = C.c_int(0)
    rc = N.lib.b2dp_parse_debugfs_firmware_info(fifo.encode(), out, 64, C.byref(n))
    t.join()
    assert rc == 0 and n.value == len(lines)
    assert sorted(e.name.decode() for e in out[:n.value]) == sorted(ln.split()[0] for ln in lines)
