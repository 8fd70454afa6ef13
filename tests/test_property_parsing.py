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
