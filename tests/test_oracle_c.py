"""Cross-check the C restatements (oracle/kfd_walk.c, oracle/probe_oracle.c -- the timed CPU
baseline of bench.py) against the Python oracle on every fixture.  CPU only."""

import numpy as np
import pytest

import fake_sysfs
from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import cbind
from oracle import plugin as oplug
from oracle import probe as oprobe
from test_oracle_golden import topo_dir


@pytest.fixture(scope="module")
def clibs():
    cbind.build()
    return cbind.probe_lib(), cbind.kfd_lib()


def _dump(gpus):
    return "".join("%s %d %d %s [%s] [%s] %d %d\n" % (k, v["card"], v["renderD"], v["devID"], v["computePartitionType"],
                                                     v["memoryPartitionType"], v["numaNode"], v["nodeId"])
                   for k, v in sorted(gpus.items()))


@pytest.mark.parametrize("name,kw", [("mi210", {}), ("mi308", dict(compute="cpx", memory="nps1")),
                                     ("cpx", dict(compute="cpx", memory="nps4")), ("cpx", {})])
def test_kfd_walk_matches_python_oracle(clibs, kfd, tmp_path, name, kw):
    import ctypes as C
    _, k = clibs
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, name), **kw)
    buf = C.create_string_buffer(1 << 16)
    n = k.kfdwalk_enumerate(root.encode(), buf, len(buf))
    gpus = oamd.GetAMDGPUs(root)
    assert n == len(gpus) and buf.value.decode() == _dump(gpus)
    assert k.kfdwalk_health(root.encode()) == int(oplug.simpleHealthCheck(root + "/sys/class/kfd/kfd"))
    out = (C.c_longlong * 3)()
    assert k.kfdwalk_pair_weights(root.encode(), out) == len(gpus)
    w = {}
    oalloc.fetchAllPairWeights(oplug.getDevices(root), w, root + "/sys/class/kfd/kfd/topology/nodes")
    pairs = sum(len(r) for r in w.values())
    assert (out[0], out[1], out[2]) == (pairs, len(w), sum(v for r in w.values() for v in r.values()))
    assert k.kfdwalk_cycle(root.encode(), 1) == 2 * len(gpus) + 1


def test_kfd_walk_no_driver(clibs, tmp_path):
    _, k = clibs
    import ctypes as C
    buf = C.create_string_buffer(64)
    assert k.kfdwalk_enumerate(str(tmp_path).encode(), buf, 64) == -1


@pytest.mark.parametrize("n_words,seed,threads", [(1, 0x5EED0000, 1), (4096, 0x5EED0001, 3), ((1 << 20) + 37, 0xDEADBEEF, 8)])
def test_probe_c_matches_numpy(clibs, n_words, seed, threads):
    import ctypes as C
    p, _ = clibs
    src = np.empty(n_words, dtype=np.uint32)
    p.oracle_fill(src.ctypes.data, n_words, seed, threads)
    assert np.array_equal(src, oprobe.pattern(n_words, seed))
    assert p.oracle_expected_checksum(n_words, seed, threads) == oprobe.expected_checksum(n_words, seed) \
        == oprobe.checksum(src)
    nxt = oprobe.next_seed(seed)
    if n_words > 100:
        src[77] ^= 0x10
        src[n_words - 1] ^= 0x80000000
    dst = np.empty_like(src)
    out = (C.c_uint64 * 3)()
    p.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, seed ^ nxt, threads, out)
    cs, bad, first, odst = oprobe.probe_pass(src, seed, nxt)
    assert (out[0], out[1], out[2]) == (cs, bad, first) and np.array_equal(dst, odst)
    assert bad == (2 if n_words > 100 else 0)


def test_probe_fixed_vectors():
    """Known answers that pin the pattern definition itself (oracle/probe.py is the spec)."""
    assert oprobe.initial_seed(3) == 0x5EED0003
    assert oprobe.next_seed(0x5EED0000) == (0x5EED0000 * 1664525 + 1013904223) & 0xFFFFFFFF
    p = oprobe.pattern(4, 0)
    assert [int(x) for x in p] == [0, 2654435761, (2 * 2654435761) & 0xFFFFFFFF, (3 * 2654435761) & 0xFFFFFFFF]
    assert oprobe.expected_checksum(1 << 10, 0x5EED0000) == oprobe.checksum(oprobe.pattern(1 << 10, 0x5EED0000))
    # bit-count closed form used by the product (csrc/cuda_backend.cu expected_checksum)
    n, seed = (1 << 12) + 5, 0xA5A55A5A
    m = oprobe.pattern(n, 0)
    cs = 0
    for b in range(32):
        c = int(((m >> np.uint32(b)) & np.uint32(1)).sum())
        cs += ((n - c) if (seed >> b) & 1 else c) << b
    assert cs & oprobe.NO_BAD == oprobe.expected_checksum(n, seed)
    assert oprobe.classify_link(True, 700.0) == 11 and oprobe.classify_link(True, 50.0) == 2
    assert oprobe.classify_link(False, 700.0) == 0


def test_probe_golden_vectors():
    """Frozen known answers for the probe specification (tests/golden/probe_vectors.json)."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "probe_vectors.json")))
    for c in g["cases"]:
        assert oprobe.initial_seed(c["device"]) == c["seed"] and oprobe.next_seed(c["seed"]) == c["next_seed"]
        src = oprobe.pattern(c["n_words"], c["seed"])
        cs, bad, first, dst = oprobe.probe_pass(src, c["seed"], c["next_seed"])
        assert (cs, bad, first) == (c["checksum"], 0, oprobe.NO_BAD)
        assert [int(x) for x in src[:4]] == c["head"] and [int(x) for x in src[-2:]] == c["tail"]
        assert [int(x) for x in dst[:4]] == c["dst_head"] and oprobe.checksum(dst) == c["dst_checksum"]
        assert np.array_equal(dst, oprobe.pattern(c["n_words"], c["next_seed"]))
    seed, want = next(iter(g["full_size"]["seeds"].items()))
    assert oprobe.expected_checksum(g["full_size"]["n_words"], int(seed)) == want
