"""The allocator's exact fast path (subset DP over whole groups, csrc/allocator.cpp
best_candidate_fast) against the reference-shaped enumeration: the Python oracle's FIFO
(oracle/allocator.py, device.go:353-442) for the chosen ids IN ORDER, and the library's own
literal enumeration (b2dp_allocator_candidates) for the winning weight and sequence count.
Random topologies with uneven groups, mixed link types and NUMA layouts => many ties."""
import math
import random

import pytest

from oracle import allocator as oalloc


def make_case(rng):
    n_groups = rng.randint(1, 7)
    devs, node = [], 2
    for g in range(n_groups):
        for p in range(rng.randint(1, 5)):
            devs.append(dict(Id=("test%d" % (g + 1)) if p == 0 else "amdgpu_xcp_%d" % (g * 8 + p), NodeId=node,
                             NumaNode=rng.randint(0, 1), DevId=str(g)))
            node += 1
    links = []
    for a in devs:
        for b in devs:
            if a["NodeId"] < b["NodeId"] and rng.random() < 0.9:
                links.append((a["NodeId"], b["NodeId"], rng.choice([11, 11, 11, 2, 5])))
    return devs, links


@pytest.mark.parametrize("seed", range(250))
def test_fast_path_equals_enumeration(pkg, seed):
    rng = random.Random(7000 + seed)
    devs, links = make_case(rng)
    odevs = [oalloc.Device(**d) for d in devs]
    opol = oalloc.BestEffortPolicy()
    by_node = {d.NodeId: d for d in odevs}
    for a, b, t in links:
        opol.p2pWeights.setdefault(a, {})[b] = oalloc.calculatePairWeight(by_node[a], by_node[b], t)
    opol.devices = odevs
    opol.devicesMap = {d.Id: d for d in odevs}
    opol.devicePartitions = oalloc.groupPartitionsByDevId(odevs)
    pol = pkg.allocator.NewBestEffortPolicy()
    err = pol.InitLinks([pkg.allocator.Device(**d) for d in devs], links)
    if not links:
        assert err is not None
        return
    assert err is None and pol.pair_weights() == opol.p2pWeights
    ids = [d["Id"] for d in devs]
    for _ in range(12):
        avail = rng.sample(ids, rng.randint(1, len(ids)))
        req = rng.sample(avail, rng.randint(0, min(3, len(avail))))
        size = rng.randint(1, len(avail))
        want, werr = opol.Allocate(list(avail), list(req), size)
        got, gerr = pol.Allocate(avail, req, size)
        assert (gerr is None) == (werr is None), (avail, req, size, gerr, werr)
        if werr is not None:
            assert str(gerr) == str(werr)
            continue
        assert got == want, (avail, req, size)
        if len(avail) != size and len(req) != size:          # not one of the two shortcuts
            (ncand, best), cerr = pol.candidates(avail, req, size)       # the literal enumeration
            assert cerr is None and best == opol.last_score and ncand == opol.last_candidates


def test_fast_path_latency(pkg, tmp_path):
    """8 single-partition GPUs, size 7: 40,320 ordered sequences in the reference's enumeration."""
    import time
    root = str(tmp_path / "t")
    ids = pkg.synth.write_b200_tree(root, n_gpus=8)
    with pkg.Context("kfd:" + root) as ctx:
        assert ctx.start() == 0
        dt = 1e9
        for _ in range(3):                                    # best of 3: a descheduled thread must not fail the bound
            t0 = time.perf_counter()
            got = ctx.preferred_allocation(ids, [], 7)
            dt = min(dt, time.perf_counter() - t0)
        assert len(got) == 7 and dt < 0.005, dt
    assert math.perm(8, 7) == 40320
