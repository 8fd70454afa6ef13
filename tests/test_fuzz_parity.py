"""Randomised parity: the C-ABI `kfd:` backend vs the oracle on generated sysfs trees that stress
the Go semantics of SURVEY Appendix C -- lexical glob order with gaps and >= 10 node ids,
unanchored / single-\\s regex matches with decoy lines, base-0 integers (octal, bad octal,
overflow), CRLF, missing keys, sticky loop variables, numa_node edge values, unknown render
minors, last-match-wins link files, bad link numbers, non-matching directory names, and drm
entries short enough to make the reference panic.  Deterministic seeds; CPU only."""
import os
import random

import pytest

from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import gosem
from oracle import labeller as olab
from oracle import plugin as oplug


def _w(path, data):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "wb") as f:
        f.write(data)


def gen_tree(root, rng):
    nodes = os.path.join(root, "sys/class/kfd/kfd/topology/nodes")
    os.makedirs(nodes)
    os.makedirs(os.path.join(root, "sys/devices/platform"))
    drv = os.path.join(root, "sys/module/amdgpu/drivers/pci:amdgpu")
    os.makedirs(drv)
    n_cpu = rng.choice([0, 1, 2])
    node_ids = list(range(n_cpu))
    gpu_nodes = []                    # (node_id, minor, group)
    next_id = n_cpu + rng.choice([0, 0, 7])     # sometimes start at >= 9 so "10" sorts before "9"
    minor = 128
    n_groups = rng.randint(1, 5)
    for g in range(n_groups):
        for p in range(rng.randint(1, 4)):
            gpu_nodes.append((next_id, minor, g))
            node_ids.append(next_id)
            next_id += rng.choice([1, 1, 1, 2])
            minor += rng.choice([1, 1, 2])
    nl = lambda: rng.choice([b"\n", b"\n", b"\n", b"\r\n"])  # noqa: E731
    for k in range(n_cpu):
        _w(os.path.join(nodes, str(k), "properties"),
           b"cpu_cores_count %d" % rng.choice([0, 16, 64]) + nl() + b"simd_count 0" + nl() + b"drm_render_minor 0" + nl())
    for nid, mn, g in gpu_nodes:
        bus = 0x10 + 0x11 * g
        lines = [b"cpu_cores_count 0", b"simd_count %d" % rng.choice([0, 80, 256, 304]),
                 b"simd_per_cu %d" % rng.choice([0, 4, 4, 4]), b"location_id %d" % ((bus << 8) | (rng.randint(0, 3) << 3)),
                 b"domain %d" % rng.choice([0, 0, 1, 65535]), b"unique_id 14073402507705256556"]
        if rng.random() < 0.85:
            lines.append(b"gfx_target_version %d" % rng.choice([0, 90402, 100000]))
        mv = rng.random()
        if mv < 0.70:
            lines.append(b"drm_render_minor %d" % mn)
        elif mv < 0.78:
            lines.append(b"drm_render_minor 0%o" % mn)              # octal spelling of the same minor
        elif mv < 0.84:
            lines.append(b"drm_render_minor 08")                     # syntax error => node skipped
        elif mv < 0.90:
            lines.append(b"drm_render_minor  %d" % mn)               # two spaces: no match
        elif mv < 0.95:
            lines.append(b"drm_render_minor 99999999999999999999")   # range error
        rng.shuffle(lines)
        if rng.random() < 0.3:
            lines.insert(0, b"xdrm_render_minor %d" % rng.choice([0, 3, mn]))   # decoy matched first (unanchored)
        if rng.random() < 0.2:
            lines = [l for l in lines if not l.startswith(rng.choice([b"location_id", b"domain"]))]
        _w(os.path.join(nodes, str(nid), "properties"), b"".join(l + nl() for l in lines))
        _w(os.path.join(nodes, str(nid), "mem_banks/0/properties"),
           b"heap_type 1\nsize_in_bytes %d\nflags 0\n" % rng.choice([17163091968, 68702699520, 206158430208, 1 << 29]))
        li = 0
        for other, _, og in gpu_nodes:
            if other == nid or rng.random() < 0.15:
                continue
            kind = rng.choice(["io_links", "io_links", "p2p_links"])
            t = rng.choice([11, 11, 2, 5, 1])
            body = [b"type %d" % t, b"version_major 0", b"node_from %d" % nid, b"node_to %d" % other, b"weight 15"]
            if rng.random() < 0.15:
                body.append(b"type %d" % rng.choice([2, 11]))       # later line wins
            if rng.random() < 0.06:
                body.append(b"node_to 09")                           # bad number: file skipped
            if rng.random() < 0.06:
                body.append(b"heap_type 7")                          # contains "type 7": unanchored match
            name = str(li) if rng.random() < 0.93 else "x%d" % li   # non-digit names are not globbed
            _w(os.path.join(nodes, str(nid), kind, name, "properties"), b"".join(l + nl() for l in body))
            li += 1
    if rng.random() < 0.2:
        os.makedirs(os.path.join(nodes, "notanode"))
        _w(os.path.join(nodes, "zz", "properties"), b"drm_render_minor 555\nlocation_id 256\ndomain 0\n")
    # PCI functions + platform partitions
    seen_groups = set()
    xcp = 0
    for nid, mn, g in gpu_nodes:
        first = g not in seen_groups
        seen_groups.add(g)
        if first:
            bdf = "%04x:%02x:00.0" % (0, 0x10 + 0x11 * g)
            base = os.path.join(drv, bdf)
            os.makedirs(base)
            nv = rng.random()
            if nv < 0.75:
                _w(os.path.join(base, "numa_node"), b"%d\n" % rng.choice([0, 1]))
            elif nv < 0.85:
                _w(os.path.join(base, "numa_node"), b"-1\n")
            elif nv < 0.92:
                _w(os.path.join(base, "numa_node"), b"abc\n")
            if rng.random() < 0.6:
                _w(os.path.join(base, "current_compute_partition"), rng.choice([b"CPX\n", b"SPX\n", b" cpx \n"]))
                _w(os.path.join(base, "current_memory_partition"), rng.choice([b"NPS1\n", b"NPS4\n"]))
                if rng.random() < 0.7:
                    _w(os.path.join(base, "available_compute_partition"), b"SPX, CPX\n")
        else:
            xcp += 1
            base = os.path.join(root, "sys/devices/platform", "amdgpu_xcp_%d" % (g * 8 + xcp))
            os.makedirs(base)
        os.makedirs(os.path.join(base, "drm", "card%d" % (mn - 127)))
        os.makedirs(os.path.join(base, "drm", "renderD%d" % (mn if rng.random() < 0.93 else 999)))
        if rng.random() < 0.1:
            os.makedirs(os.path.join(base, "drm", "controlD64"))
        if rng.random() < 0.03:
            os.makedirs(os.path.join(base, "drm", rng.choice(["x", "abcde"])))     # reference panics
        cls = os.path.join(root, "sys/class/drm/card%d/device" % (mn - 127))
        _w(os.path.join(cls, "device"), rng.choice([b"0x74a1\n", b"0x2901\n", b"74b5\n"]))
        _w(os.path.join(cls, "product_name"), rng.choice([b"AMD Instinct MI300X (OAM)\n", b"NVIDIA B200\n", b"\n"]))
        _w(os.path.join(cls, "driver/module/version"), b"6.8.5\n")
    if rng.random() < 0.15:
        os.makedirs(os.path.join(drv, "zzzz:00:00.0"))              # not hex: not globbed
        os.makedirs(os.path.join(drv, "module"))


GENS = ["driver-version", "driver-src-version", "device-id", "product-name", "vram", "simd-count", "cu-count",
        "compute-memory-partition", "compute-partitioning-supported", "memory-partitioning-supported", "family", "firmware"]


@pytest.mark.parametrize("seed", range(400))
def test_random_tree_parity(pkg, tmp_path, seed):
    rng = random.Random(0xB200 + seed)
    root = str(tmp_path / "t")
    gen_tree(root, rng)
    kfd = root + "/sys/class/kfd/kfd"
    # stateless readers
    assert pkg.amdgpu.GetDevIdsFromTopology(kfd) == oamd.GetDevIdsFromTopology(kfd)
    assert pkg.amdgpu.GetNodeIdsFromTopology(kfd) == oamd.GetNodeIdsFromTopology(kfd)
    assert pkg.plugin.countGPUDevFromTopology(kfd) == oplug.countGPUDevFromTopology(kfd)
    assert pkg.plugin.simpleHealthCheck(kfd) == oplug.simpleHealthCheck(kfd)
    with pkg.Context("kfd:" + root) as ctx:
        try:
            want = oamd.GetAMDGPUs(root)
        except gosem.GoPanic:
            with pytest.raises(pkg._native.B2dpError) as ei:
                ctx.enumerate()
            assert ei.value.code == pkg._native.E_PANIC
            return
        got = ctx.enumerate()
        assert got == want
        assert ctx.partition_histogram() == oamd.UniquePartitionConfigCount(want)
        for strategy in ("single", "mixed"):
            ores, oerr = oplug.getResourceList(strategy, root)
            if oerr is not None:
                with pytest.raises(pkg._native.B2dpError) as ei:
                    ctx.resource_list(strategy)
                assert ei.value.message == str(oerr)
            else:
                assert ctx.resource_list(strategy) == ores
        # ListAndWatch list per resource name
        for res in set(["gpu"] + list(oamd.UniquePartitionConfigCount(want))):
            wire, st = ctx.list_and_watch(res, pkg._native.LW_INITIAL)
            homog, lw = oplug.list_and_watch_devices(want, res)
            msg = pkg.v1beta1.ListAndWatchResponse.FromString(wire)
            assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in msg.devices] == (lw or [])
            assert msg.SerializeToString() == wire
        # labels
        assert ctx.generate_labels(GENS) == olab.generateLabels({g: True for g in GENS}, root)
        # allocator: pair weights, then random requests through the context's own policy
        odevs = oplug.getDevices(root)
        opol = oalloc.BestEffortPolicy()
        oerr = opol.Init(odevs, kfd + "/topology/nodes")
        rc = ctx.start()
        assert (rc == 0) == (oerr is None)
        if oerr is not None:
            assert pkg._native.lib.b2dp_strerror(rc).decode() == (
                "Besteffort Policy init failed to initialize p2pWeights" if odevs else str(oerr)) or not odevs
            return
        pol = pkg.allocator.NewBestEffortPolicy()
        pdevs = [pkg.allocator.Device(Id=d.Id, NodeId=d.NodeId, NumaNode=d.NumaNode, DevId=d.DevId) for d in odevs]
        assert pol.Init(pdevs, kfd + "/topology/nodes") is None
        assert pol.pair_weights() == opol.p2pWeights
        ids = [d.Id for d in odevs]
        for _ in range(6):
            avail = rng.sample(ids, rng.randint(1, len(ids)))
            req = rng.sample(avail, rng.randint(0, min(2, len(avail))))
            size = rng.randint(0, len(avail) + 1)
            try:
                want_ids, werr = opol.Allocate(list(avail), list(req), size)
            except gosem.GoPanic:
                # e.g. sticky loop variables gave two devices the same NodeId: no candidate reaches
                # the requested size and the reference dereferences a nil DeviceSet
                with pytest.raises(pkg._native.B2dpError) as ei:
                    ctx.preferred_allocation(avail, req, size)
                assert ei.value.code == pkg._native.E_PANIC
                continue
            try:
                got_ids = ctx.preferred_allocation(avail, req, size)
                assert werr is None and got_ids == want_ids, (avail, req, size)
            except pkg._native.B2dpError as e:
                assert werr is not None and e.message == str(werr), (avail, req, size, e, werr)
