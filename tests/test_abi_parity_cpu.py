"""CPU parity: the C-ABI library (kfd: backend + stateless readers + allocator + labels) against
the oracle on the reference's own fixtures, plus the reference's known answers called through
the ABI (so these read like internal/pkg/*/_test.go).  No GPU needed."""
import collections
import ctypes
import os
import re

import pytest

import fake_sysfs
from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import gosem
from oracle import labeller as olab
from oracle import plugin as oplug
from test_oracle_golden import (BEST_POLICY_CASES, CPX, EXP_FEAT, EXP_FW, MI210, MI308, MI308_DEVIDS, MI308_MINORS,
                                TOPO_CFG, topo_dir)

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def P(pkg):
    return pkg


def test_library_exports_every_declared_symbol(P):
    """include/b200dp.h <-> libb200dp.so <-> the ctypes table agree, symbol for symbol."""
    hdr = open(os.path.join(REPO, "include", "b200dp.h")).read()
    declared = set(re.findall(r"B2DP_API\s+[\w\s\*]+?\b(b2dp_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = ctypes.CDLL(P._native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(P._native.SIGNATURES)
    assert lib.b2dp_abi_version() == int(re.search(r"#define B2DP_ABI_VERSION (\d+)", hdr).group(1)) == P._native.ABI_VERSION


def test_no_cpu_fallback_for_the_probe(P, kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi210"))
    with P.Context("kfd:" + root) as ctx:
        with pytest.raises(P._native.B2dpError) as ei:
            ctx.probe_health()
        assert ei.value.code == P._native.E_UNSUPPORTED
        with pytest.raises(P._native.B2dpError):
            ctx.p2p_matrix()


def test_open_errors(P, tmp_path):
    with pytest.raises(P._native.B2dpError) as ei:
        P.Context("kfd:" + str(tmp_path))               # no sys/module/amdgpu/drivers => Fatalf analogue
    assert ei.value.code == P._native.E_NODRIVER
    with pytest.raises(P._native.B2dpError) as ei:
        P.Context("bogus:")
    assert ei.value.code == P._native.E_INVAL
    with pytest.raises(P._native.B2dpError) as ei:
        P.Context("cuda:busy=sometimes")
    assert ei.value.code == P._native.E_INVAL
    import torch
    if not torch.cuda.is_available():            # no GPU: the cuda: backend fails loudly, no fallback
        with pytest.raises(P._native.B2dpError) as ei:
            P.Context("cuda:")
        assert ei.value.code == P._native.E_NOGPU


# ---- amdgpu_test.go through the ABI -----------------------------------------------------------
def test_parse_topology_properties(P, kfd):
    tp = kfd.root("topology-parsing")
    f = P.amdgpu.ParseTopologyProperties
    assert f(tp + "/topology/nodes/1/mem_banks/0/properties", "size_in_bytes")[0] == 17163091968
    assert f(tp + "/topology/nodes/1/mem_banks/0/properties", "flags") == (0, None)
    assert f(tp + "/topology/nodes/2/properties", "simd_count")[0] == 256
    assert f(tp + "/topology/nodes/2/properties", "simd_id_base")[0] == 2147487744
    v, e = f(tp + "/topology/nodes/2/properties", "asdf")
    assert e is not None and e.code == P._native.E_NOTFOUND
    v, e = f(tp + "/topology/nodes/1/properties", "unique_id")       # > int64: clamped + range error
    assert v == (1 << 63) - 1 and e.code == P._native.E_RANGE
    v, e = f(tp + "/nope", "x")
    assert e.code == P._native.E_IO


def test_parse_property_go_semantics(P, tmp_path):
    """Unanchored match, single \\s, first match wins, base-0 integers."""
    p = tmp_path / "properties"
    p.write_bytes(b"xsimd_count 7\nsimd_count  9\nsimd_count 017\nsimd_count 5\nbad 08\nhex 0x10\ncr_key 12\r\n"
                  b"tab_key\t44\nneg -3\nbig 99999999999999999999\n")
    f = P.amdgpu.ParseTopologyProperties
    assert f(str(p), "simd_count") == (7, None)          # matches inside "xsimd_count 7"
    assert f(str(p), "bad")[1].code == P._native.E_SYNTAX   # 08 is bad octal
    assert f(str(p), "hex") == (0, None)                 # \d+ stops at 'x' => "0"
    assert f(str(p), "cr_key") == (12, None)
    assert f(str(p), "tab_key") == (44, None)
    assert f(str(p), "neg")[1].code == P._native.E_NOTFOUND
    assert f(str(p), "big")[1].code == P._native.E_RANGE
    for key in ["simd_count", "bad", "hex", "cr_key", "tab_key", "neg", "big"]:
        ov, oe = oamd.ParseTopologyProperties(str(p), gosem.compile_re2(key + r"\s(\d+)"))
        v, e = f(str(p), key)
        assert v == ov and (e is None) == (oe is None), key


def test_parse_debugfs_firmware_info(P, kfd):
    feat, fw = P.amdgpu.parseDebugFSFirmwareInfo(kfd.root("debugfs-parsing") + "/amdgpu_firmware_info")
    assert feat == EXP_FEAT and fw == EXP_FW
    assert P.amdgpu.parseDebugFSFirmwareInfo("/nonexistent") == ({}, {})


@pytest.mark.parametrize("fixture", ["topology-parsing", "topology-parsing-mi308", "topo-mi300-cpx"])
def test_topology_maps_match_oracle(P, kfd, fixture):
    root = kfd.root(fixture)
    assert P.amdgpu.GetDevIdsFromTopology(root) == oamd.GetDevIdsFromTopology(root)
    assert P.amdgpu.GetNodeIdsFromTopology(root) == oamd.GetNodeIdsFromTopology(root)
    assert P.plugin.countGPUDevFromTopology(root) == oplug.countGPUDevFromTopology(root)
    assert P.plugin.simpleHealthCheck(root) == oplug.simpleHealthCheck(root)


def test_render_dev_ids_mi308(P, kfd):
    got = P.amdgpu.GetDevIdsFromTopology(kfd.root("topology-parsing-mi308"))
    assert sorted(got) == MI308_MINORS
    assert [got[b] for b in range(128, 192, 8)] == MI308_DEVIDS
    assert P.plugin.countGPUDevFromTopology(kfd.root("topology-parsing")) == 2      # plugin_test.go:23-30


# ---- GetAMDGPUs on fake sysfs trees -------------------------------------------------------------
def _enum_both(P, root):
    with P.Context("kfd:" + root) as ctx:
        got = ctx.enumerate()
    return got, oamd.GetAMDGPUs(root)


@pytest.mark.parametrize("name,kw,count", [
    ("mi210", {}, 8),
    ("mi308", dict(compute="cpx", memory="nps1"), 32),
    ("cpx", dict(compute="cpx", memory="nps4"), 63),
    ("cpx", {}, 8),                       # no partition files: xcp devices never inherit a numa => dropped
])
def test_get_amdgpus_matches_oracle(P, kfd, tmp_path, name, kw, count):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, name), **kw)
    got, want = _enum_both(P, root)
    assert got == want and len(got) == count
    assert list(got) == sorted(got)


def test_get_amdgpus_quirks(P, kfd, tmp_path):
    """Sticky loop variables (amdgpu.go:157-159), missing numa_node (skip), unknown render minor."""
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi308"), compute="cpx", memory="nps1",
                            drop_numa_for=("0000:80:00.0",), odd_drm=("amdgpu_xcp_99", 300, 999))
    got, want = _enum_both(P, root)
    assert got == want
    assert "0000:80:00.0" not in got and "amdgpu_xcp_99" not in got
    # the partitions of the dropped GPU found no parent with partition types => dropped too
    assert len(got) == 32 - 4
    # a drm/ entry shorter than 4 chars makes the reference panic (amdgpu.go:202)
    os.makedirs(os.path.join(root, "sys/module/amdgpu/drivers/pci:amdgpu/0000:0a:00.0/drm/x"))
    with pytest.raises(gosem.GoPanic):
        oamd.GetAMDGPUs(root)
    with pytest.raises(P._native.B2dpError) as ei:
        P.Context("kfd:" + root).enumerate()
    assert ei.value.code == P._native.E_PANIC


def test_homogeneity_and_resource_list(P, kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "a"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    with P.Context("kfd:" + root) as ctx:
        assert ctx.is_homogeneous() and ctx.partition_histogram() == {"cpx_nps4": 63}
        assert ctx.resource_list("single") == ["gpu"] == oplug.getResourceList("single", root)[0]
        assert ctx.resource_list("mixed") == ["cpx_nps4"] == oplug.getResourceList("mixed", root)[0]
        assert ctx.partition_supported(0) == oamd.IsComputePartitionSupported(root) is True
        assert ctx.partition_supported(1) == oamd.IsMemoryPartitionSupported(root) is False
        with pytest.raises(P._native.B2dpError):
            ctx.resource_list("both")
    root = fake_sysfs.build(str(tmp_path / "b"), topo_dir(kfd, "mi308"), compute="cpx", memory="nps1",
                            hetero_second=("spx", "nps1"))
    with P.Context("kfd:" + root) as ctx:
        assert not ctx.is_homogeneous()
        assert ctx.partition_histogram() == oamd.UniquePartitionConfigCount(oamd.GetAMDGPUs(root))
        with pytest.raises(P._native.B2dpError) as ei:
            ctx.resource_list("single")
        assert ei.value.code == P._native.E_HETEROGENEOUS
        assert str(oplug.getResourceList("single", root)[1]) == ei.value.message
        assert ctx.resource_list("mixed") == oplug.getResourceList("mixed", root)[0] == ["cpx_nps1", "spx_nps1"]
    root = fake_sysfs.build(str(tmp_path / "c"), topo_dir(kfd, "mi210"))
    with P.Context("kfd:" + root) as ctx:
        assert ctx.resource_list("mixed") == ["gpu"] == oplug.getResourceList("mixed", root)[0]


# ---- allocator: device_test.go / besteffort_policy_test.go through the ABI -----------------------
def _devs(P, cfg):
    return [P.allocator.Device(Id=d.Id, NodeId=d.NodeId, NumaNode=d.NumaNode, DevId=d.DevId)
            for d in oalloc.getTestDevices(**cfg)]


def test_pair_weights_empty_devices(P, kfd):
    pol = P.allocator.NewBestEffortPolicy()
    assert pol.Init([], topo_dir(kfd, "mi308")) is not None                      # device_test.go:80-88


@pytest.mark.parametrize("name,cfg,rows", [("mi308", MI308, 31), ("mi210", MI210, 7), ("cpx", CPX, 62)])
def test_pair_weights_match_oracle(P, kfd, name, cfg, rows):
    pol = P.allocator.NewBestEffortPolicy()
    assert pol.Init(_devs(P, cfg), topo_dir(kfd, name)) is None
    want = {}
    oalloc.fetchAllPairWeights(oalloc.getTestDevices(**cfg), want, topo_dir(kfd, name))
    got = pol.pair_weights()
    assert got == want and len(got) == rows                                       # device_test.go:105
    assert pol.group_count() == len(oalloc.groupPartitionsByDevId(oalloc.getTestDevices(**cfg)))


@pytest.mark.parametrize("size,expected", [(3, 4), (12, 12)])
def test_get_subsets_method(P, kfd, size, expected):                              # device_test.go:125-169
    pol = P.allocator.NewBestEffortPolicy()
    devs = _devs(P, MI308)
    pol.Init(devs, topo_dir(kfd, "mi308"))
    (n, _), err = pol.candidates([d.Id for d in devs], None, size)
    assert err is None and n == expected


@pytest.mark.parametrize("topo,size,available,filtered,required,expected,score", BEST_POLICY_CASES)
def test_best_policy_allocator(P, kfd, topo, size, available, filtered, required, expected, score):
    devs = _devs(P, TOPO_CFG[topo])
    pol = P.allocator.NewBestEffortPolicy()
    assert pol.Init(devs, topo_dir(kfd, topo)) is None
    av = list(available) if available else [d.Id for d in devs]
    av = [a for a in av if a not in filtered]
    result, err = pol.Allocate(av, list(required), size)
    assert err is None and len(result) == size
    if expected is not None:
        assert sorted(result) == sorted(expected)
    # identical to the oracle including ORDER (subset insertion order), score and candidate count
    opol = oalloc.BestEffortPolicy()
    opol.Init(oalloc.getTestDevices(**TOPO_CFG[topo]), topo_dir(kfd, topo))
    oresult, oerr = opol.Allocate(list(av), list(required), size)
    assert oerr is None and result == oresult
    (ncand, best), cerr = pol.candidates(av, list(required), size)
    assert cerr is None and (best, ncand) == score == (opol.last_score, opol.last_candidates)


def test_allocate_validation_errors(P, kfd):
    devs = _devs(P, MI210)
    ids = [d.Id for d in devs]
    pol = P.allocator.NewBestEffortPolicy()
    assert str(pol.Allocate(ids, [], 2)[1]) == oalloc.invalidInit
    pol.Init(devs, topo_dir(kfd, "mi210"))
    assert str(pol.Allocate(ids, [], 0)[1]) == oalloc.invalidSize
    assert str(pol.Allocate(ids[:2], [], 3)[1]) == oalloc.invalidAvailable
    assert str(pol.Allocate(ids, ids[:3], 2)[1]) == oalloc.invalidRequired
    assert str(pol.Allocate(ids[:4], ["nope"], 2)[1]) == oalloc.noCandidateFound
    assert pol.Allocate(["x", "y"], [], 2) == (["x", "y"], None)                  # shortcut, unvalidated
    assert pol.Allocate(ids, ["q", "r"], 2) == (["q", "r"], None)
    _, err = pol.Allocate(ids + ["ghost"], [], 2)                                 # nil *Device => panic in Go
    assert err.code == P._native.E_PANIC
    bad = P.allocator.NewBestEffortPolicy()
    assert str(bad.Init(devs, "/nonexistent")) == "Besteffort Policy init failed to initialize p2pWeights"


def test_allocator_from_links_equals_from_files(P, kfd):
    """Init from an explicit link list (what the P2P matrix produces) == Init from sysfs files."""
    devs = _devs(P, MI210)
    links = []
    nodes = topo_dir(kfd, "mi210")
    res = [gosem.compile_re2(k + r"\s(\d+)") for k in ("node_from", "node_to", "type")]
    for nd in gosem.glob(nodes + "/[0-9]*"):
        for lp in gosem.glob(nd + "/io_links/[0-9]*") + gosem.glob(nd + "/p2p_links/[0-9]*"):
            vals, err = oalloc.fetchTopoProperties(lp + "/properties", res)
            links.append(tuple(vals))
    a, b = P.allocator.NewBestEffortPolicy(), P.allocator.NewBestEffortPolicy()
    assert a.Init(devs, nodes) is None and b.InitLinks(devs, links) is None
    assert a.pair_weights() == b.pair_weights()
    ids = [d.Id for d in devs]
    for size in (1, 2, 3, 5):
        assert a.Allocate(ids, [], size) == b.Allocate(ids, [], size)


# ---- ctx-level Start / GetPreferredAllocation / ListAndWatch / Allocate --------------------------
def test_plugin_on_kfd_backend(P, kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    V = P.v1beta1
    gpus = oamd.GetAMDGPUs(root)
    with P.Context("kfd:" + root) as ctx:
        plug = P.plugin.AMDGPUPlugin(ctx, "gpu")
        plug.Start()
        opts = V.DevicePluginOptions.FromString(plug.GetDevicePluginOptions())
        assert opts.get_preferred_allocation_available and not opts.pre_start_required
        # initial ListAndWatch send == oracle list, byte-identical to protobuf's own encoding
        wire, st = ctx.list_and_watch("gpu", P._native.LW_INITIAL)
        homog, want = oplug.list_and_watch_devices(gpus, "gpu")
        msg = V.ListAndWatchResponse.FromString(wire)
        assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in msg.devices] == want and homog == st.homogeneous
        assert msg.SerializeToString() == wire and st.n_devices == 63
        # heartbeat, exporter absent: node-level health for everyone (health.go:93-95)
        wire, st = ctx.list_and_watch("gpu", P._native.LW_HEARTBEAT | P._native.LW_NO_PROBE)
        assert st.node_healthy == oplug.simpleHealthCheck(root + "/sys/class/kfd/kfd") is True
        assert all(d.health == "Healthy" for d in V.ListAndWatchResponse.FromString(wire).devices)
        # heartbeat with an external per-device source (the exporter merge rule)
        ids = sorted(gpus)
        ext = {ids[0]: False, ids[5]: True, "unknown-device": False}
        wire, st = plug.heartbeat_cycle(external=ext)
        got = [d.health for d in V.ListAndWatchResponse.FromString(wire).devices]
        want = oplug.merge_health(ids, "Healthy", {k: ("Healthy" if v else "Unhealthy") for k, v in ext.items()})
        assert got == want and st.n_unhealthy == 1
        # GetPreferredAllocation through the context's own allocator == oracle on the same devices
        opol = oalloc.BestEffortPolicy()
        assert opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes") is None
        for size in (1, 3, 8, 9):
            want_ids, err = opol.Allocate(list(ids), [], size)
            resp = V.PreferredAllocationResponse.FromString(plug.GetPreferredAllocation([(ids, [], size)]))
            assert err is None and list(resp.container_responses[0].deviceIDs) == want_ids
        with pytest.raises(P.plugin.PluginError) as ei:
            plug.GetPreferredAllocation([(ids[:2], [], 3)])
        assert str(ei.value) == "unable to get preferred allocation list. Error:" + oalloc.invalidAvailable
        # Allocate
        resp = V.AllocateResponse.FromString(plug.Allocate([[ids[0], "bogus", ids[9]], []]))
        got = [[(d.host_path, d.container_path, d.permissions) for d in c.devices] for c in resp.container_responses]
        assert got == [oplug.allocate_device_specs(gpus, [ids[0], "bogus", ids[9]]),
                       oplug.allocate_device_specs(gpus, [])]
        assert ctx.device_specs([ids[0]]) == oplug.allocate_device_specs(gpus, [ids[0]])


def test_allocate_reads_the_stream_snapshot(P, kfd, tmp_path):
    """plugin.go:375 indexes p.AMDGPUs -- the table built when the ListAndWatch stream started
    (plugin.go:231) -- not the live sysfs: a device that disappears afterwards is still allocatable
    with its old paths until the stream restarts, and heartbeats keep sending the old list."""
    import shutil
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi210"))
    gpus = oamd.GetAMDGPUs(root)
    ids = sorted(gpus)
    gone = ids[1]
    V = P.v1beta1
    with P.Context("kfd:" + root) as ctx:
        ctx.list_and_watch("gpu", P._native.LW_INITIAL)
        shutil.rmtree(os.path.join(root, "sys/module/amdgpu/drivers/pci:amdgpu", gone))
        assert gone not in ctx.enumerate() and gone not in oamd.GetAMDGPUs(root)
        assert ctx.device_specs([gone]) == oplug.allocate_device_specs(gpus, [gone])          # snapshot, like p.AMDGPUs
        wire, st = ctx.list_and_watch("gpu", P._native.LW_HEARTBEAT | P._native.LW_NO_PROBE)
        assert [d.ID for d in V.ListAndWatchResponse.FromString(wire).devices] == ids
        ctx.list_and_watch("gpu", P._native.LW_INITIAL)                                       # stream restart
        now = oamd.GetAMDGPUs(root)
        assert ctx.device_specs([gone]) == oplug.allocate_device_specs(now, [gone]) == [("/dev/kfd", "/dev/kfd", "rw")]


def test_list_and_watch_heterogeneous(P, kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi308"), compute="cpx", memory="nps1",
                            hetero_second=("spx", "nps1"))
    gpus = oamd.GetAMDGPUs(root)
    V = P.v1beta1
    with P.Context("kfd:" + root) as ctx:
        for res in ("cpx_nps1", "spx_nps1", "gpu"):
            wire, st = ctx.list_and_watch(res, P._native.LW_INITIAL)
            homog, want = oplug.list_and_watch_devices(gpus, res)
            assert not homog and not st.homogeneous
            if want is None:
                assert wire == b"" and st.n_devices == 0
            else:
                msg = V.ListAndWatchResponse.FromString(wire)
                assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in msg.devices] == want


def test_start_failure_degrades(P, kfd, tmp_path):
    """Allocator init failure => GetPreferredAllocationAvailable false (plugin.go:86-90,211-213)."""
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi210"))
    os.unlink(os.path.join(root, "sys/class/kfd/kfd/topology/nodes"))
    os.makedirs(os.path.join(root, "sys/class/kfd/kfd/topology/nodes"))
    with P.Context("kfd:" + root) as ctx:
        plug = P.plugin.AMDGPUPlugin(ctx)
        plug.Start()
        assert plug.allocatorInitError
        assert not P.v1beta1.DevicePluginOptions.FromString(plug.GetDevicePluginOptions()).get_preferred_allocation_available


def test_negative_numa_encoding(P, kfd, tmp_path):
    """numa_node -1 is kept by the first loop (amdgpu.go:185-195) and encodes as a 10-byte varint."""
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi210"), numa_of_group=lambda g: -1 if g == 0 else 0)
    with P.Context("kfd:" + root) as ctx:
        wire, _ = ctx.list_and_watch("gpu")
        msg = P.v1beta1.ListAndWatchResponse.FromString(wire)
        assert msg.devices[0].topology.nodes[0].ID == -1 and msg.SerializeToString() == wire


# ---- exporter merge rule ---------------------------------------------------------------------------
def test_populate_per_gpu_health(P):
    ids = ["a", "b", "c"]
    f = P.exporter.PopulatePerGPUDHealth
    assert f(ids, "Unhealthy", None) == oplug.merge_health(ids, "Unhealthy", None) == ["Unhealthy"] * 3
    states = [("a", "healthy"), ("b", "HEALTHY"), ("zzz", "healthy")]
    hmap = P.exporter.gpu_states_to_map(states)
    assert hmap == oplug.exporter_states_to_map(states)
    assert f(ids, "Healthy", hmap) == oplug.merge_health(ids, "Healthy", hmap) == ["Healthy", "Unhealthy", "Healthy"]
    assert f(ids, "Unhealthy", {}) == ["Unhealthy"] * 3


# ---- labeller ------------------------------------------------------------------------------------------
def test_label_keys_and_removal(P):
    assert P.labeller.labelGeneratorNames() == sorted(olab.GENERATOR_NAMES)
    labels = {"amd.com/gpu.cu-count": "104", "amd.com/gpu.vram": "64G", "beta.amd.com/gpu.cu-count": "104",
              "beta.amd.com/gpu.cu-count.104": "1", "beta.amd.com/gpu.family": "AI",
              "beta.amd.com/gpu.family.AI": "1", "amd.com/cpu": "true", "dummyLabel1": "1"}   # main_test.go:59-125
    assert P.labeller.removeOldNodeLabels(dict(labels)) == olab.removeOldNodeLabels(dict(labels)) \
        == {"amd.com/cpu": "true", "dummyLabel1": "1"}
    assert P.labeller.removeOldNodeLabels(None) is None
    for entries in ({"16G": 2}, {"16G": 2, "64G": 1}, {}):
        assert P.labeller.createLabels("vram", entries) == olab.createLabels("vram", entries)


@pytest.mark.parametrize("name,kw", [("cpx", dict(compute="cpx", memory="nps4")), ("mi308", dict(compute="spx", memory="nps1")),
                                     ("mi210", {})])
def test_generate_labels_match_oracle(P, kfd, tmp_path, name, kw):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, name), **kw)
    enabled = {n: True for n in olab.GENERATOR_NAMES}
    want = olab.generateLabels(enabled, root)
    with P.Context("kfd:" + root) as ctx:
        got = P.labeller.generateLabels(ctx, enabled)
        assert got == want and len(got) >= 10
        one = P.labeller.generateLabels(ctx, {"vram": True})
        assert one == olab.generateLabels({"vram": True}, root)


def test_synthetic_b200_tree(P, tmp_path):
    """The N x B200-shaped synthetic tree: oracle == kfd backend on devices, weights, allocation."""
    root = str(tmp_path / "b200")
    ids = P.synth.write_b200_tree(root, n_gpus=8)
    with P.Context("kfd:" + root) as ctx:
        got = ctx.enumerate()
        assert got == oamd.GetAMDGPUs(root) and sorted(got) == ids and len(ids) == 8
        assert ctx.start() == 0
        opol = oalloc.BestEffortPolicy()
        assert opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes") is None
        hist = collections.Counter(w for row in opol.p2pWeights.values() for w in row.values())
        assert hist == {40: 12, 50: 16}          # all NVLink-class: same-numa 20+10+10, cross-numa 20+10+20
        for size in range(1, 8):
            assert ctx.preferred_allocation(ids, [], size) == opol.Allocate(list(ids), [], size)[0]
    root2 = str(tmp_path / "mig")
    ids2 = P.synth.write_b200_tree(root2, n_gpus=8, partitions=7, compute_partition="1g", memory_partition="23gb")
    with P.Context("kfd:" + root2) as ctx:
        got = ctx.enumerate()
        assert got == oamd.GetAMDGPUs(root2) and len(got) == 56 == len(ids2)
        assert ctx.resource_list("mixed") == ["1g_23gb"]


# main_test.go:59-125, both cases verbatim, through the ABI and the oracle
REMOVE_OLD_CASES = [
    ({"amd.com/gpu.cu-count": "104", "amd.com/gpu.device-id": "740f", "amd.com/gpu.driver-version": "6.10.5",
      "amd.com/gpu.family": "AI", "amd.com/gpu.product-name": "Instinct_MI210", "amd.com/gpu.simd-count": "416",
      "amd.com/gpu.vram": "64G", "beta.amd.com/gpu.cu-count": "104", "beta.amd.com/gpu.cu-count.104": "1",
      "beta.amd.com/gpu.device-id": "740f", "beta.amd.com/gpu.device-id.740f": "1", "beta.amd.com/gpu.family": "HPC",
      "beta.amd.com/gpu.family.HPC": "1", "beta.amd.com/gpu.product-name": "Instinct_MI300X",
      "beta.amd.com/gpu.product-name.Instinct_MI300X": "1", "beta.amd.com/gpu.simd-count": "416",
      "beta.amd.com/gpu.simd-count.416": "1", "beta.amd.com/gpu.vram": "64G", "beta.amd.com/gpu.vram.64G": "1",
      "dummyLabel1": "1", "dummyLabel2": "2"},
     {"dummyLabel1": "1", "dummyLabel2": "2"}),
    ({"amd.com/cpu": "true", "amd.com/gpu": "true", "amd.com/mi300x": "true", "dummyLabel1": "1", "dummyLabel2": "2"},
     {"amd.com/cpu": "true", "amd.com/gpu": "true", "amd.com/mi300x": "true", "dummyLabel1": "1", "dummyLabel2": "2"}),
]


@pytest.mark.parametrize("labels,expect", REMOVE_OLD_CASES)
def test_remove_old_node_labels_reference_cases(P, labels, expect):
    assert olab.removeOldNodeLabels(dict(labels)) == expect
    assert P.labeller.removeOldNodeLabels(dict(labels)) == expect


def test_reconcile(P, kfd, tmp_path):
    """controller.go:23-58: stale labels of this labeller vanish, foreign labels stay, new ones land."""
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi210"))
    with P.Context("kfd:" + root) as ctx:
        new = P.labeller.generateLabels(ctx, {"vram": True, "cu-count": True, "device-id": True})
    node = dict(REMOVE_OLD_CASES[0][0])
    got = P.labeller.Reconcile(node, new)
    assert got == {**REMOVE_OLD_CASES[0][1], **new} and "beta.amd.com/gpu.family.HPC" not in got
    assert P.labeller.Reconcile(None, new) == new


def test_node_label_merge_patch(P):
    import json
    before = dict(REMOVE_OLD_CASES[0][0])
    after = P.labeller.Reconcile(dict(before), {"amd.com/gpu.vram": "179G", "beta.amd.com/gpu.vram": "179G",
                                                "beta.amd.com/gpu.vram.179G": "8"})
    patch = json.loads(P.labeller.node_label_merge_patch(before, after))["metadata"]["labels"]
    applied = {k: v for k, v in {**before, **patch}.items() if v is not None}
    assert applied == after
    assert patch["amd.com/gpu.family"] is None and patch["amd.com/gpu.vram"] == "179G" and "dummyLabel1" not in patch
    assert P.labeller.node_label_merge_patch(after, after) == '{"metadata":{"labels":{}}}'


def test_vendor_domain_is_configurable(P):
    try:
        P.labeller.setVendorDomain("nvidia.com")
        assert P.labeller.createLabels("vram", {"179G": 8}) == {
            "beta.nvidia.com/gpu.vram.179G": "8", "beta.nvidia.com/gpu.vram": "179G", "nvidia.com/gpu.vram": "179G"}
        assert P.labeller.removeOldNodeLabels({"nvidia.com/gpu.vram": "1G", "amd.com/gpu.vram": "1G"}) == {"amd.com/gpu.vram": "1G"}
        with pytest.raises(P._native.B2dpError):
            P.labeller.setVendorDomain("bad/domain")
    finally:
        P.labeller.setVendorDomain("amd.com")
    assert P.labeller.createLabelPrefix("vram", True) == "beta.amd.com/gpu.vram"


@pytest.mark.parametrize("n,mig", [(1, 1), (8, 1), (8, 7), (3, 2)])
def test_synthetic_backend_equals_the_written_tree(P, tmp_path, n, mig):
    """`synthetic:<N>[,mig=<k>]` (BASELINE config 5 ii) generates the same node shape as synth.write_b200_tree:
    device table, resource names, ListAndWatch list, pair weights, allocations and labels equal the kfd: backend
    on the written tree and the reference algorithm (oracle) on it; the generated tree is removed at close."""
    import glob
    root = str(tmp_path / "t")
    comp, mem = ("mig%d" % mig, "nps1") if mig > 1 else ("", "")
    ids = P.synth.write_b200_tree(root, n_gpus=n, partitions=mig, compute_partition=comp, memory_partition=mem)
    before = set(glob.glob("/dev/shm/b2dp_syn_*") + glob.glob("/tmp/b2dp_syn_*"))
    with P.Context("synthetic:%d,mig=%d" % (n, mig)) as syn, P.Context("kfd:" + root) as ref:
        made = set(glob.glob("/dev/shm/b2dp_syn_*") + glob.glob("/tmp/b2dp_syn_*")) - before
        assert len(made) == 1
        want = oamd.GetAMDGPUs(root)
        assert syn.enumerate() == ref.enumerate() == want and sorted(want) == ids
        assert syn.resource_list("mixed") == ref.resource_list("mixed") == oplug.getResourceList("mixed", root)[0]
        res = syn.resource_list("single")[0]
        assert syn.list_and_watch(res, P._native.LW_INITIAL)[0] == ref.list_and_watch(res, P._native.LW_INITIAL)[0]
        assert syn.node_health() == ref.node_health() is True
        assert syn.start() == ref.start() == (0 if len(ids) > 1 else P._native.E_ALLOC_NO_WEIGHTS)
        if len(ids) > 1:
            opol = oalloc.BestEffortPolicy()
            assert opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes") is None
            for size in sorted({1, 2, min(len(ids), 7), len(ids)}):
                got = syn.preferred_allocation(ids, [], size)
                assert got == ref.preferred_allocation(ids, [], size) == opol.Allocate(list(ids), [], size)[0]
        gens = P.labeller.labelGeneratorNames()
        assert syn.generate_labels(gens) == ref.generate_labels(gens) == olab.generateLabels({g: True for g in gens}, root)
        with pytest.raises(P._native.B2dpError) as ei:
            syn.probe_health()
        assert ei.value.code == P._native.E_UNSUPPORTED          # CPU only: no probe, no fallback
    assert not (set(glob.glob("/dev/shm/b2dp_syn_*") + glob.glob("/tmp/b2dp_syn_*")) - before)
    for bad in ("synthetic:0", "synthetic:8,mig=9", "synthetic:8,frob=1", "synthetic:abc"):
        with pytest.raises(P._native.B2dpError):
            P.Context(bad)


def test_struct_layouts_match_the_c_header(P, tmp_path):
    """Struct layouts are part of the ABI: sizeof/offsetof as a C compiler sees include/b200dp.h == the ctypes mirror
    (a cgo or ctypes host built against another layout would read garbage, so B2DP_ABI_VERSION moves with them)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    N = P._native
    pairs = [("b2dp_device", N.Device), ("b2dp_kv_count", N.KvCount), ("b2dp_label", N.Label), ("b2dp_devspec", N.DevSpec),
             ("b2dp_pair_weight", N.PairWeight), ("b2dp_link", N.Link), ("b2dp_fw_entry", N.FwEntry),
             ("b2dp_probe_opts", N.ProbeOpts), ("b2dp_probe_result", N.ProbeResult), ("b2dp_cycle_opts", N.CycleOpts),
             ("b2dp_cycle_stats", N.CycleStats), ("b2dp_p2p_opts", N.P2pOpts), ("b2dp_probe_info", N.ProbeInfo)]
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "b200dp.h"', 'int main(void) {']
    for cname, ct in pairs:
        src.append('printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in ct._fields_:
            src.append('printf(" %s=%%zu", offsetof(%s, %s));' % (fname, cname, fname))
        src.append('printf("\\n");')
    src += ['return 0; }']
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "layout")
    r = subprocess.run(["gcc", "-std=c11", "-I", os.path.join(REPO, "include"), str(c), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True).stdout.splitlines()
    assert len(out) == len(pairs)
    for line, (cname, ct) in zip(out, pairs):
        parts = line.split()
        assert parts[0] == cname and int(parts[1]) == ctypes.sizeof(ct), (cname, parts[1], ctypes.sizeof(ct))
        for tok, (fname, _) in zip(parts[2:], ct._fields_):
            k, v = tok.split("=")
            assert k == fname and int(v) == getattr(ct, fname).offset, (cname, fname, v, getattr(ct, fname).offset)
    assert ctypes.sizeof(N.ProbeResult) == 88 and ctypes.sizeof(N.CycleStats) == 64
