"""Pin the oracle against every known-answer value the reference's own tests hold for the
hot path, on the reference's own captured kfd trees.

Reference tests mirrored here (file:line in /root/reference):
  internal/pkg/amdgpu/amdgpu_test.go:122-163   TestParseTopologyProperties
  internal/pkg/amdgpu/amdgpu_test.go:165-218   TestParseDebugFSFirmwareInfo
  internal/pkg/amdgpu/amdgpu_test.go:220-264   TestRenderDevIdsFromTopology (stale values; grouping only)
  internal/pkg/plugin/plugin_test.go:23-30     TestCountGPUDevFromTopology
  internal/pkg/allocator/device_test.go:80-169 pair weights / grouping / candidate counts
  internal/pkg/allocator/besteffort_policy_test.go:25-216 TestBestPolicyAllocator
  cmd/k8s-node-labeller/main_test.go:11-125    label keys / removeOldNodeLabels
"""
import collections

import pytest

from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import gosem
from oracle import labeller as olab
from oracle import plugin as oplug


# ---- amdgpu_test.go:122-163 -----------------------------------------------------------
def test_parse_topology_properties(kfd):
    tp = kfd.root("topology-parsing")
    re2 = gosem.compile_re2
    v, _ = oamd.ParseTopologyProperties(tp + "/topology/nodes/1/mem_banks/0/properties", re2(r"size_in_bytes\s(\d+)"))
    assert v == 17163091968
    v, _ = oamd.ParseTopologyProperties(tp + "/topology/nodes/1/mem_banks/0/properties", re2(r"flags\s(\d+)"))
    assert v == 0
    v, _ = oamd.ParseTopologyProperties(tp + "/topology/nodes/2/properties", re2(r"simd_count\s(\d+)"))
    assert v == 256
    v, _ = oamd.ParseTopologyProperties(tp + "/topology/nodes/2/properties", re2(r"simd_id_base\s(\d+)"))
    assert v == 2147487744
    _, e = oamd.ParseTopologyProperties(tp + "/topology/nodes/2/properties", re2(r"asdf\s(\d+)"))
    assert e is not None


def test_parse_topology_unique_id_overflow(kfd):
    # unique_id 14073402507705256556 > int64 => range error, clamped value
    tp = kfd.root("topology-parsing")
    v, e = oamd.ParseTopologyProperties(tp + "/topology/nodes/1/properties", gosem.compile_re2(r"unique_id\s(\d+)"))
    assert isinstance(e, gosem.ParseError) and e.kind == "range" and v == (1 << 63) - 1


# ---- amdgpu_test.go:165-218 -----------------------------------------------------------
EXP_FEAT = {"VCE": 0, "UVD": 0, "MC": 0, "ME": 35, "PFP": 35, "CE": 35, "RLC": 0, "MEC": 33, "MEC2": 33,
            "SOS": 0, "ASD": 0, "SMC": 0, "SDMA0": 40, "SDMA1": 40}
EXP_FW = {"VCE": 0x352D0400, "UVD": 0x01571100, "MC": 0, "ME": 0x94, "PFP": 0xA4, "CE": 0x4A, "RLC": 0x58,
          "MEC": 0x160, "MEC2": 0x160, "SOS": 0x161A92, "ASD": 0x16129A, "SMC": 0x1C2800, "SDMA0": 0x197,
          "SDMA1": 0x197}


def test_parse_debugfs_firmware_info(kfd):
    feat, fw = oamd.parseDebugFSFirmwareInfo(kfd.root("debugfs-parsing") + "/amdgpu_firmware_info")
    assert feat == EXP_FEAT and fw == EXP_FW


# ---- amdgpu_test.go:220-264 (values stale; key set + grouping authoritative) ------------
MI308_MINORS = [m for base in range(128, 192, 8) for m in range(base, base + 4)]
MI308_DEVIDS = ["0000:0a:00:0", "0000:80:00:0", "0000:a4:00:0", "0000:c8:00:0",
                "0001:0b:00:0", "0001:81:00:0", "0001:a5:00:0", "0001:c9:00:0"]


def test_render_dev_ids_from_topology_mi308(kfd):
    got = oamd.GetDevIdsFromTopology(kfd.root("topology-parsing-mi308"))
    assert sorted(got) == MI308_MINORS
    # grouping asserted by the (stale) reference test: 8 groups x 4 consecutive minors
    groups = collections.defaultdict(list)
    for minor, dev in got.items():
        groups[dev].append(minor)
    assert sorted(sorted(v) for v in groups.values()) == [list(range(b, b + 4)) for b in range(128, 192, 8)]
    # value format follows the shipped code (amdgpu.go:139-142), SURVEY B.2
    assert [got[b] for b in range(128, 192, 8)] == MI308_DEVIDS


def test_dev_ids_other_fixtures(kfd):
    assert oamd.GetDevIdsFromTopology(kfd.root("topology-parsing")) == {}       # no `domain` key
    cpx = oamd.GetDevIdsFromTopology(kfd.root("topo-mi300-cpx"))
    assert len(cpx) == 63
    assert sorted(set(cpx.values())) == ["0000:%s:00:0" % b for b in ["05", "27", "47", "65", "85", "a7", "c7", "e5"]]


def test_node_ids_from_topology(kfd):
    assert oamd.GetNodeIdsFromTopology(kfd.root("topology-parsing")) == {128: 1, 129: 2}
    mi = oamd.GetNodeIdsFromTopology(kfd.root("topology-parsing-mi308"))
    assert mi == {m: 2 + i for i, m in enumerate(MI308_MINORS)}
    cpx = oamd.GetNodeIdsFromTopology(kfd.root("topo-mi300-cpx"))
    assert cpx == {128 + i: 2 + i for i in range(63)}


# ---- plugin_test.go:23-30 ---------------------------------------------------------------
def test_count_gpu_dev_from_topology(kfd):
    assert oplug.countGPUDevFromTopology(kfd.root("topology-parsing")) == 2
    assert oplug.countGPUDevFromTopology(kfd.root("topology-parsing-mi308")) == 32
    assert oplug.countGPUDevFromTopology(kfd.root("topo-mi300-cpx")) == 63


def test_simple_health_check_logic(kfd):
    assert oplug.simpleHealthCheck(kfd.root("topology-parsing")) is False   # no gfx_target_version
    assert oplug.simpleHealthCheck(kfd.root("topology-parsing-mi308")) is True
    assert oplug.simpleHealthCheck(kfd.root("topo-mi300-cpx")) is True


# ---- device_test.go ---------------------------------------------------------------------
MI308 = dict(devCount=4, partitionCountPerDev=8, numanodeCount=2, startNodeId=2, endNodeId=33)
MI210 = dict(devCount=8, partitionCountPerDev=1, numanodeCount=2, startNodeId=2, endNodeId=9)
CPX = dict(devCount=8, partitionCountPerDev=8, numanodeCount=2, startNodeId=2, endNodeId=64)


def topo_dir(kfd, name):
    return {"mi308": kfd.root("topology-parsing-mi308") + "/topology/nodes",
            "mi210": kfd.root("topo-mi210-xgmi-pcie") + "/nodes",
            "cpx": kfd.root("topo-mi300-cpx") + "/topology/nodes"}[name]


def test_pair_weights_empty_devices(kfd):
    assert oalloc.fetchAllPairWeights([], {}, topo_dir(kfd, "mi308")) is not None   # device_test.go:80-88


def test_pair_weights_calculation(kfd):
    w = {}
    assert oalloc.fetchAllPairWeights(oalloc.getTestDevices(**MI308), w, topo_dir(kfd, "mi308")) is None
    assert len(w) == 31                                                               # device_test.go:105


def _hist(w):
    h = collections.Counter()
    for row in w.values():
        h.update(row.values())
    return dict(h), sum(k * v for k, v in h.items()), sum(len(r) for r in w.values())


@pytest.mark.parametrize("name,cfg,hist,total,pairs,rows", [
    ("mi308", MI308, {30: 112, 40: 128, 50: 256}, 21280, 496, 31),
    ("mi210", MI210, {40: 12, 80: 16}, 1760, 28, 7),
    ("cpx", CPX, {30: 217, 40: 744, 50: 992}, 85870, 1953, 62),
])
def test_pair_weight_histograms(kfd, name, cfg, hist, total, pairs, rows):
    w = {}
    oalloc.fetchAllPairWeights(oalloc.getTestDevices(**cfg), w, topo_dir(kfd, name))
    h, s, n = _hist(w)
    assert (h, s, n, len(w)) == (hist, total, pairs, rows)


def test_group_partitions_by_dev_id():
    assert len(oalloc.groupPartitionsByDevId(oalloc.getTestDevices(**MI308))) == 4   # device_test.go:110-123


@pytest.mark.parametrize("size,expected", [(3, 4), (12, 12)])
def test_get_subsets_method(kfd, size, expected):
    devices = oalloc.getTestDevices(**MI308)
    w = {}
    oalloc.fetchAllPairWeights(devices, w, topo_dir(kfd, "mi308"))
    subsets, err = oalloc.getCandidateDeviceSubsets(oalloc.groupPartitionsByDevId(devices), devices, list(devices),
                                                    [], size, w)
    assert err is None and len(subsets) == expected


# ---- besteffort_policy_test.go:25-216 ---------------------------------------------------
def X(*n):
    return ["amdgpu_xcp_%d" % i for i in n]


T = lambda *n: ["test%d" % i for i in n]  # noqa: E731
SAME_NUMA_AVAIL = T(3, 4, 5, 6, 7, 8)

# (topology, size, available|None, filtered, required, expectedIds|None, (score, candidates)|None)
BEST_POLICY_CASES = [
    ("mi308", 1, None, [], [], None, (0, 4)),
    ("mi308", 3, None, [], [], None, (90, 4)),
    ("mi308", 12, None, [], [], None, (2300, 12)),
    ("mi210", 1, None, [], [], T(1), (0, 8)),
    ("mi210", 3, None, [], [], T(1, 2, 3), (120, 336)),
    ("mi210", 5, None, [], [], T(1, 2, 3, 4, 5), (560, 6720)),
    ("mi210", 3, SAME_NUMA_AVAIL, [], [], T(5, 6, 7), (120, 120)),
    ("cpx", 1, None, [], [], T(8), (0, 8)),
    ("cpx", 3, None, [], [], T(8) + X(57, 58), (90, 8)),
    ("cpx", 5, None, [], [], T(8) + X(57, 58, 59, 60), (300, 8)),
    ("cpx", 3, SAME_NUMA_AVAIL, [], [], T(5, 6, 7), (120, 120)),
    ("cpx", 3, SAME_NUMA_AVAIL, [], T(5), T(5, 6, 7), (120, 20)),
    ("cpx", 30, None, [], [], None, (16410, 1680)),
    ("cpx", 8, None, [], [], T(1) + X(1, 2, 3, 4, 5, 6, 7), (840, 14)),
    ("cpx", 7, None, [], [], T(8) + X(57, 58, 59, 60, 61, 62), (630, 8)),
    ("cpx", 4, None, T(8) + X(57, 58), [], X(59, 60, 61, 62), (180, 8)),
    ("cpx", 10, None, T(1, 2, 3, 4, 8) + X(57), [], T(5) + X(33, 34, 35, 36, 37, 38, 39, 58, 59), (1510, 56)),
]
TOPO_CFG = {"mi308": MI308, "mi210": MI210, "cpx": CPX}


@pytest.mark.parametrize("topo,size,available,filtered,required,expected,score", BEST_POLICY_CASES)
def test_best_policy_allocator(kfd, topo, size, available, filtered, required, expected, score):
    devices = oalloc.getTestDevices(**TOPO_CFG[topo])
    pol = oalloc.BestEffortPolicy()
    assert pol.Init(devices, topo_dir(kfd, topo)) is None
    av = list(available) if available else [d.Id for d in devices]
    av = [a for a in av if a not in filtered]
    result, err = pol.Allocate(av, list(required), size)
    assert err is None and len(result) == size
    if expected is not None:
        assert sorted(result) == sorted(expected)
    if score is not None:                      # SURVEY Appendix B.2 regression pins
        assert (pol.last_score, pol.last_candidates) == score


def test_mi308_allocation_sets(kfd):
    devices = oalloc.getTestDevices(**MI308)
    pol = oalloc.BestEffortPolicy()
    pol.Init(devices, topo_dir(kfd, "mi308"))
    ids = [d.Id for d in devices]
    assert pol.Allocate(list(ids), [], 1)[0] == ["test1"]
    assert pol.Allocate(list(ids), [], 3)[0] == ["test1", "amdgpu_xcp_1", "amdgpu_xcp_2"]
    assert pol.Allocate(list(ids), [], 12)[0] == T(1) + X(1, 2, 3, 4, 5, 6, 7) + T(2) + X(9, 10, 11)


def test_allocate_validation_errors(kfd):
    devices = oalloc.getTestDevices(**MI210)
    pol = oalloc.BestEffortPolicy()
    ids = [d.Id for d in devices]
    assert str(pol.Allocate(ids, [], 2)[1]) == oalloc.invalidInit
    pol.Init(devices, topo_dir(kfd, "mi210"))
    assert str(pol.Allocate(ids, [], 0)[1]) == oalloc.invalidSize
    assert str(pol.Allocate(ids[:2], [], 3)[1]) == oalloc.invalidAvailable
    assert str(pol.Allocate(ids, ids[:3], 2)[1]) == oalloc.invalidRequired
    assert str(pol.Allocate(ids[:4], ["nope"], 2)[1]) == oalloc.noCandidateFound
    # shortcuts return the caller's own lists, unvalidated (besteffort_policy.go:110-116)
    assert pol.Allocate(["x", "y"], [], 2) == (["x", "y"], None)
    assert pol.Allocate(ids, ["q", "r"], 2) == (["q", "r"], None)
    with pytest.raises(gosem.GoPanic):
        pol.Allocate(ids + ["ghost"], [], 2)


# ---- main_test.go -----------------------------------------------------------------------
KINDS = ["family", "driver-version", "driver-src-version", "firmware", "device-id", "product-name", "vram",
         "simd-count", "cu-count", "compute-memory-partition", "compute-partitioning-supported",
         "memory-partitioning-supported"]


def test_init_label_lists():
    keys, exp = olab.initLabelLists()
    assert keys == sorted("amd.com/gpu." + k for k in KINDS)
    assert exp == sorted("beta.amd.com/gpu." + k for k in KINDS)


def test_remove_old_node_labels():
    # main_test.go:59-125
    labels = {"amd.com/gpu.cu-count": "104", "amd.com/gpu.vram": "64G", "beta.amd.com/gpu.cu-count": "104",
              "beta.amd.com/gpu.cu-count.104": "1", "beta.amd.com/gpu.family": "AI",
              "beta.amd.com/gpu.family.AI": "1", "amd.com/cpu": "true", "dummyLabel1": "1"}
    olab.removeOldNodeLabels(labels)
    assert labels == {"amd.com/cpu": "true", "dummyLabel1": "1"}
    assert olab.removeOldNodeLabels(None) is None


def test_create_labels_scheme():
    assert olab.createLabels("vram", {"16G": 2}) == {
        "beta.amd.com/gpu.vram.16G": "2", "beta.amd.com/gpu.vram": "16G", "amd.com/gpu.vram": "16G"}
    assert olab.createLabels("vram", {"16G": 2, "64G": 1}) == {
        "beta.amd.com/gpu.vram.16G": "2", "beta.amd.com/gpu.vram.64G": "1",
        "amd.com/gpu.vram.16G": "2", "amd.com/gpu.vram.64G": "1"}
    assert olab.vram_label_value(17163091968) == "16G"


# ---- Go stdlib semantics ----------------------------------------------------------------
def test_gosem_parse_int():
    assert gosem.parse_int(b"0", 0, 64) == 0
    assert gosem.parse_int(b"017", 0, 64) == 15                 # leading 0 => octal under base 0
    with pytest.raises(gosem.ParseError):
        gosem.parse_int(b"08", 0, 64)
    assert gosem.parse_int(b"0x352d0400", 0, 32) == 0x352D0400
    with pytest.raises(gosem.ParseError) as ei:
        gosem.parse_int(b"0x80000000", 0, 32)
    assert ei.value.kind == "range" and ei.value.value == (1 << 31) - 1
    assert gosem.atoi("-1") == -1 and gosem.atoi("+7") == 7
    assert gosem.atoi_ignore_err("12x") == 0


def test_gosem_glob_order(kfd):
    names = [p.split("/")[-2] for p in gosem.glob(kfd.root("topo-mi300-cpx") + "/topology/nodes/*/properties")]
    assert names == sorted(str(i) for i in range(65))           # lexical: 0,1,10,11,...
    assert names[:4] == ["0", "1", "10", "11"]
