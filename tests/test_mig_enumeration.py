"""BASELINE.json configs[4] (iii): LIVE MIG enumeration through NVML on the cuda: backend -- the B200 analogue of the
reference's amdgpu_xcp_* platform-device loop (amdgpu.go:221-265) -- exercised without hardware through a stand-in
NVML (tests/native/nvml_stub.cpp, loaded via B2DP_NVML_LIBRARY) that describes MIG-partitioned nodes.

Contract: the table the cuda: backend builds from NVML == the table of `synthetic:<N>,mig=<k>` (the generated kfd tree of
the same node) == what the reference algorithm (oracle) reads from the tree the backend exports; pair weights and the
preferred allocation of every size agree across the three; resource names, Allocate specs (parent node +
/dev/nvidia-caps nodes), NVIDIA_VISIBLE_DEVICES / CDI names (MIG UUIDs) and labels follow.  probe=off: no CUDA at all
(this container has no GPU)."""
import os
import shutil
import subprocess

import pytest

from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import labeller as olab
from oracle import plugin as oplug

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "native", "_build")
STUB = os.path.join(BUILD, "libnvml_stub.so")


@pytest.fixture(scope="module")
def stub():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "native", "nvml_stub.cpp")
    if not os.path.exists(STUB) or os.path.getmtime(src) > os.path.getmtime(STUB):
        tmp = "%s.tmp%d" % (STUB, os.getpid())                     # atomic: xdist workers may build at once
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-fvisibility=hidden", src, "-o", tmp],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(tmp, STUB)
    return STUB


def _sysroot(tmp_path, n_gpus, mig, numa_nodes=2):
    """What the backend reads next to NVML: numa_node + PCI device id per GPU, the NUMA node count, mig-minors."""
    root = tmp_path / "sysroot"
    shutil.rmtree(root, ignore_errors=True)
    for k in range(numa_nodes):
        (root / "sys/devices/system/node" / ("node%d" % k)).mkdir(parents=True)
    caps = []
    for g in range(n_gpus):
        d = root / "sys/bus/pci/devices" / ("0000:%02x:00.0" % (0x19 + 0x10 * g))
        d.mkdir(parents=True)
        (d / "numa_node").write_text("%d\n" % (0 if g < (n_gpus + 1) // 2 else 1))
        (d / "device").write_text("0x2901\n")
        minor = n_gpus - 1 - g                                     # the stub's minors are reversed
        for s in range(mig):
            caps.append("gpu%d/gi%d/access %d" % (minor, 7 + s, 100 + 20 * g + 2 * s))
            caps.append("gpu%d/gi%d/ci0/access %d" % (minor, 7 + s, 101 + 20 * g + 2 * s))
    (root / "proc/driver/nvidia-caps").mkdir(parents=True)
    (root / "proc/driver/nvidia-caps/mig-minors").write_text("\n".join(caps) + "\n")
    (root / "sys/module/nvidia").mkdir(parents=True)
    (root / "sys/module/nvidia/srcversion").write_text("STUBSRCVERSION\n")
    return str(root)


def _open(pkg, monkeypatch, stub, tmp_path, gpus, mig, extra="", migmask=None):
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=%d,mig=%d%s" % (gpus, mig, "" if migmask is None else ",migmask=%d" % migmask))
    return pkg.Context("cuda:probe=off,sysroot=%s%s" % (_sysroot(tmp_path, gpus, mig), extra))


def test_live_mig_equals_synthetic_tree_equals_oracle_on_export(pkg, monkeypatch, stub, tmp_path):
    """8 x B200, 7 x 1g.23gb each (B200's maximum): 56 schedulable devices."""
    with _open(pkg, monkeypatch, stub, tmp_path, 8, 7) as ctx, \
            pkg.Context("synthetic:8,mig=7,compute=1g,memory=23gb") as syn:
        devs = ctx.enumerate()
        assert len(devs) == 56 and list(devs) == sorted(devs)
        assert devs == syn.enumerate()                                            # (i) == the generated kfd tree
        root = str(tmp_path / "export")
        ctx.export_kfd_tree(root)
        assert devs == oamd.GetAMDGPUs(root)                                      # (ii) == the reference algorithm on the export
        # the reference's CPX shape: first instance under the PCI id, the others amdgpu_xcp_<8g+p>, one devID per GPU
        assert devs["0000:19:00.0"]["devID"] == devs["amdgpu_xcp_1"]["devID"] == devs["amdgpu_xcp_6"]["devID"] == "0000:19:00:0"
        assert devs["amdgpu_xcp_9"]["devID"] == "0000:29:00:0" and "amdgpu_xcp_8" not in devs and "amdgpu_xcp_7" not in devs
        assert {(v["computePartitionType"], v["memoryPartitionType"]) for v in devs.values()} == {("1g", "23gb")}
        assert ctx.partition_histogram() == {"1g_23gb": 56} == syn.partition_histogram()
        assert ctx.is_homogeneous() and ctx.resource_list("single") == ["gpu"]
        assert ctx.resource_list("mixed") == ["1g_23gb"] == oplug.getResourceList("mixed", root)[0]
        assert ctx.node_health() is True
        # Start(): declared links (no P2P between MIG instances) -> the same p2pWeights as the kfd reader / the oracle
        assert ctx.start() == 0 and syn.start() == 0
        opol = oalloc.BestEffortPolicy()
        assert opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes") is None
        assert ctx.pair_weights() == syn.pair_weights() == opol.p2pWeights
        assert len(ctx.pair_weights()) == 55
        ids = sorted(devs)
        for size in range(1, 57):                                                # every size, product vs product
            got = ctx.preferred_allocation(ids, [], size)
            assert got == syn.preferred_allocation(ids, [], size) and len(got) == size
        for size in (1, 2, 3, 7, 8):                                             # and vs the oracle where it is affordable
            sub = ids[:21]
            assert ctx.preferred_allocation(sub, [], size) == opol.Allocate(list(sub), [], size)[0]
        assert ctx.preferred_allocation(ids, ["amdgpu_xcp_9"], 7) == syn.preferred_allocation(ids, ["amdgpu_xcp_9"], 7)
        # ListAndWatch: one stream per resource, all Healthy at start, numa from the parent GPU
        wire, st = ctx.list_and_watch("1g_23gb", pkg._native.LW_INITIAL)
        homog, lw = oplug.list_and_watch_devices(oamd.GetAMDGPUs(root), "1g_23gb")
        msg = pkg.v1beta1.ListAndWatchResponse.FromString(wire)
        assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in msg.devices] == lw and st.n_devices == 56
        # heartbeat with probe=off: the default (node) health for everyone, like the reference without an exporter
        wire, st = ctx.list_and_watch("1g_23gb", pkg._native.LW_HEARTBEAT | pkg._native.LW_NO_PROBE)
        assert st.n_unhealthy == 0
        with pytest.raises(pkg._native.B2dpError) as ei:
            ctx.probe_health()
        assert ei.value.code == pkg._native.E_UNSUPPORTED


def test_mig_allocate_specs_and_runtime_ids(pkg, monkeypatch, stub, tmp_path):
    """Allocate for MIG devices: the parent's /dev/nvidia<minor> (once per GPU) and the instance's two
    /dev/nvidia-caps nodes; NVIDIA_VISIBLE_DEVICES / CDI names carry MIG UUIDs (or "<gpu>:<slot>" with
    id_strategy=index) -- never minors, which here differ from the NVML indices."""
    V = pkg.v1beta1
    with _open(pkg, monkeypatch, stub, tmp_path, 2, 3, ",cdi=nvidia.com/gpu") as ctx:
        devs = ctx.enumerate()
        assert sorted(devs) == ["0000:19:00.0", "0000:29:00.0", "amdgpu_xcp_1", "amdgpu_xcp_10", "amdgpu_xcp_2", "amdgpu_xcp_9"]
        specs = [s[0] for s in ctx.device_specs(["0000:19:00.0", "amdgpu_xcp_2", "amdgpu_xcp_9", "nope"])]
        assert specs == ["/dev/nvidiactl", "/dev/nvidia-uvm", "/dev/nvidia-uvm-tools",
                         "/dev/nvidia1", "/dev/nvidia-caps/nvidia-cap100", "/dev/nvidia-caps/nvidia-cap101",      # GPU 0 has minor 1
                         "/dev/nvidia-caps/nvidia-cap104", "/dev/nvidia-caps/nvidia-cap105",                       # same parent: node not repeated
                         "/dev/nvidia0", "/dev/nvidia-caps/nvidia-cap122", "/dev/nvidia-caps/nvidia-cap123"]
        resp = V.ContainerAllocateResponse.FromString(ctx.allocate_response(["amdgpu_xcp_2", "amdgpu_xcp_9"]))
        assert dict(resp.envs) == {"NVIDIA_VISIBLE_DEVICES": "MIG-00000000-0002-4000-8000-00000000b200,MIG-00000001-0001-4000-8000-00000000b200"}
        assert [c.name for c in resp.cdi_devices] == ["nvidia.com/gpu=MIG-00000000-0002-4000-8000-00000000b200",
                                                      "nvidia.com/gpu=MIG-00000001-0001-4000-8000-00000000b200"]
        assert [d.host_path for d in resp.devices] == [s for s in specs if s not in ("/dev/nvidia-caps/nvidia-cap100", "/dev/nvidia-caps/nvidia-cap101")]
        assert resp.SerializeToString() == ctx.allocate_response(["amdgpu_xcp_2", "amdgpu_xcp_9"])     # protobuf's own canonical bytes
        info = pkg._native.ProbeInfo()
        assert pkg._native.lib.b2dp_probe_describe(ctx._h, 3, info) == 0      # "amdgpu_xcp_10": GPU 1, third instance
        assert info.uuid.decode() == "MIG-00000001-0002-4000-8000-00000000b200" and info.sm_count == 18 and info.via_helper == 0
    with _open(pkg, monkeypatch, stub, tmp_path, 2, 3, ",id_strategy=index") as ctx:
        resp = V.ContainerAllocateResponse.FromString(ctx.allocate_response(["amdgpu_xcp_2", "0000:29:00.0"]))
        assert dict(resp.envs) == {"NVIDIA_VISIBLE_DEVICES": "0:2,1:0"}


def test_whole_gpus_through_nvml_and_mixed_node(pkg, monkeypatch, stub, tmp_path):
    """probe=off on whole GPUs (what a labeller-only pod uses: no CUDA context, no HBM ring), and a node where only
    some GPUs are MIG-enabled: heterogeneous => `single` refuses, `mixed` serves one resource per partition style."""
    with _open(pkg, monkeypatch, stub, tmp_path, 4, 0) as ctx, pkg.Context("synthetic:4") as syn:
        devs = ctx.enumerate()
        want = syn.enumerate()
        assert sorted(devs) == sorted(want)
        for k in devs:                                                          # same table but `card` = the REAL device minor
            assert {f: v for f, v in devs[k].items() if f != "card"} == {f: v for f, v in want[k].items() if f != "card"}
        assert [devs[k]["card"] for k in sorted(devs)] == [3, 2, 1, 0]
        assert ctx.resource_list("single") == ["gpu"]
        assert ctx.start() == 0
        ids = sorted(devs)
        syn.start()
        for size in range(1, 5):
            assert ctx.preferred_allocation(ids, [], size) == syn.preferred_allocation(ids, [], size)
        resp = pkg.v1beta1.ContainerAllocateResponse.FromString(ctx.allocate_response(ids[:2]))
        assert dict(resp.envs) == {"NVIDIA_VISIBLE_DEVICES": "GPU-00000000-0000-4000-8000-00000000b200,GPU-00000001-0000-4000-8000-00000000b200"}
        assert [d.host_path for d in resp.devices][-2:] == ["/dev/nvidia3", "/dev/nvidia2"]
        root = str(tmp_path / "export_whole")
        ctx.export_kfd_tree(root)
        assert devs == oamd.GetAMDGPUs(root)
        gens = ["driver-version", "driver-src-version", "device-id", "product-name", "vram", "simd-count", "cu-count",
                "compute-memory-partition", "compute-partitioning-supported", "memory-partitioning-supported", "family", "firmware"]
        got = ctx.generate_labels(gens)
        drm = {}
        for v in devs.values():
            b = "%s/sys/class/drm/card%d/device/" % (root, v["card"])
            fw = dict(ln.split(" ", 1) for ln in open(b + "b2dp_firmware").read().splitlines())
            fw = {k: "".join(ch if ch.isalnum() or ch in ".-_" else "_" for ch in ver.strip()) for k, ver in fw.items()}
            drm["card%d" % v["card"]] = {"family": open(b + "b2dp_family").read().strip(), "feat": {}, "fw": fw}
        assert got == olab.generateLabels({g: True for g in gens}, root, drm=drm)
        assert got["amd.com/gpu.cu-count"] == "148" and got["amd.com/gpu.vram"] == "179G" and got["amd.com/gpu.family"] == "Blackwell"
        assert got["beta.amd.com/gpu.firmware.vbios.fw.97.00.88.00.0F"] == "4" and got["beta.amd.com/gpu.firmware.gsp.fw.580.159.03"] == "4"
        assert got["beta.amd.com/gpu.firmware.inforom-pwr.fw.N_A_x"] == "4"     # label-safe version strings
        with pkg.Context("kfd:" + root) as kctx:
            assert kctx.generate_labels(gens) == got
    with _open(pkg, monkeypatch, stub, tmp_path, 4, 3, migmask=5) as ctx:       # GPUs 0 and 2 partitioned, 1 and 3 whole
        devs = ctx.enumerate()
        assert len(devs) == 3 + 1 + 3 + 1
        assert ctx.partition_histogram() == {"1g_23gb": 6, "7g_179gb": 2} and not ctx.is_homogeneous()
        with pytest.raises(pkg._native.B2dpError) as ei:
            ctx.resource_list("single")
        assert ei.value.code == pkg._native.E_HETEROGENEOUS
        assert ctx.resource_list("mixed") == ["1g_23gb", "7g_179gb"]
        wire, st = ctx.list_and_watch("7g_179gb", pkg._native.LW_INITIAL)
        assert st.n_devices == 2
        root = str(tmp_path / "export_mixed")
        ctx.export_kfd_tree(root)
        assert devs == oamd.GetAMDGPUs(root)
        wire, st = ctx.list_and_watch("1g_23gb", pkg._native.LW_INITIAL)
        assert st.n_devices == 6 and not st.homogeneous
        labels = ctx.generate_labels(["compute-memory-partition", "cu-count", "vram"])
        assert "amd.com/gpu.compute-memory-partition" not in labels             # heterogeneous: main.go:355-368 emits nothing
        assert labels["beta.amd.com/gpu.cu-count.18"] == "6" and labels["beta.amd.com/gpu.cu-count.148"] == "2"
        assert labels["beta.amd.com/gpu.vram.23G"] == "6"
    with pkg.Context("nvml:sysroot=%s" % _sysroot(tmp_path, 4, 3)) as ctx:       # the short name for cuda:probe=off
        assert len(ctx.enumerate()) == 8
    # mig=off lists the physical GPUs even when MIG mode is enabled
    with _open(pkg, monkeypatch, stub, tmp_path, 2, 3, ",mig=off") as ctx:
        assert sorted(ctx.enumerate()) == ["0000:19:00.0", "0000:29:00.0"]


def test_helpers_mode_needs_the_helper_binary_and_nvml(pkg, monkeypatch, stub, tmp_path):
    monkeypatch.setenv("B2DP_NVML_LIBRARY", "/nonexistent/libnvidia-ml.so")
    with pytest.raises(pkg._native.B2dpError) as ei:
        pkg.Context("cuda:probe=off")
    assert ei.value.code == pkg._native.E_NOGPU and "NVML" in str(ei.value)
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=1,mig=0")
    monkeypatch.setenv("B2DP_PROBE_HELPER", "/nonexistent/helper")
    with pytest.raises(pkg._native.B2dpError) as ei:
        pkg.Context("cuda:probe=helpers,sysroot=%s" % _sysroot(tmp_path, 1, 0))
    assert "probe helper executable not found" in str(ei.value)
    for bad in ("cuda:probe=sometimes", "cuda:mig=maybe", "cuda:mig_bytes=7"):
        with pytest.raises(pkg._native.B2dpError):
            pkg.Context(bad)


# ---- probe=helpers end to end on a MIG node, without a GPU: the parent side of csrc/units_backend.hpp against a
# protocol-speaking stand-in child (tests/native/fake_probe_helper.cpp) ----------------------------------------------------
FAKE = os.path.join(BUILD, "fake_probe_helper")


@pytest.fixture(scope="module")
def fake_helper():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "native", "fake_probe_helper.cpp")
    if not os.path.exists(FAKE) or os.path.getmtime(src) > os.path.getmtime(FAKE):
        tmp = "%s.tmp%d" % (FAKE, os.getpid())
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", src, "-o", tmp], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(tmp, FAKE)
    return FAKE


def _children():
    me, pids = os.getpid(), []
    for d in os.listdir("/proc"):
        if d.isdigit():
            try:
                stat = open("/proc/%s/stat" % d).read()
                if int(stat.rsplit(")", 1)[1].split()[1]) == me and "fake_probe_hel" in stat:
                    pids.append(int(d))
            except OSError:
                pass
    return sorted(pids)


def test_mig_units_probed_through_helper_processes(pkg, monkeypatch, stub, fake_helper, tmp_path):
    """A MIG node (2 GPUs x 3 instances) with probe=helpers (what mig=auto selects on such a node): one child per MIG
    instance, started with CUDA_VISIBLE_DEVICES=<MIG UUID>; the fan-out sends to all before waiting on any; each unit
    follows the seed schedule of its enumeration index and is judged against the host's closed form; ListAndWatch
    carries per-instance verdicts; a deadline miss is Unhealthy for that unit only and its late answer is discarded; a
    child that dies is Unhealthy once and restarted; xid=1 attributes a critical Xid to one GPU instance; closing
    reaps every child."""
    import time
    from oracle import probe as oprobe
    N = pkg._native
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=2,mig=3")
    monkeypatch.setenv("B2DP_PROBE_HELPER", fake_helper)
    mark = str(tmp_path / "died_once")
    monkeypatch.setenv("FAKE_HELPER_MARK", mark)
    monkeypatch.setenv("FAKE_HELPER_DIE_UNIT", "4")
    monkeypatch.setenv("FAKE_HELPER_DIE_AFTER", "3")
    monkeypatch.setenv("FAKE_HELPER_SLOW_UNIT", "2")
    monkeypatch.setenv("FAKE_HELPER_SLOW_MS", "0")
    sysroot = _sysroot(tmp_path, 2, 3)
    mig_bytes = 64 << 20
    with pkg.Context("cuda:sysroot=%s,mig_bytes=%d,xid=1" % (sysroot, mig_bytes)) as ctx:       # mig=auto => helpers
        ids = sorted(ctx.enumerate())
        assert ids == ["0000:19:00.0", "0000:29:00.0", "amdgpu_xcp_1", "amdgpu_xcp_10", "amdgpu_xcp_2", "amdgpu_xcp_9"]
        assert len(_children()) == 6
        d = ctx.probe_describe(3)                                     # "amdgpu_xcp_10" = GPU 1, third instance
        assert d["via_helper"] and d["usable"] and d["slot_bytes"] == mig_bytes and d["gbs_ref"] == 900.0
        assert d["uuid"] == "MIG-00000001-0002-4000-8000-00000000b200" and d["name"] == "FAKE B200 " + d["uuid"]   # the child saw ITS uuid
        seeds = [oprobe.initial_seed(i) for i in range(6)]
        for step in range(3):
            res = ctx.probe_health(timed=False)
            assert [r.device for r in res] == list(range(6)) and [r.seed for r in res] == seeds
            for r in res:
                assert r.err == 0 and r.healthy and r.bytes == 2 * mig_bytes and abs(r.min_gbs_applied - 720.0) < 1e-3
                assert r.checksum == r.expected_checksum == oprobe.expected_checksum(mig_bytes // 4, r.seed)
            seeds = [oprobe.next_seed(s) for s in seeds]
        # unit 4 dies instead of answering its 4th probe: Unhealthy (E_CUDA) for that instance only, restarted next pass
        res = ctx.probe_health(timed=False)
        assert [(r.healthy, r.err) for r in res] == [(True, 0)] * 4 + [(False, N.E_CUDA)] + [(True, 0)]
        wire, st = ctx.list_and_watch("1g_23gb", N.LW_HEARTBEAT)
        msg = pkg.v1beta1.ListAndWatchResponse.FromString(wire)
        assert [d_.health for d_ in msg.devices] == ["Healthy"] * 6 and st.n_devices == 6      # the restarted child answers
        assert len(_children()) == 6 and os.path.exists(mark)
        # fault on one instance: reported once
        ctx.probe_inject_fault(1, 777, 1)
        res = ctx.probe_health(timed=False)
        assert [r.healthy for r in res] == [True, False, True, True, True, True] and res[1].first_bad_word == 777
        assert all(r.healthy for r in ctx.probe_health(timed=False))
        # the 0.8 line, per instance
        ctx.probe_set_ref(5, 2000.0)
        res = ctx.probe_health(timed=False)
        assert [r.healthy for r in res] == [True] * 5 + [False] and res[5].flags & N.RES_SLOW and abs(res[5].min_gbs_applied - 1600.0) < 1e-3
        ctx.probe_set_ref(5, 0.0)
        assert all(r.healthy for r in ctx.probe_health(timed=False))
        # xid=1: a critical Xid is latched on the instance (sticky until reset); an application-level one is not
        UINT64_MAX = (1 << 64) - 1
        ctx.probe_inject_fault(2, UINT64_MAX, 31)
        assert all(r.healthy for r in ctx.probe_health(timed=False))
        ctx.probe_inject_fault(2, UINT64_MAX, 79)
        for _ in range(2):
            res = ctx.probe_health(timed=False)
            assert [r.healthy for r in res] == [True, True, False, True, True, True] and res[2].flags & N.RES_XID
        ctx.probe_reset(2)
        assert all(r.healthy for r in ctx.probe_health(timed=False))
        # Start() on declared links, allocation prefers instances of one GPU
        assert ctx.start() == 0
        got = ctx.preferred_allocation(ids, [], 3)
        assert len({ctx.enumerate()[i]["devID"] for i in got}) == 1
    time.sleep(0.2)
    assert _children() == []


def test_helper_deadline_and_stale_answers(pkg, monkeypatch, stub, fake_helper, tmp_path):
    """health.go:37 gives the exporter RPC a deadline; here a unit that misses `timeout_ms` is Unhealthy with
    B2DP_E_TIMEOUT while the others' verdicts stand, the caller returns at the deadline, and the late answer is
    recognised by its sequence number and dropped before the next verdict."""
    import time
    N = pkg._native
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=1,mig=2")
    monkeypatch.setenv("B2DP_PROBE_HELPER", fake_helper)
    monkeypatch.setenv("FAKE_HELPER_SLOW_UNIT", "1")
    monkeypatch.setenv("FAKE_HELPER_SLOW_MS", "400")
    with pkg.Context("cuda:sysroot=%s,mig_bytes=%d" % (_sysroot(tmp_path, 1, 2), 1 << 20)) as ctx:
        t0 = time.perf_counter()
        res = ctx.probe_health(timed=False, timeout_ms=100, min_gbs=1.0)
        waited = time.perf_counter() - t0
        assert [(r.healthy, r.err) for r in res] == [(True, 0), (False, N.E_TIMEOUT)] and waited < 0.35
        wire, st = ctx.list_and_watch("1g_23gb", N.LW_HEARTBEAT, timeout_ms=100, min_gbs=1.0)
        assert st.n_unhealthy == 1
        res = ctx.probe_health(timed=False, timeout_ms=5000, min_gbs=1.0)       # waits out the slow child: both fresh answers
        assert [(r.healthy, r.err) for r in res] == [(True, 0), (True, 0)]
        assert res[1].seed != res[0].seed


def test_helper_restart_does_not_stall_the_heartbeat(pkg, monkeypatch, stub, fake_helper, tmp_path):
    """A fresh CUDA process needs seconds to come up.  The heartbeat that finds a dead helper, and the ones that follow
    while its replacement starts, return promptly with that unit Unhealthy (the others' verdicts stand); the first
    heartbeat after the replacement answered HELLO puts the unit back in service, on the ceiling its predecessor had."""
    import time
    N = pkg._native
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=1,mig=3")
    monkeypatch.setenv("B2DP_PROBE_HELPER", fake_helper)
    monkeypatch.setenv("FAKE_HELPER_MARK", str(tmp_path / "died"))
    monkeypatch.setenv("FAKE_HELPER_DIE_UNIT", "1")
    monkeypatch.setenv("FAKE_HELPER_DIE_AFTER", "1")
    monkeypatch.setenv("FAKE_HELPER_RESTART_HELLO_MS", "900")
    logs = []
    N.set_log_callback(lambda level, msg: logs.append((level, msg)))       # what glog would have shown (b2dp_set_log_callback)
    try:
        _restart_scenario(pkg, tmp_path, logs)
    finally:
        N.set_log_callback(None)


def _restart_scenario(pkg, tmp_path, logs):
    import time
    N = pkg._native
    with pkg.Context("cuda:sysroot=%s,mig_bytes=%d" % (_sysroot(tmp_path, 1, 3), 1 << 20)) as ctx:
        assert any(lv == 0 and "3 unit(s), 3 MIG instance(s), probe=helpers (forced by MIG" in m for lv, m in logs), logs
        ctx.probe_set_ref(-1, 1000.0)
        assert all(r.healthy for r in ctx.probe_health(timed=False))
        res = ctx.probe_health(timed=False)                           # unit 1 dies instead of answering
        assert [(r.healthy, r.err) for r in res] == [(True, 0), (False, N.E_CUDA), (True, 0)]
        seen_starting = 0
        t_end = time.time() + 10
        while time.time() < t_end:
            t0 = time.perf_counter()
            res = ctx.probe_health(timed=False)
            assert time.perf_counter() - t0 < 0.6                     # never the 0.9 s the replacement takes
            assert res[0].healthy and res[2].healthy
            if res[1].err == 0:
                break
            assert res[1].err == N.E_CUDA and not ctx.probe_describe(1)["usable"]
            seen_starting += 1
            time.sleep(0.1)
        assert res[1].healthy and seen_starting >= 1 and abs(res[1].gbs_ref - 1000.0) < 1e-3     # the inherited ceiling
        assert ctx.probe_describe(1)["usable"]
        assert any(lv == 2 and "amdgpu_xcp_1: probe helper" in m and "exited" in m for lv, m in logs), logs
        assert any(lv == 0 and "amdgpu_xcp_1: probe helper" in m and "back in service" in m for lv, m in logs), logs
        # the reason a device flipped is logged with the heartbeat that flips it
        ctx.probe_inject_fault(2, 4242, 1)
        wire, st = ctx.list_and_watch("1g_23gb", N.LW_HEARTBEAT)
        assert st.n_unhealthy == 1
        assert any(lv == 1 and m.startswith("amdgpu_xcp_2 Unhealthy:") and "mismatches=1 first_bad_word=4242 checksum_BAD" in m for lv, m in logs), logs


@pytest.mark.parametrize("seed", range(10))
def test_random_mig_layouts_equal_the_oracle_on_the_export(pkg, monkeypatch, stub, tmp_path, seed):
    """Random nodes (1-8 GPUs, 1-7 instances, any subset of GPUs partitioned): the table built from NVML, the pair weights
    and the preferred allocations equal what the reference algorithm computes on the exported tree."""
    import random
    rnd = random.Random(1000 + seed)
    gpus, mig = rnd.randint(1, 8), rnd.randint(1, 7)
    mask = rnd.randint(0, (1 << gpus) - 1)
    with _open(pkg, monkeypatch, stub, tmp_path, gpus, mig, migmask=mask) as ctx:
        devs = ctx.enumerate()
        n_part = bin(mask).count("1")
        assert len(devs) == n_part * mig + (gpus - n_part)
        root = str(tmp_path / "export")
        ctx.export_kfd_tree(root)
        assert devs == oamd.GetAMDGPUs(root)
        ids = sorted(devs)
        rc = ctx.start()
        opol = oalloc.BestEffortPolicy()
        oerr = opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes")
        if len(ids) < 2:
            assert rc != 0 and oerr is not None
            return
        assert rc == 0 and oerr is None and ctx.pair_weights() == opol.p2pWeights
        sub = ids if len(ids) <= 12 else rnd.sample(ids, 12)         # the oracle's search is factorial in the group count
        for size in sorted({1, 2, min(3, len(sub)), min(5, len(sub)), len(sub)}):
            assert ctx.preferred_allocation(sub, [], size) == opol.Allocate(list(sub), [], size)[0], (gpus, mig, mask, size)
        must = sub[-1:]
        for size in (1, min(4, len(sub))):
            assert ctx.preferred_allocation(sub, must, size) == opol.Allocate(list(sub), list(must), size)[0]
        hist = ctx.partition_histogram()
        strategy_single_ok = len(hist) <= 1
        if strategy_single_ok:
            assert ctx.resource_list("single") == oplug.getResourceList("single", root)[0]
        assert ctx.resource_list("mixed") == oplug.getResourceList("mixed", root)[0]


def test_library_watch_loop_over_mig_helpers(pkg, monkeypatch, stub, fake_helper, tmp_path):
    """b2dp_watch_* (the library-owned ListAndWatch loop with its -pulse ticker, plugin.go:229-330) on a MIG node probed
    through helpers: initial list without a probe, then one verified cycle per 5 ms tick for the partition resource; a
    fault injected into one instance shows up in exactly one response."""
    import queue
    N = pkg._native
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=2,mig=2")
    monkeypatch.setenv("B2DP_PROBE_HELPER", fake_helper)
    got = queue.Queue()
    with pkg.Context("cuda:sysroot=%s,mig_bytes=%d" % (_sysroot(tmp_path, 2, 2), 1 << 20)) as ctx:
        w = ctx.watch(lambda rc, wire, st: got.put((rc, wire, st)), resource="1g_23gb", pulse_ms=5)
        try:
            rc, wire, st = got.get(timeout=10)
            assert rc == 0 and st.probe_bytes == 0 and st.n_devices == 4
            healths = []
            for i in range(8):
                if i == 3:
                    ctx.probe_inject_fault(2, 99, 1)
                rc, wire, st = got.get(timeout=10)
                assert rc == 0 and st.n_devices == 4 and st.probe_bytes == 4 * 2 * (1 << 20)
                healths.append([d.health for d in pkg.v1beta1.ListAndWatchResponse.FromString(wire).devices])
        finally:
            w.stop()
        bad = [h for h in healths if "Unhealthy" in h]
        assert len(bad) == 1 and bad[0] == ["Healthy", "Healthy", "Unhealthy", "Healthy"], healths


def test_mig_enabled_gpu_without_instances_lists_nothing(pkg, monkeypatch, stub, tmp_path):
    """A GPU with MIG mode enabled but no instance created has nothing to schedule: it is not listed (the whole GPU is
    not usable by CUDA in that state); a node where that leaves no device at all fails to open with a clear message."""
    monkeypatch.setenv("B2DP_NVML_LIBRARY", stub)
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=2,mig=3,empty=2")
    with pkg.Context("cuda:probe=off,sysroot=%s" % _sysroot(tmp_path, 2, 3)) as ctx:
        assert sorted(ctx.enumerate()) == ["0000:19:00.0", "amdgpu_xcp_1", "amdgpu_xcp_2"]
    monkeypatch.setenv("B2DP_NVML_STUB", "gpus=2,mig=3,empty=3")
    with pytest.raises(pkg._native.B2dpError) as ei:
        pkg.Context("cuda:probe=off,sysroot=%s" % _sysroot(tmp_path, 2, 3))
    assert ei.value.code == pkg._native.E_NOGPU and "no instance" in str(ei.value)
