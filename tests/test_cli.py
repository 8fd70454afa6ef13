"""The compiled C++ host (tools/b200dp_cli) over the C ABI == the oracle, on the kfd: backend. CPU only."""
import os
import subprocess

import pytest

import fake_sysfs
from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import labeller as olab
from oracle import plugin as oplug
from test_oracle_golden import topo_dir

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(REPO, "tools", "b200dp_cli")


def run(*args):
    r = subprocess.run([CLI, *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    return r.returncode, r.stdout, r.stderr


@pytest.mark.skipif(not os.path.exists(CLI), reason="tools/b200dp_cli not built (run __graft_entry__.build())")
def test_cli_matches_oracle(kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    uri = "kfd:" + root
    gpus = oamd.GetAMDGPUs(root)
    rc, out, _ = run(uri, "enumerate")
    assert rc == 0
    want = ["%s devID=%s card=%d renderD=%d node=%d numa=%d partition=%s_%s" % (
        k, v["devID"], v["card"], v["renderD"], v["nodeId"], v["numaNode"], v["computePartitionType"],
        v["memoryPartitionType"]) for k, v in sorted(gpus.items())]
    assert out.splitlines() == want
    rc, out, _ = run(uri, "resources", "mixed")
    assert (rc, out.split()) == (0, ["amd.com/cpx_nps4"])
    rc, out, _ = run(uri, "labels", "vram,simd-count")
    assert rc == 0
    assert dict(l.split("=", 1) for l in out.splitlines()) == olab.generateLabels({"vram": True, "simd-count": True}, root)
    opol = oalloc.BestEffortPolicy()
    opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes")
    rc, out, _ = run(uri, "alloc", "9")
    assert rc == 0 and out.split() == opol.Allocate(sorted(gpus), [], 9)[0]
    rc, out, _ = run(uri, "health")
    assert (rc, out.strip()) == (0, "node Healthy")
    rc, _, err = run(uri, "probe", "1")                      # no CPU fallback for the GPU probe
    assert rc == 1 and "not supported by this backend" in err
    rc, _, err = run("kfd:" + str(tmp_path / "nowhere"), "enumerate")
    assert rc == 1 and "amdgpu driver unavailable" in err     # the reference's Fatalf, as an error code
