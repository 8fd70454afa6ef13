"""The reference's own Go unit tests re-written in C++ against include/b200dp_host.hpp -- the compiled
host-side mirror of its packages (amdgpu, allocator, exporter, plugin, labeller) above the C ABI --
built with g++ against libb200dp.so and run on the reference's captured fixtures
(tests/native/host_mirror_test.cpp names the reference test behind each function)."""
import os
import shutil
import subprocess

import pytest

import fake_sysfs

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIBDIR = os.path.join(ROOT, "k8s-device-plugin_b200")


def test_reference_unit_tests_through_the_cpp_host_mirror(kfd, tmp_path):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "host_mirror_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(HERE, "native", "host_mirror_test.cpp"),
                        "-o", exe, "-L", LIBDIR, "-lb200dp", "-Wl,-rpath," + LIBDIR], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    sysroot = fake_sysfs.build(str(tmp_path / "cpx"), kfd.root("topo-mi300-cpx") + "/topology/nodes", compute="cpx",
                               memory="nps4")
    r = subprocess.run([exe, kfd.base_dir(), sysroot], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    assert "PASS" in r.stdout and ", 0 failed" in r.stdout
    for name in ("TestParseTopologyProperties", "TestParseDebugFSFirmwareInfo", "TestRenderDevIdsFromTopology",
                 "TestCountGPUDevFromTopology", "TestPairWeightCalculation", "TestGroupPartitionsByDevId",
                 "TestGetSubsetsMethod", "TestBestPolicyAllocator", "TestRemoveOldNodeLabels", "TestPluginFlowOnSysroot"):
        assert "ok   " + name in r.stdout


def test_hpack_rfc7541_vectors_and_grpc_framing(tmp_path):
    """csrc/host/h2grpc.hpp against the worked examples of RFC 7541 Appendix C (integers C.1, literals C.2,
    Huffman-coded requests sharing one dynamic table C.4) and its own malformed-input rejections."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "h2_hpack_test")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", os.path.join(HERE, "native", "h2_hpack_test.cpp"),
                        "-o", exe, "-lpthread"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "PASS" in r.stdout, (r.stdout, r.stderr[-3000:])
