"""The CPU half of libb200dp.so (kfd.cpp, allocator.cpp, labels.cpp, ctx.cpp) rebuilt with plain g++
under AddressSanitizer + UBSan and under ThreadSanitizer (tests/native/cuda_stub.cpp stands in for
the CUDA backend) and driven through the public C ABI by tests/native/abi_stress.cpp:

* sweep   -- every CPU-side entry point, including the B2DP_E_NOSPC and bad-argument paths, over the
             reference's captured trees and 24 generated hostile trees (tests/test_fuzz_parity.gen_tree:
             decoy matches, overflowing integers, CRLF, malformed link files, reference-panic inputs);
* threads -- 6 threads on ONE shared context (enumerate / ListAndWatch / Start / GetPreferredAllocation /
             Allocate / labels) next to the library's own watch loop: no data race, and every pass
             returns the single-threaded answer.

The reference has no race/sanitizer build and carries a latent race on p.AMDGPUs
(plugin.go:231 vs :375); the library's contract is "callable from any thread" (include/b200dp.h).
The GPU side of the same contract is covered by compute-sanitizer (profiles/r01_compute_sanitizer_smoke.txt).
"""
import os
import random
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "k8s-device-plugin_b200", "csrc")
SOURCES = [os.path.join(CSRC, f) for f in ("kfd.cpp", "allocator.cpp", "labels.cpp", "ctx.cpp")] + [
    os.path.join(HERE, "native", "cuda_stub.cpp"), os.path.join(HERE, "native", "abi_stress.cpp")]


CACHE = os.path.join(HERE, "native", "_build")          # git-ignored; binaries are re-used while no source is newer
HEADERS = [os.path.join(ROOT, "include", "b200dp.h")] + [os.path.join(CSRC, f) for f in
           ("gosem.hpp", "internal.hpp", "pbwire.hpp", "hpack_huffman.inc", "host/h2grpc.hpp", "host/pbread.hpp",
            "units_backend.hpp", "nvml_dyn.hpp", "helper_proto.hpp", "pattern_math.hpp")]


def _fresh(out, sources):
    return os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in list(sources) + HEADERS)


def _build(out, flags, sources=None, werror=True):
    sources = sources or SOURCES
    if _fresh(out, sources):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fno-omit-frame-pointer", "-Wall"] + (["-Werror"] if werror else []) + \
          flags + list(sources) + ["-o", out + ".tmp%d" % os.getpid(), "-lpthread", "-ldl"]   # unique: xdist workers may build at once
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    os.replace(out + ".tmp%d" % os.getpid(), out)
    return out


@pytest.fixture(scope="module")
def bins(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    d = tmp_path_factory.mktemp("san")
    probe = subprocess.run(["g++", "-fsanitize=address,undefined", "-x", "c++", "-", "-o", str(d / "probe")],
                           input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("sanitizer runtimes not installed: " + probe.stderr[-200:])
    return {
        "asan": _build(os.path.join(CACHE, "abi_asan"), ["-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]),
        "tsan": _build(os.path.join(CACHE, "abi_tsan"), ["-fsanitize=thread"]),
    }


@pytest.fixture(scope="module")
def trees(kfd, tmp_path_factory):
    """Sysroots: the reference's captured topologies wrapped into a fake "/" + generated hostile trees."""
    sys.path.insert(0, HERE)
    import fake_sysfs
    from test_fuzz_parity import gen_tree
    d = tmp_path_factory.mktemp("trees")
    out = []
    for name, sub, kw in [("topology-parsing", "topology/nodes", {}),
                          ("topology-parsing-mi308", "topology/nodes", {"compute": "cpx", "memory": "nps1"}),
                          ("topo-mi210-xgmi-pcie", "nodes", {}),
                          ("topo-mi300-cpx", "topology/nodes", {"compute": "cpx", "memory": "nps4"})]:
        src = os.path.join(kfd.root(name), sub)
        assert os.path.isdir(src), src
        dst = str(d / ("fx_" + name))
        fake_sysfs.build(dst, src, **kw)
        out.append(dst)
    for seed in range(24):
        dst = str(d / ("fz_%d" % seed))
        gen_tree(dst, random.Random(seed))
        out.append(dst)
    out.append(str(d / "missing"))          # no driver dir at all: B2DP_E_NODRIVER
    return out


def _env():
    e = dict(os.environ)
    e["ASAN_OPTIONS"] = "detect_leaks=1:abort_on_error=0:halt_on_error=1"
    e["UBSAN_OPTIONS"] = "print_stacktrace=1:halt_on_error=1"
    e["TSAN_OPTIONS"] = "halt_on_error=1:second_deadlock_stack=1"
    return e


def test_asan_ubsan_sweep_over_fixtures_and_hostile_trees(bins, trees, tmp_path):
    r = subprocess.run([bins["asan"], "sweep", str(tmp_path)] + trees, capture_output=True, text=True, env=_env(),
                       timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-6000:])
    assert "sweep ok" in r.stdout and " 0 failures" in r.stdout
    assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr and "LeakSanitizer" not in r.stderr


@pytest.mark.parametrize("which", [0, 3])
def test_tsan_shared_context_many_threads(bins, trees, which):
    r = subprocess.run([bins["tsan"], "threads", trees[which], "6", "12" if which == 0 else "3"], capture_output=True,
                       text=True, env=_env(), timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-6000:])
    assert "threads ok" in r.stdout and " 0 mismatching passes, 0 failures" in r.stdout
    assert "ThreadSanitizer" not in r.stderr


@pytest.mark.parametrize("which", ["asan", "tsan"])
def test_mig_enumeration_under_sanitizers(bins, which, tmp_path):
    """The NVML-driven units backend (csrc/units_backend.hpp: MIG devices in the reference's CPX shape, declared links,
    Allocate specs with /dev/nvidia-caps nodes, labels) on an 8 x 7-instance node described by tests/native/nvml_stub.cpp:
    several threads on one shared `cuda:probe=off` context plus private contexts opening and closing, every pass equal
    to the single-threaded digest."""
    import test_mig_enumeration as tm
    stub = tm.STUB
    os.makedirs(tm.BUILD, exist_ok=True)
    src = os.path.join(HERE, "native", "nvml_stub.cpp")
    if not os.path.exists(stub) or os.path.getmtime(src) > os.path.getmtime(stub):
        tmp = "%s.tmp%d" % (stub, os.getpid())
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-shared", "-fPIC", "-fvisibility=hidden", src, "-o", tmp],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(tmp, stub)
    env = _env()
    env["B2DP_NVML_LIBRARY"] = stub
    env["B2DP_NVML_STUB"] = "gpus=8,mig=7"
    env["ASAN_OPTIONS"] += ":detect_leaks=0" if which == "tsan" else ""
    uri = "cuda:probe=off,cdi=nvidia.com/gpu,sysroot=" + tm._sysroot(tmp_path, 8, 7)
    r = subprocess.run([bins[which], "threads", uri, "4", "3"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-6000:])
    assert "threads ok" in r.stdout and " 0 mismatching passes, 0 failures" in r.stdout
    assert "Sanitizer" not in r.stderr and "runtime error" not in r.stderr


@pytest.mark.parametrize("which", ["asan", "tsan"])
def test_helper_fan_out_under_sanitizers(bins, which, tmp_path):
    """The parent side of probe=helpers (spawn, socketpair protocol, fan-out, reaping; csrc/units_backend.hpp) on a
    2 x 3-instance MIG node with protocol-speaking stand-in children: 3 threads x 12 passes on one context."""
    import test_mig_enumeration as tm
    for out, src, extra in ((tm.STUB, "nvml_stub.cpp", ["-shared", "-fPIC", "-fvisibility=hidden"]), (tm.FAKE, "fake_probe_helper.cpp", [])):
        os.makedirs(tm.BUILD, exist_ok=True)
        srcp = os.path.join(HERE, "native", src)
        if not os.path.exists(out) or os.path.getmtime(srcp) > os.path.getmtime(out):
            tmp = "%s.tmp%d" % (out, os.getpid())
            r = subprocess.run(["g++", "-std=c++17", "-O1"] + extra + [srcp, "-o", tmp], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            os.replace(tmp, out)
    env = _env()
    env.update(B2DP_NVML_LIBRARY=tm.STUB, B2DP_NVML_STUB="gpus=2,mig=3", B2DP_PROBE_HELPER=tm.FAKE)
    uri = "cuda:probe=helpers,mig_bytes=1048576,sysroot=" + tm._sysroot(tmp_path, 2, 3)
    r = subprocess.run([bins[which], "helpers", uri, "3", "12"], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-6000:])
    assert "helpers ok: 6 units, 3 threads, 36 passes, 0 bad, 0 failures" in r.stdout
    assert "Sanitizer" not in r.stderr and "runtime error" not in r.stderr


@pytest.mark.parametrize("san", ["address,undefined", "thread"])
def test_native_daemon_under_sanitizers(san, tmp_path):
    """b200dp_plugind (csrc/host: HTTP/2 + HPACK + gRPC, the watch threads, the signal loop) rebuilt under
    ASan+UBSan / TSan and put through tests/test_native_plugind.py again: grpcio as the kubelet on both
    sockets, concurrent streams, cancellation, kubelet restart, SIGUSR1, SIGTERM.  A sanitizer report makes
    the daemon exit non-zero, which those tests check."""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    probe = subprocess.run(["g++", "-fsanitize=" + san, "-x", "c++", "-", "-o", str(tmp_path / "probe")],
                           input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("sanitizer runtimes not installed")
    srcs = SOURCES[:-1] + [os.path.join(CSRC, "host", "plugind.cpp")]
    exe = _build(os.path.join(CACHE, "plugind_" + san.replace(",", "_")), ["-fsanitize=" + san], srcs, werror=False)
    env = _env()
    env["B200DP_PLUGIND"] = exe
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "test_native_plugind.py"), "-x", "-q",
                        "-p", "no:cacheprovider"], capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    assert r.returncode == 0, (r.stdout[-4000:], r.stderr[-2000:])
