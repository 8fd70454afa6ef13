#!/usr/bin/env python3
"""Random and structurally-valid-but-wild request bodies (Allocate, GetPreferredAllocation, PreStartContainer) over real gRPC
against the ASan+UBSan daemon on the reference's CPX capture: exercises the protobuf reader and the allocator's
validation paths.  Round 2: 39,869 calls in 60 s, no sanitizer report.

    python tools/fuzz_rpc_bodies.py        (CPU only)
"""
import os, random, subprocess, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib, grpc
pkg = importlib.import_module("k8s-device-plugin_b200")
import fake_sysfs, kfd_fixtures
from test_grpc_host import FakeKubelet
EXE = os.path.join(ROOT, "tests", "native", "_build", "plugind_address_undefined")
root = fake_sysfs.build(tempfile.mkdtemp() + "/r", os.path.join(kfd_fixtures.root("topo-mi300-cpx"), "topology/nodes"), compute="cpx", memory="nps4")
d = tempfile.mkdtemp(prefix="b2f_", dir="/tmp")
V = pkg.v1beta1
kubelet = FakeKubelet(os.path.join(d, "kubelet.sock"), V)
env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
proc = subprocess.Popen([EXE, "-pulse=0", "-backend=kfd:" + root, "-plugin_dir", d], stderr=subprocess.PIPE, text=True, env=env)
kubelet.requests.get(timeout=20)
sock = os.path.join(d, "amd.com_gpu")
rnd = random.Random(7)
ids = ["amdgpu_xcp_%d" % i for i in range(1, 64)] + ["0000:0c:00.0", "", "x" * 200, "\xff\xfe"]
def rand_msg():
    k = rnd.random()
    if k < 0.35: return os.urandom(rnd.choice([0, 1, 2, 3, 7, 30, 300, 5000]))
    if k < 0.7:   # structurally valid, semantically wild
        req = V.PreferredAllocationRequest()
        for _ in range(rnd.randint(0, 3)):
            c = req.container_requests.add()
            c.available_deviceIDs.extend(rnd.sample(ids, rnd.randint(0, 20)))
            c.must_include_deviceIDs.extend(rnd.sample(ids, rnd.randint(0, 5)))
            c.allocation_size = rnd.choice([-5, 0, 1, 2, 8, 63, 64, 1000, 2**31 - 1, -2**31])
        b = req.SerializeToString()
        if rnd.random() < 0.3 and b: b = b[:rnd.randint(0, len(b))]        # truncated
        return b
    req = V.AllocateRequest()
    for _ in range(rnd.randint(0, 4)):
        req.container_requests.add().devices_ids.extend(rnd.sample(ids, rnd.randint(0, 30)))
    return req.SerializeToString()
n = 0
with grpc.insecure_channel("unix://" + sock) as ch:
    t_end = time.time() + 60
    while time.time() < t_end and proc.poll() is None:
        method = rnd.choice([V.GET_PREFERRED_ALLOCATION, V.ALLOCATE, V.PRE_START_CONTAINER, V.GET_OPTIONS])
        try:
            ch.unary_unary(method, request_serializer=lambda b: b, response_deserializer=lambda b: b)(rand_msg(), timeout=5)
        except grpc.RpcError:
            pass
        n += 1
alive = proc.poll() is None
proc.terminate()
err = proc.communicate(timeout=20)[1]
print("calls", n, "alive", alive, "rc", proc.returncode)
print(err[-1500:] if ("Sanitizer" in err or "runtime error" in err) else "no sanitizer report")
kubelet.server.stop(0); shutil.rmtree(d, ignore_errors=True)
