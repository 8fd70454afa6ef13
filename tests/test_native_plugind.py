"""The native daemon (k8s-device-plugin_b200/b200dp_plugind: C++ HTTP/2 + HPACK + gRPC from
csrc/host/h2grpc.hpp, no Python, no gRPC library) against grpcio playing the kubelet: interop in both
directions -- grpcio's client (Huffman-coded, dynamically indexed HPACK, PINGs, flow control) drives the
native server; the native client registers with a grpcio Registration server -- and the same RPC
answers as the reference logic (oracle) on the reference's CPX capture.  CPU only (kfd: backend)."""
V_ABI = 3          # B2DP_ABI_VERSION
import os
import signal
import subprocess
import time

import grpc
import pytest

import fake_sysfs
from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import plugin as oplug
from test_grpc_host import FakeKubelet, _call
from test_oracle_golden import topo_dir

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
EXE = os.environ.get("B200DP_PLUGIND", os.path.join(ROOT, "k8s-device-plugin_b200", "b200dp_plugind"))


def _wait_for(path, timeout=10.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if os.path.exists(path):
            return True
        time.sleep(0.02)
    return False


@pytest.fixture
def daemon_env(pkg, kfd, tmp_path, short_dir):
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    return pkg.v1beta1, root, short_dir


def test_native_daemon_kubelet_round_trip(daemon_env):
    V, root, plug_dir = daemon_env
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    gpus = oamd.GetAMDGPUs(root)
    ids = sorted(gpus)
    proc = subprocess.Popen([EXE, "-pulse=1", "-resource_naming_strategy=single", "-backend=kfd:" + root,
                             "-plugin_dir", plug_dir], stderr=subprocess.PIPE, text=True)
    try:
        reg = kubelet.requests.get(timeout=30)                 # the native gRPC client -> grpcio server
        assert (reg.version, reg.endpoint, reg.resource_name) == ("v1beta1", "amd.com_gpu", "amd.com/gpu")
        assert reg.options.get_preferred_allocation_available and not reg.options.pre_start_required
        sock = os.path.join(plug_dir, "amd.com_gpu")
        assert _wait_for(sock)
        with grpc.insecure_channel("unix://" + sock) as ch:     # grpcio client -> the native gRPC server
            opts = _call(ch, V.GET_OPTIONS, V.Empty(), V.DevicePluginOptions)
            assert opts.get_preferred_allocation_available and not opts.pre_start_required
            stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                     response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
            first = next(stream)
            _, want = oplug.list_and_watch_devices(gpus, "gpu")
            assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in first.devices] == want
            t0 = time.time()
            second = next(stream)                               # the daemon's own -pulse=1 ticker
            third = next(stream)
            assert 0.5 < time.time() - t0 < 8.0                 # two 1 s ticks; generous upper bound for loaded hosts
            for m in (second, third):
                assert [d.ID for d in m.devices] == ids and all(d.health == "Healthy" for d in m.devices)
            proc.send_signal(signal.SIGUSR1)                    # operator's "heartbeat now" (timed in the soak test, no ticker there)
            assert [d.ID for d in next(stream).devices] == ids

            opol = oalloc.BestEffortPolicy()
            opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes")
            # repeated calls on one connection exercise grpcio's HPACK dynamic table (indexed re-use)
            for size, must in ((4, [ids[3]]), (8, []), (1, []), (9, [ids[0], ids[60]])):
                req = V.PreferredAllocationRequest(container_requests=[
                    V.ContainerPreferredAllocationRequest(available_deviceIDs=ids, must_include_deviceIDs=must,
                                                          allocation_size=size)])
                resp = _call(ch, V.GET_PREFERRED_ALLOCATION, req, V.PreferredAllocationResponse)
                assert list(resp.container_responses[0].deviceIDs) == opol.Allocate(list(ids), list(must), size)[0]
            bad = V.PreferredAllocationRequest(container_requests=[
                V.ContainerPreferredAllocationRequest(available_deviceIDs=ids[:2], allocation_size=5)])
            with pytest.raises(grpc.RpcError) as ei:
                _call(ch, V.GET_PREFERRED_ALLOCATION, bad, V.PreferredAllocationResponse)
            assert ei.value.code() == grpc.StatusCode.UNKNOWN
            assert ei.value.details() == "unable to get preferred allocation list. Error:" + oalloc.invalidAvailable

            areq = V.AllocateRequest(container_requests=[V.ContainerAllocateRequest(devices_ids=[ids[0], ids[1]]),
                                                         V.ContainerAllocateRequest(devices_ids=["bogus"])])
            aresp = _call(ch, V.ALLOCATE, areq, V.AllocateResponse)
            got = [[(d.host_path, d.container_path, d.permissions) for d in c.devices] for c in aresp.container_responses]
            assert got == [oplug.allocate_device_specs(gpus, [ids[0], ids[1]]), oplug.allocate_device_specs(gpus, ["bogus"])]
            assert _call(ch, V.PRE_START_CONTAINER, V.PreStartContainerRequest(devices_ids=ids[:2]),
                         V.PreStartContainerResponse) is not None
            with pytest.raises(grpc.RpcError) as ei:
                _call(ch, "/v1beta1.DevicePlugin/NoSuchMethod", V.Empty(), V.Empty)
            assert ei.value.code() == grpc.StatusCode.UNIMPLEMENTED
            # a second concurrent ListAndWatch stream on the same connection, then cancel both
            stream2 = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                      response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
            assert [d.ID for d in next(stream2).devices] == ids
            stream.cancel()
            stream2.cancel()
        # kubelet restart: kubelet.sock re-created -> the daemon serves again and re-registers (dpm/manager.go:73-84)
        kubelet.server.stop(0)
        time.sleep(0.1)
        if os.path.exists(os.path.join(plug_dir, "kubelet.sock")):
            os.unlink(os.path.join(plug_dir, "kubelet.sock"))
        kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
        reg = kubelet.requests.get(timeout=30)
        assert reg.resource_name == "amd.com/gpu"
        with grpc.insecure_channel("unix://" + sock) as ch:
            assert _call(ch, V.GET_OPTIONS, V.Empty(), V.DevicePluginOptions).get_preferred_allocation_available
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            _, err = proc.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            proc.kill()
            _, err = proc.communicate()
        kubelet.server.stop(0)
    assert proc.returncode == 0, err[-2000:]
    assert "Received signal, exiting" in err
    assert not os.path.exists(os.path.join(plug_dir, "amd.com_gpu"))      # socket removed on stop


def test_native_daemon_flag_and_start_errors(daemon_env):
    V, root, plug_dir = daemon_env
    r = subprocess.run([EXE, "-resource_naming_strategy=bogus", "-backend=kfd:" + root], capture_output=True, text=True)
    assert r.returncode == 1 and "invalid resource naming strategy: bogus" in r.stderr          # main.go:42-51
    # the reference image's glog flags are accepted (and ignored): Dockerfile:33
    r = subprocess.run([EXE, "-logtostderr=true", "-stderrthreshold=INFO", "-v=5", "-logtostderr", "-resource_naming_strategy=bogus"],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "invalid resource naming strategy: bogus" in r.stderr
    r = subprocess.run([EXE, "-version"], capture_output=True, text=True)
    assert r.returncode == 0 and "libb200dp ABI %d" % V_ABI in r.stdout
    r = subprocess.run([EXE, "-no_such_flag=1"], capture_output=True, text=True)
    assert r.returncode == 2 and "flag provided but not defined: -no_such_flag" in r.stderr
    r = subprocess.run([EXE, "-backend=kfd:" + root + "/nope"], capture_output=True, text=True)
    assert r.returncode == 1 and "amdgpu driver unavailable" in r.stderr                        # amdgpu.go:150-152
    # no kubelet: registration fails, start is retried 3 times (dpm/manager.go:16-20,205-219), the daemon stays up
    proc = subprocess.Popen([EXE, "-backend=kfd:" + root, "-plugin_dir", plug_dir, "-start_retry_wait=0.05"],
                            stderr=subprocess.PIPE, text=True)
    time.sleep(1.0)
    proc.send_signal(signal.SIGINT)
    _, err = proc.communicate(timeout=10)
    assert proc.returncode == 0 and "Failed to start plugin gpu: Register:" in err


def test_native_daemon_survives_hostile_peers(daemon_env):
    """Anything can connect to a unix socket: garbage instead of the preface, truncated and oversized frames,
    invalid HPACK (bad indices, bad Huffman padding, runaway integers), DATA on unknown streams, window
    overflow, a flood of random frames.  The daemon must drop such connections, stay up, and keep answering
    a well-behaved kubelet -- also when rebuilt under ASan/UBSan/TSan (tests/test_sanitizers.py)."""
    import random
    import socket
    import struct
    V, root, plug_dir = daemon_env
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    proc = subprocess.Popen([EXE, "-backend=kfd:" + root, "-plugin_dir", plug_dir], stderr=subprocess.PIPE, text=True)
    sock = os.path.join(plug_dir, "amd.com_gpu")
    PREFACE = b"PRI * HTTP/2.0\r\n\r\nSM\r\n\r\n"

    def frame(ftype, flags, stream, payload):
        return struct.pack(">I", len(payload))[1:] + bytes([ftype, flags]) + struct.pack(">I", stream) + payload

    def shoot(data, linger=0.05):
        s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        s.settimeout(2.0)
        s.connect(sock)
        try:
            s.sendall(data)
            time.sleep(linger)
            try:
                s.recv(65536)
            except (socket.timeout, ConnectionError):
                pass
        except (BrokenPipeError, ConnectionError):
            pass
        finally:
            s.close()

    try:
        kubelet.requests.get(timeout=30)
        assert _wait_for(sock)
        rng = random.Random(7)
        hdr_ok = frame(4, 0, 0, b"")                                            # SETTINGS
        cases = [
            b"GET / HTTP/1.1\r\n\r\n",                                          # not HTTP/2
            PREFACE[:10],                                                       # truncated preface, then close
            PREFACE + b"\xff" * 64,                                             # frame header with a 16 MiB length
            PREFACE + hdr_ok + frame(1, 0x4, 1, b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff"),  # runaway HPACK int
            PREFACE + hdr_ok + frame(1, 0x4, 1, b"\xbe"),                       # index 62 with an empty dynamic table
            PREFACE + hdr_ok + frame(1, 0x4, 1, b"\x00\x83\x00\x00\x00\x01a"),  # Huffman name with bad padding
            PREFACE + hdr_ok + frame(1, 0x4, 1, b"\x00\x05abc"),                # string longer than the block
            PREFACE + hdr_ok + frame(1, 0x0, 1, b"\x82") + frame(0, 0, 1, b"x"),  # DATA while a header block is open
            PREFACE + hdr_ok + frame(0, 0x1, 7, b"\x00\x00\x00\x00\x01z"),      # DATA on a stream that never opened
            PREFACE + hdr_ok + frame(8, 0, 0, b"\x7f\xff\xff\xff") * 4,         # window increments past 2^31
            PREFACE + hdr_ok + frame(4, 0, 0, b"\x00\x04"),                     # SETTINGS with a bad length
            PREFACE + hdr_ok + frame(6, 0, 0, b"1234"),                         # PING with a bad length
            PREFACE + hdr_ok + frame(1, 0x5 | 0x8, 1, b"\xf0" + b"\x82"),       # pad length larger than the payload
            PREFACE + hdr_ok + frame(1, 0x5, 1, b"\x83\x86\x44\x01/") + frame(3, 0, 1, b"\x00\x00\x00\x08"),  # GET + RST
        ]
        cases.append(PREFACE + hdr_ok + b"".join(frame(1, 0x4, 1 + 2 * i, b"\x83\x86\x44\x01/") for i in range(400)))  # 400 streams never closed
        cases.append(PREFACE + hdr_ok + frame(1, 0x0, 1, b"\x00\x01a\x01b") + frame(9, 0x0, 1, b"\x00\x01a\x01b" * 3000) * 8)  # endless CONTINUATION
        for _ in range(40):                                                     # random frame soup
            blob = PREFACE + hdr_ok
            for _ in range(rng.randint(1, 12)):
                blob += frame(rng.randint(0, 12), rng.randint(0, 255), rng.choice([0, 1, 1, 3, 2 ** 31 - 1]),
                              bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 40))))
            cases.append(blob)
        for c in cases:
            shoot(c)
            assert proc.poll() is None, proc.stderr.read()[-3000:]
        # random bytes where a protobuf request belongs: a gRPC status (or an empty answer), never a crash;
        # and the well-behaved kubelet is still served
        with grpc.insecure_channel("unix://" + sock) as ch:
            for method in (V.ALLOCATE, V.GET_PREFERRED_ALLOCATION, V.PRE_START_CONTAINER):
                raw = ch.unary_unary(method, request_serializer=lambda b: b, response_deserializer=lambda b: b)
                for _ in range(60):
                    blob = bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 48)))
                    try:
                        raw(blob, timeout=5)
                    except grpc.RpcError as e:
                        assert e.code() in (grpc.StatusCode.INTERNAL, grpc.StatusCode.UNKNOWN), e
                assert proc.poll() is None
            assert _call(ch, V.GET_OPTIONS, V.Empty(), V.DevicePluginOptions).get_preferred_allocation_available
            stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                     response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
            assert len(next(stream).devices) == 63
            stream.cancel()
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            _, err = proc.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            proc.kill()
            _, err = proc.communicate()
        kubelet.server.stop(0)
    assert proc.returncode == 0, err[-3000:]


def test_native_daemon_exporter_socket_on_kfd_is_a_clean_error(daemon_env):
    """-exporter_socket serves metricssvc.MetricsService from the HBM probe; the kfd: backend has no probe and
    no CPU fallback, so List answers with a gRPC error (the reference's client then uses the default health,
    health.go:66-69) instead of inventing verdicts."""
    import importlib
    srv_mod = importlib.import_module("k8s-device-plugin_b200.server")
    V, root, plug_dir = daemon_env
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    exp = os.path.join(plug_dir, "exp", "m.sock")
    proc = subprocess.Popen([EXE, "-backend=kfd:" + root, "-plugin_dir", plug_dir, "-exporter_socket", exp],
                            stderr=subprocess.PIPE, text=True)
    try:
        kubelet.requests.get(timeout=30)
        assert _wait_for(exp)
        with grpc.insecure_channel("unix://" + exp) as ch:
            call = ch.unary_unary(srv_mod.METRICS_LIST, request_serializer=lambda b: b,
                                  response_deserializer=srv_mod.GPUStateResponse.FromString)
            with pytest.raises(grpc.RpcError) as ei:
                call(b"", timeout=5.0)
            assert ei.value.code() == grpc.StatusCode.UNKNOWN
    finally:
        proc.send_signal(signal.SIGTERM)
        _, err = proc.communicate(timeout=10)
        kubelet.server.stop(0)
    assert proc.returncode == 0, err[-2000:]
    assert not os.path.exists(exp)


def test_native_daemon_soak_streams_come_and_go(daemon_env):
    """A kubelet that reconnects over and over (each time: options, a ListAndWatch stream, a few heartbeats,
    an allocation, cancel) next to a storm of SIGUSR1 heartbeats: answers stay right, finished streams are
    reaped (no growing thread count), memory stays flat; under ASan the exit-time leak check runs as well."""
    V, root, plug_dir = daemon_env
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    proc = subprocess.Popen([EXE, "-backend=kfd:" + root, "-plugin_dir", plug_dir], stderr=subprocess.PIPE, text=True)
    sock = os.path.join(plug_dir, "amd.com_gpu")

    def threads_and_rss():
        with open("/proc/%d/status" % proc.pid) as f:
            st = dict(line.split(":", 1) for line in f if ":" in line)
        return int(st["Threads"].split()[0]), int(st["VmRSS"].split()[0])

    try:
        kubelet.requests.get(timeout=30)
        assert _wait_for(sock)
        marks = []
        for it in range(60):
            with grpc.insecure_channel("unix://" + sock) as ch:
                assert _call(ch, V.GET_OPTIONS, V.Empty(), V.DevicePluginOptions).get_preferred_allocation_available
                stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                         response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
                first = next(stream)
                assert len(first.devices) == 63
                for _ in range(5):
                    proc.send_signal(signal.SIGUSR1)
                    assert len(next(stream).devices) == 63
                ids = [d.ID for d in first.devices]
                areq = V.AllocateRequest(container_requests=[V.ContainerAllocateRequest(devices_ids=ids[:3])])
                assert len(_call(ch, V.ALLOCATE, areq, V.AllocateResponse).container_responses[0].devices) == 7
                stream.cancel()
            if it in (9, 59):
                proc.send_signal(signal.SIGUSR1)        # lets the daemon reap cancelled streams
                time.sleep(0.2)
                marks.append(threads_and_rss())
        (t0, rss0), (t1, rss1) = marks
        assert t1 <= t0 + 4, marks                       # watch / connection threads do not accumulate
        if "B200DP_PLUGIND" not in os.environ:           # sanitizer builds: ASan's quarantine keeps freed memory mapped;
            assert rss1 <= rss0 * 1.5 + 8192, marks      # (kB) their exit-time leak check is the memory test there
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            _, err = proc.communicate(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill()
            _, err = proc.communicate()
        kubelet.server.stop(0)
    assert proc.returncode == 0, err[-3000:]


def test_native_daemon_mixed_strategy_two_resources(pkg, kfd, tmp_path, short_dir):
    """-resource_naming_strategy=mixed on a heterogeneous node (main.go:62-89): one socket and one registration
    per "<compute>_<memory>" resource, each stream carrying only its own devices; `single` refuses such a node."""
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    V = pkg.v1beta1
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi308"), compute="cpx", memory="nps1",
                            hetero_second=("spx", "nps1"))
    gpus = oamd.GetAMDGPUs(root)
    r = subprocess.run([EXE, "-resource_naming_strategy=single", "-backend=kfd:" + root, "-plugin_dir", short_dir],
                       capture_output=True, text=True, timeout=30)
    assert r.returncode == 1 and "Partitions of different styles" in r.stderr                   # main.go:79
    kubelet = FakeKubelet(os.path.join(short_dir, "kubelet.sock"), V)
    proc = subprocess.Popen([EXE, "-resource_naming_strategy=mixed", "-backend=kfd:" + root, "-plugin_dir", short_dir],
                            stderr=subprocess.PIPE, text=True)
    try:
        regs = sorted((kubelet.requests.get(timeout=30) for _ in range(2)), key=lambda q: q.resource_name)
        assert [q.resource_name for q in regs] == ["amd.com/cpx_nps1", "amd.com/spx_nps1"]
        assert [q.endpoint for q in regs] == ["amd.com_cpx_nps1", "amd.com_spx_nps1"]
        for q in regs:
            res = q.resource_name.split("/")[1]
            _, want = oplug.list_and_watch_devices(gpus, res)
            with grpc.insecure_channel("unix://" + os.path.join(short_dir, q.endpoint)) as ch:
                stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                         response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
                first = next(stream)
                assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in first.devices] == want and want
                stream.cancel()
    finally:
        proc.send_signal(signal.SIGTERM)
        _, err = proc.communicate(timeout=10)
        kubelet.server.stop(0)
    assert proc.returncode == 0, err[-2000:]


def test_native_kubelet_sim_drives_the_native_daemon():
    """tools/b200dp_kubelet_sim: the kubelet's side (Registration server + ListAndWatch streaming client) on the
    same native gRPC, against the daemon on a generated 8-GPU node -- native to native, no Python on either side."""
    import json
    sim = os.path.join(ROOT, "tools", "b200dp_kubelet_sim")
    if not (os.path.exists(sim) and os.path.exists(EXE)):
        import __graft_entry__
        __graft_entry__.build()
    r = subprocess.run([sim, EXE, "synthetic:8,mig=7", "100"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, (r.stdout, r.stderr[-2000:])
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["resource"] == "amd.com/gpu" and d["iterations"] == 100 and d["response_bytes"] > 56 * 20
    assert 0 < d["heartbeat_to_kubelet_ms_median"] < 50


def test_native_daemon_labeller_mode(daemon_env, pkg, monkeypatch, tmp_path):
    """`b200dp_plugind -labels=...` is the node labeller without Python (cmd/k8s-node-labeller/main.go:383-479 minus
    the controller-runtime client): the generated map equals the oracle's on the CPX capture; -reconcile applies
    controller.go:23-58 to a node's label map read from stdin; -patch prints the JSON merge patch a K8s client sends."""
    import json
    from oracle import labeller as olab
    V, root, plug_dir = daemon_env
    gens = ["vram", "cu-count", "simd-count", "compute-memory-partition", "device-id", "driver-version"]
    r = subprocess.run([EXE, "-labels=" + ",".join(gens), "-backend=kfd:" + root], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = olab.generateLabels({g: True for g in gens}, root)
    assert json.loads(r.stdout) == want and want["amd.com/gpu.compute-memory-partition"] == "cpx_nps4"
    node = {"keep": "me", "beta.amd.com/gpu.vram": "1G", "beta.amd.com/gpu.vram.1G": "8", "amd.com/gpu.cu-count": "7",
            "quote\"d": "tab\there", "amd.com/gpu.compute-memory-partition": "cpx_nps4"}
    r = subprocess.run([EXE, "-labels=" + ",".join(gens), "-reconcile", "-backend=kfd:" + root], input=json.dumps(node),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    after = olab.Reconcile(dict(node), want) if hasattr(olab, "Reconcile") else None
    lab = pkg.labeller
    assert json.loads(r.stdout) == lab.Reconcile(dict(node), dict(want)) and (after is None or after == json.loads(r.stdout))
    r = subprocess.run([EXE, "-labels=" + ",".join(gens), "-patch", "-backend=kfd:" + root], input=json.dumps(node),
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert json.loads(r.stdout) == json.loads(lab.node_label_merge_patch(node, lab.Reconcile(dict(node), dict(want))))
    patch = json.loads(r.stdout)["metadata"]["labels"]
    assert patch["beta.amd.com/gpu.vram.1G"] is None and "keep" not in patch and "amd.com/gpu.compute-memory-partition" not in patch
    r = subprocess.run([EXE, "-labels=vram", "-patch", "-backend=kfd:" + root], input="[1,2]", capture_output=True, text=True)
    assert r.returncode == 2 and "JSON object" in r.stderr
    r = subprocess.run([EXE, "-labels=vram", "-reconcile", "-backend=kfd:" + root], input="null", capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout) == olab.generateLabels({"vram": True}, root)
    # `-labels` (all generators) on the cuda: backend needs neither a CUDA context nor an HBM ring: NVML answers
    # (probe=off is implied) -- here a MIG-partitioned node described by the stand-in NVML
    import test_mig_enumeration as tm
    stub = tm.STUB
    if not os.path.exists(stub):
        pytest.skip("nvml stub not built (tests/test_mig_enumeration.py builds it)")
    env = dict(os.environ, B2DP_NVML_LIBRARY=stub, B2DP_NVML_STUB="gpus=2,mig=3")
    r = subprocess.run([EXE, "-labels", "-backend=cuda:sysroot=" + tm._sysroot(tmp_path, 2, 3)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    got = json.loads(r.stdout)
    assert got["amd.com/gpu.compute-memory-partition"] == "1g_23gb" and got["amd.com/gpu.cu-count"] == "18"
    assert got["amd.com/gpu.p2p-link"] == "nvlink" and got["amd.com/gpu.family"] == "Blackwell"
    assert got["amd.com/gpu.driver-version"] == "580.159.03" and got["beta.amd.com/gpu.vram.23G"] == "6"


def test_deploy_manifests_use_flags_the_daemon_accepts():
    """deploy/*.yaml and the Helm chart's values: every daemon argument parses (`-version` after them exits 0 only
    if the flag parser accepted all of them), the backend URIs are well-formed option lists, and the chart carries the
    reference chart's four templates (helm/amd-gpu/templates)."""
    import re
    import yaml
    dep = os.path.join(ROOT, "deploy")
    seen = 0
    for fn in ("k8s-ds-b200-dp-health.yaml", "k8s-ds-b200-labeller.yaml"):
        for doc in yaml.safe_load_all(open(os.path.join(dep, fn))):
            if not doc or doc.get("kind") != "DaemonSet":
                continue
            for c in doc["spec"]["template"]["spec"]["containers"]:
                args = list(c.get("args", []))
                assert c.get("command", ["/opt/b200dp/k8s-device-plugin_b200/b200dp_plugind"])[0].endswith("b200dp_plugind")
                r = subprocess.run([EXE] + args + ["-version"], capture_output=True, text=True)
                assert r.returncode == 0, (fn, args, r.stderr)
                for a in args:
                    if a.startswith("-backend=cuda:"):
                        assert re.fullmatch(r"-backend=cuda:([a-z_]+=[^,]+(,[a-z_]+=[^,]+)*)?", a), a
                seen += 1
    assert seen == 2
    chart = os.path.join(dep, "helm", "b200-gpu")
    vals = yaml.safe_load(open(os.path.join(chart, "values.yaml")))
    assert vals["labeller"]["enabled"] is False and vals["dp"]["pulse"] == 10 and "min_frac=0.8" in vals["dp"]["backend"]
    assert yaml.safe_load(open(os.path.join(chart, "Chart.yaml")))["name"] == "b200-gpu"
    for t in ("deviceplugin-daemonset.yaml", "labeller.yaml", "rbac.yaml", "serviceaccount.yaml", "_helpers.tpl", "NOTES.txt"):
        assert os.path.exists(os.path.join(chart, "templates", t)), t


def test_native_daemon_on_a_mig_node_end_to_end(pkg, short_dir, tmp_path):
    """BASELINE configs[4] at the kubelet's sockets: the native daemon on a node where GPU 0 is MIG-partitioned
    (3 x 1g.23gb) and GPU 1 is whole (stand-in NVML + protocol-speaking stand-in probe helpers, no GPU): `mixed` strategy
    registers one resource per partition style, each ListAndWatch stream carries its own devices, heartbeats run the
    probe through one helper per unit, Allocate names MIG UUIDs and mounts the /dev/nvidia-caps nodes, and
    GetPreferredAllocation keeps instances of one GPU together."""
    import shutil
    import test_mig_enumeration as tm
    V = pkg.v1beta1
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    for out, src, extra in ((tm.STUB, "nvml_stub.cpp", ["-shared", "-fPIC", "-fvisibility=hidden"]), (tm.FAKE, "fake_probe_helper.cpp", [])):
        os.makedirs(tm.BUILD, exist_ok=True)
        srcp = os.path.join(HERE, "native", src)
        if not os.path.exists(out) or os.path.getmtime(srcp) > os.path.getmtime(out):
            tmp = "%s.tmp%d" % (out, os.getpid())
            r = subprocess.run(["g++", "-std=c++17", "-O1"] + extra + [srcp, "-o", tmp], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr
            os.replace(tmp, out)
    env = dict(os.environ, B2DP_NVML_LIBRARY=tm.STUB, B2DP_NVML_STUB="gpus=2,mig=3,migmask=1", B2DP_PROBE_HELPER=tm.FAKE)
    plug_dir = short_dir
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    proc = subprocess.Popen([EXE, "-pulse=0", "-resource_naming_strategy=mixed", "-plugin_dir", plug_dir,
                             "-backend=cuda:mig_bytes=1048576,cdi=nvidia.com/gpu,sysroot=" + tm._sysroot(tmp_path, 2, 3)],
                            stderr=subprocess.PIPE, text=True, env=env)
    try:
        regs = sorted((kubelet.requests.get(timeout=30) for _ in range(2)), key=lambda r: r.resource_name)
        assert [r.resource_name for r in regs] == ["amd.com/1g_23gb", "amd.com/7g_179gb"]
        assert [r.endpoint for r in regs] == ["amd.com_1g_23gb", "amd.com_7g_179gb"]
        mig_ids = ["0000:19:00.0", "amdgpu_xcp_1", "amdgpu_xcp_2"]
        with grpc.insecure_channel("unix://" + os.path.join(plug_dir, "amd.com_1g_23gb")) as ch:
            stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                     response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
            first = next(stream)
            assert [(d.ID, d.health) for d in first.devices] == [(i, "Healthy") for i in mig_ids]
            proc.send_signal(signal.SIGUSR1)                                  # heartbeat: the probe runs in the helpers
            beat = next(stream)
            assert [(d.ID, d.health) for d in beat.devices] == [(i, "Healthy") for i in mig_ids]
            areq = V.AllocateRequest(container_requests=[V.ContainerAllocateRequest(devices_ids=["amdgpu_xcp_2"])])
            aresp = _call(ch, V.ALLOCATE, areq, V.AllocateResponse).container_responses[0]
            assert dict(aresp.envs) == {"NVIDIA_VISIBLE_DEVICES": "MIG-00000000-0002-4000-8000-00000000b200"}
            assert [c.name for c in aresp.cdi_devices] == ["nvidia.com/gpu=MIG-00000000-0002-4000-8000-00000000b200"]
            assert [d.host_path for d in aresp.devices] == ["/dev/nvidiactl", "/dev/nvidia-uvm", "/dev/nvidia-uvm-tools", "/dev/nvidia1",
                                                            "/dev/nvidia-caps/nvidia-cap104", "/dev/nvidia-caps/nvidia-cap105"]
            req = V.PreferredAllocationRequest(container_requests=[
                V.ContainerPreferredAllocationRequest(available_deviceIDs=mig_ids + ["0000:29:00.0"], allocation_size=2)])
            resp = _call(ch, V.GET_PREFERRED_ALLOCATION, req, V.PreferredAllocationResponse)
            assert set(resp.container_responses[0].deviceIDs) <= set(mig_ids)         # two instances of the same GPU
            stream.cancel()
        with grpc.insecure_channel("unix://" + os.path.join(plug_dir, "amd.com_7g_179gb")) as ch:
            stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                     response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
            assert [(d.ID, d.health) for d in next(stream).devices] == [("0000:29:00.0", "Healthy")]
            stream.cancel()
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            _, err = proc.communicate(timeout=10)
        except subprocess.TimeoutExpired:
            proc.kill()
            _, err = proc.communicate()
        kubelet.server.stop(0)
    assert proc.returncode == 0, err[-2000:]
    if "B200DP_PLUGIND" not in os.environ:                 # (the sanitizer builds link a stand-in for cuda_backend.cu, which logs this)
        assert "NVML enumeration: 4 unit(s), 3 MIG instance(s), probe=helpers (forced by MIG" in err      # the library's log line
