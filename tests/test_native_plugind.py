"""The native daemon (k8s-device-plugin_b200/b200dp_plugind: C++ HTTP/2 + HPACK + gRPC from
csrc/host/h2grpc.hpp, no Python, no gRPC library) against grpcio playing the kubelet: interop in both
directions -- grpcio's client (Huffman-coded, dynamically indexed HPACK, PINGs, flow control) drives the
native server; the native client registers with a grpcio Registration server -- and the same RPC
answers as the reference logic (oracle) on the reference's CPX capture.  CPU only (kfd: backend)."""
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
        reg = kubelet.requests.get(timeout=10)                 # the native gRPC client -> grpcio server
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
            assert 0.5 < time.time() - t0 < 4.0
            for m in (second, third):
                assert [d.ID for d in m.devices] == ids and all(d.health == "Healthy" for d in m.devices)
            t0 = time.time()
            proc.send_signal(signal.SIGUSR1)                    # operator's "heartbeat now"; the ticker just fired
            assert [d.ID for d in next(stream).devices] == ids and time.time() - t0 < 0.5

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
        reg = kubelet.requests.get(timeout=10)
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
