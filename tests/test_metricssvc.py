"""serve the probe's verdicts in the exporter's wire format so an UNMODIFIED
reference plugin consumes them through internal/pkg/exporter/health.go.  The client below restates
getGPUHealth (health.go:42-82) over a real unix socket; the server is server.MetricsServer with a
stub context (CPU test) or the cuda: backend (GPU test)."""
import importlib
import os
from types import SimpleNamespace

import grpc
import pytest

from oracle import plugin as oplug


def reference_get_gpu_health(srv_mod, socket_path):
    """health.go:42-82: stat the socket, List(Empty) with a 5 s timeout, map Device -> health."""
    if not os.path.exists(socket_path):
        return None                                              # err != nil => caller uses the default
    with grpc.insecure_channel("unix://" + socket_path) as ch:
        call = ch.unary_unary(srv_mod.METRICS_LIST, request_serializer=lambda b: b,
                              response_deserializer=srv_mod.GPUStateResponse.FromString)
        resp = call(b"", timeout=5.0)
    return oplug.exporter_states_to_map([(g.Device, g.Health) for g in resp.GPUState])


class StubCtx:
    def __init__(self, verdicts):
        self.verdicts = verdicts

    def enumerate(self):
        return {k: {} for k in self.verdicts}

    def probe_health(self, min_gbs=0.0):
        return [SimpleNamespace(device=i, healthy=v) for i, (k, v) in enumerate(sorted(self.verdicts.items()))]


def test_metrics_server_feeds_the_reference_merge(pkg, short_dir):
    srv_mod = importlib.import_module("k8s-device-plugin_b200.server")
    sock = os.path.join(short_dir, "exp", "amdgpu_device_metrics_exporter_grpc.socket")
    assert reference_get_gpu_health(srv_mod, sock) is None
    verdicts = {"0000:19:00.0": True, "0000:29:00.0": False, "0000:39:00.0": True}
    server = srv_mod.MetricsServer(StubCtx(verdicts), socket_path=sock).start()
    try:
        hmap = reference_get_gpu_health(srv_mod, sock)
        assert hmap == {"0000:19:00.0": "Healthy", "0000:29:00.0": "Unhealthy", "0000:39:00.0": "Healthy"}
        ids = sorted(verdicts) + ["amdgpu_xcp_9"]
        # the reference-side merge (health.go:86-106) and ours agree on what kubelet will see
        want = oplug.merge_health(ids, "Healthy", hmap)
        assert pkg.exporter.PopulatePerGPUDHealth(ids, "Healthy", hmap) == want == ["Healthy", "Unhealthy", "Healthy", "Healthy"]
        with grpc.insecure_channel("unix://" + sock) as ch:
            get = ch.unary_unary(srv_mod.METRICS_GET, request_serializer=lambda m: m.SerializeToString(),
                                 response_deserializer=srv_mod.GPUStateResponse.FromString)
            one = get(srv_mod.GPUGetRequest(ID=["1"]), timeout=5)
            assert [(g.ID, g.Device, g.Health) for g in one.GPUState] == [("1", "0000:29:00.0", "unhealthy")]
    finally:
        server.stop()
    assert not os.path.exists(sock)


@pytest.mark.gpu
def test_metrics_server_on_cuda_backend(pkg, short_dir):
    srv_mod = importlib.import_module("k8s-device-plugin_b200.server")
    sock = os.path.join(short_dir, "amdgpu_device_metrics_exporter_grpc.socket")
    with pkg.Context("cuda:bytes=%d" % (64 << 20)) as ctx:
        ids = sorted(ctx.enumerate())
        server = srv_mod.MetricsServer(ctx, socket_path=sock, min_gbs=1e-3).start()
        try:
            assert reference_get_gpu_health(srv_mod, sock) == {i: "Healthy" for i in ids}
            ctx.probe_inject_fault(0, 1000, 0x8000)
            assert reference_get_gpu_health(srv_mod, sock)[ids[0]] == "Unhealthy"
            assert reference_get_gpu_health(srv_mod, sock)[ids[0]] == "Healthy"
        finally:
            server.stop()


@pytest.mark.gpu
def test_native_daemon_serves_the_exporter_contract(pkg, short_dir):
    """The native daemon's -exporter_socket: the reference's own exporter client (restated above) reads one
    verdict per kubelet device id, produced by the HBM pass."""
    import signal
    import subprocess
    import time
    from concurrent import futures
    srv_mod = importlib.import_module("k8s-device-plugin_b200.server")
    dp = importlib.import_module("k8s-device-plugin_b200.daemon_probe")
    V = pkg.v1beta1
    if not os.path.exists(dp.DAEMON):
        import __graft_entry__
        __graft_entry__.build()

    class Kubelet(grpc.GenericRpcHandler):
        def service(self, det):
            if det.method != V.REGISTER:
                return None
            return grpc.unary_unary_rpc_method_handler(lambda req, ctx: V.Empty().SerializeToString(), lambda b: b, lambda b: b)
    kubelet = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
    kubelet.add_generic_rpc_handlers((Kubelet(),))
    kubelet.add_insecure_port("unix://" + os.path.join(short_dir, "kubelet.sock"))
    kubelet.start()
    sock = os.path.join(short_dir, "amdgpu_device_metrics_exporter_grpc.socket")
    proc = subprocess.Popen([dp.DAEMON, "-backend=cuda:bytes=%d,min_gbs=0.001" % (64 << 20), "-plugin_dir", short_dir,
                             "-exporter_socket", sock], stderr=subprocess.PIPE, text=True)
    try:
        t0 = time.time()
        while not os.path.exists(sock) and time.time() - t0 < 60 and proc.poll() is None:
            time.sleep(0.05)
        with pkg.Context("cuda:bytes=%d" % (1 << 20)) as ctx:
            ids = sorted(ctx.enumerate())
        for _ in range(3):
            assert reference_get_gpu_health(srv_mod, sock) == {i: "Healthy" for i in ids}
    finally:
        proc.send_signal(signal.SIGTERM)
        _, err = proc.communicate(timeout=20)
        kubelet.stop(0)
    assert proc.returncode == 0, err[-2000:]
