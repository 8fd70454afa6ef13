"""Measure the native daemon (b200dp_plugind) the way a kubelet sees it: a grpcio Registration server on
<dir>/kubelet.sock accepts its Register call, a grpcio client opens ListAndWatch on the plugin socket, and
every SIGUSR1 ("heartbeat now") is timed until the next ListAndWatchResponse arrives on the stream.
Used by bench.py and tools/fanout_bench.py; the daemon itself contains no Python."""
import os
import shutil
import signal
import statistics
import subprocess
import tempfile
import time
from concurrent import futures

import grpc

from . import v1beta1 as V

DAEMON = os.path.join(os.path.dirname(os.path.abspath(__file__)), "b200dp_plugind")


def heartbeat_latency_ms(backend_uri: str, iters: int = 100, warmup: int = 5, timeout: float = 60.0):
    """-> dict(median_ms, p99_ms, n_devices, response_bytes) or raises."""
    if not os.path.exists(DAEMON):
        raise FileNotFoundError(DAEMON + " (run __graft_entry__.build())")
    d = tempfile.mkdtemp(prefix="b2d_", dir="/tmp")
    registered = []

    class Kubelet(grpc.GenericRpcHandler):
        def service(self, det):
            if det.method != V.REGISTER:
                return None

            def reg(req, ctx):
                registered.append(V.RegisterRequest.FromString(req))
                return V.Empty().SerializeToString()
            return grpc.unary_unary_rpc_method_handler(reg, lambda b: b, lambda b: b)

    kubelet = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
    kubelet.add_generic_rpc_handlers((Kubelet(),))
    kubelet.add_insecure_port("unix://" + os.path.join(d, "kubelet.sock"))
    kubelet.start()
    proc = subprocess.Popen([DAEMON, "-pulse=0", "-backend=" + backend_uri, "-plugin_dir", d],
                            stderr=subprocess.PIPE, text=True)
    try:
        t0 = time.time()
        while not registered:
            if proc.poll() is not None or time.time() - t0 > timeout:
                raise RuntimeError("daemon did not register: " + (proc.stderr.read() if proc.poll() is not None else "timeout"))
            time.sleep(0.02)
        sock = os.path.join(d, registered[0].endpoint)
        with grpc.insecure_channel("unix://" + sock) as ch:
            stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                     response_deserializer=lambda b: b)(V.Empty())
            first = next(stream)
            lat = []
            for i in range(iters + warmup):
                t0 = time.perf_counter()
                os.kill(proc.pid, signal.SIGUSR1)
                msg = next(stream)
                if i >= warmup:
                    lat.append((time.perf_counter() - t0) * 1e3)
            stream.cancel()
        devs = V.ListAndWatchResponse.FromString(msg).devices
        if not all(x.health == "Healthy" for x in devs):
            raise RuntimeError("daemon reported an unhealthy device")
        lat.sort()
        return {"median_ms": round(statistics.median(lat), 4), "p99_ms": round(lat[min(len(lat) - 1, int(len(lat) * 0.99))], 4),
                "n_devices": len(devs), "response_bytes": len(first)}
    finally:
        if proc.poll() is None:
            proc.send_signal(signal.SIGTERM)
            try:
                proc.wait(timeout=10)
            except subprocess.TimeoutExpired:
                proc.kill()
        kubelet.stop(0)
        shutil.rmtree(d, ignore_errors=True)
