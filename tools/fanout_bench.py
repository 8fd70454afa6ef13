#!/usr/bin/env python3
"""Single-process N-GPU fan-out: the deployment shape of the plugin (one daemon per node, the
library drives every GPU from one process).  Measures BASELINE.json configs[2]/[3]:

  - full ListAndWatch heartbeat cycle (enumerate -> node health -> probe x N concurrently ->
    merge -> serialized response) wall clock, median of K after W warm-ups
  - per-GPU probe GB/s (CUDA events) vs the measured HBM peak
  - the NVLink P2P matrix (GB/s per directed pair) and the link classes it yields
  - GetPreferredAllocation latency on the measured topology
  - heartbeat tick -> response received by a (fake) kubelet over a real unix socket (grpcio)

    gpurun --gpus 8 -- python tools/fanout_bench.py --sweep --out gpurun_out/fanout_sweep.json
"""
import argparse
import importlib
import json
import os
import statistics
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run(pkg, n_gpus, steps, warmup, p2p=True, grpc_leg=True):
    N = pkg._native
    uri = "cuda:bytes=%d" % (1 << 30)
    if n_gpus:
        uri += ",devices=" + "+".join(str(i) for i in range(n_gpus))
    t0 = time.perf_counter()
    ctx = pkg.Context(uri)
    open_s = time.perf_counter() - t0
    devs = ctx.enumerate()
    n = len(devs)
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    ctx.list_and_watch("gpu", N.LW_INITIAL)
    for _ in range(warmup):
        ctx.list_and_watch("gpu", N.LW_HEARTBEAT)
    cyc, probe_ms, per_gpu = [], [], [[] for _ in range(n)]
    for _ in range(steps):
        t0 = time.perf_counter()
        wire, st = ctx.list_and_watch("gpu", N.LW_HEARTBEAT)
        cyc.append((time.perf_counter() - t0) * 1e3)
        probe_ms.append(st.ms_probe)
        assert st.n_devices == n and st.n_unhealthy == 0, st
    for _ in range(steps):
        for r in ctx.probe_health():
            per_gpu[r.device].append(r.gbs)
    out = {
        "n_gpus": n, "devices": list(devs), "open_s": round(open_s, 3), "steps": steps,
        "cycle_ms_median": round(statistics.median(cyc), 4), "cycle_ms_p99": round(sorted(cyc)[int(0.99 * (len(cyc) - 1))], 4),
        "cycle_ms_max": round(max(cyc), 4), "probe_fanout_ms_median": round(statistics.median(probe_ms), 4),
        "per_gpu_gbs_median": [round(statistics.median(g), 1) for g in per_gpu],
        "per_gpu_frac_of_peak": [round(statistics.median(g) / peak, 4) for g in per_gpu], "hbm_peak_gbs": peak,
        "aggregate_gbs_in_cycle": round(n * 2 * (1 << 30) / (statistics.median(cyc) * 1e-3) / 1e9, 1),
        "response_bytes": len(wire),
    }
    if n > 1 and p2p:
        t0 = time.perf_counter()
        gbs, lt, mm = ctx.p2p_matrix()
        out["p2p_matrix_s"] = round(time.perf_counter() - t0, 3)
        out["p2p_gbs"] = [[round(float(x), 1) for x in row] for row in gbs]
        out["p2p_link_type"] = [[int(x) for x in row] for row in lt]
        out["p2p_mismatches"] = int(mm.sum())
        off = [float(gbs[i, j]) for i in range(n) for j in range(n) if i != j]
        out["p2p_gbs_min"], out["p2p_gbs_median"] = round(min(off), 1), round(statistics.median(off), 1)
        out["p2p_frac_of_770_median"] = round(statistics.median(off) / 770.0, 4)
        t0 = time.perf_counter()
        gb2, lt2, mm2 = ctx.p2p_matrix(bidir=True)
        out["p2p_bidir_matrix_s"] = round(time.perf_counter() - t0, 3)
        off2 = [float(gb2[i, j]) for i in range(n) for j in range(n) if i != j]
        out["p2p_bidir_gbs_min"], out["p2p_bidir_gbs_median"] = round(min(off2), 1), round(statistics.median(off2), 1)
        out["p2p_bidir_mismatches"] = int(mm2.sum())
        assert ctx.start() == 0
        ids = sorted(devs)
        lat = []
        for size in range(1, n):
            ctx.preferred_allocation(ids, [], size)
            t0 = time.perf_counter()
            got = ctx.preferred_allocation(ids, [], size)
            lat.append((time.perf_counter() - t0) * 1e3)
            assert len(got) == size
        out["preferred_allocation_ms_by_size"] = [round(x, 4) for x in lat]
        out["labels_p2p_link"] = ctx.generate_labels(["p2p-link"])
    if grpc_leg:
        # kubelet <-> plugin over a real unix socket (grpcio): heartbeat tick -> probe on every GPU ->
        # ListAndWatchResponse received by the (fake) kubelet's stream
        try:
            import grpc
            srv_mod = importlib.import_module("k8s-device-plugin_b200.server")
            V = pkg.v1beta1
            d = tempfile.mkdtemp(prefix="b2f_", dir="/tmp")
            plugin = pkg.plugin.AMDGPUPlugin(ctx, "gpu")
            plugin.Start = lambda: None            # the allocator is not under test here
            server = srv_mod.PluginServer(plugin, plugin_dir=d).start()
            with grpc.insecure_channel("unix://" + server.socket_path) as ch:
                stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                         response_deserializer=lambda b: b)(V.Empty())
                next(stream)
                lat = []
                for i in range(steps + 5):
                    t0 = time.perf_counter()
                    plugin.Heartbeat.put(True)
                    wire2 = next(stream)
                    if i >= 5:
                        lat.append((time.perf_counter() - t0) * 1e3)
                stream.cancel()
            server.stop()
            assert len(V.ListAndWatchResponse.FromString(wire2).devices) == n
            out["grpc_heartbeat_to_kubelet_ms_median"] = round(statistics.median(lat), 4)
            out["grpc_heartbeat_to_kubelet_ms_p99"] = round(sorted(lat)[int(0.99 * (len(lat) - 1))], 4)
        except Exception as e:      # noqa: BLE001
            out["grpc_error"] = repr(e)
    ctx.close()
    if grpc_leg:
        # the native daemon (C++ gRPC host in its own process) over the same GPUs: SIGUSR1 -> probe fan-out ->
        # response received by a grpcio client playing the kubelet
        try:
            dp = importlib.import_module("k8s-device-plugin_b200.daemon_probe")
            r = dp.heartbeat_latency_ms(uri, iters=steps)
            out["native_daemon_heartbeat_to_kubelet_ms_median"] = r["median_ms"]
            out["native_daemon_heartbeat_to_kubelet_ms_p99"] = r["p99_ms"]
        except Exception as e:      # noqa: BLE001
            out["native_daemon_error"] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--sweep", action="store_true", help="run at 1, 2, 4, ... up to all visible GPUs")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-p2p", action="store_true")
    args = ap.parse_args()
    pkg = importlib.import_module("k8s-device-plugin_b200")
    if args.sweep:
        import torch
        total = torch.cuda.device_count()
        counts = [c for c in (1, 2, 4, 8) if c <= total]
        if total not in counts:
            counts.append(total)
        res = [run(pkg, c, args.steps, args.warmup, p2p=(not args.no_p2p and c == total)) for c in counts]
        summary = [{k: r.get(k) for k in ("n_gpus", "cycle_ms_median", "cycle_ms_p99", "per_gpu_frac_of_peak",
                                           "aggregate_gbs_in_cycle", "grpc_heartbeat_to_kubelet_ms_median",
                                           "native_daemon_heartbeat_to_kubelet_ms_median")} for r in res]
        out = {"summary": summary, "runs": res}
    else:
        out = run(pkg, args.gpus, args.steps, args.warmup, p2p=not args.no_p2p)
    s = json.dumps(out)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")


if __name__ == "__main__":
    main()
