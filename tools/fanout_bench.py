#!/usr/bin/env python3
"""Single-process N-GPU fan-out: the deployment shape of the plugin (one daemon per node, one
worker thread + stream per GPU inside libb200dp).  Measures BASELINE.json configs[2]/[3]:

  - full ListAndWatch heartbeat cycle (enumerate -> node health -> probe x N concurrently ->
    merge -> serialized response) wall clock, median of K after W warm-ups
  - per-GPU probe GB/s (CUDA events) vs the measured HBM peak
  - the NVLink P2P matrix (GB/s per directed pair) and the link classes it yields
  - GetPreferredAllocation latency on the measured topology

    gpurun --gpus 8 -- python tools/fanout_bench.py --out gpurun_out/fanout_8.json
"""
import argparse
import importlib
import json
import os
import statistics
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--gpus", type=int, default=0, help="0 = all visible")
    ap.add_argument("--out", default="")
    ap.add_argument("--no-p2p", action="store_true")
    args = ap.parse_args()
    pkg = importlib.import_module("k8s-device-plugin_b200")
    N = pkg._native
    uri = "cuda:bytes=%d" % (1 << 30)
    if args.gpus:
        uri += ",devices=" + "+".join(str(i) for i in range(args.gpus))
    t0 = time.perf_counter()
    ctx = pkg.Context(uri)
    open_s = time.perf_counter() - t0
    devs = ctx.enumerate()
    n = len(devs)
    try:
        peak = float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    for _ in range(args.warmup):
        ctx.list_and_watch("gpu", N.LW_HEARTBEAT)
    cyc, probe_ms, per_gpu = [], [], [[] for _ in range(n)]
    for _ in range(args.steps):
        t0 = time.perf_counter()
        wire, st = ctx.list_and_watch("gpu", N.LW_HEARTBEAT)
        cyc.append((time.perf_counter() - t0) * 1e3)
        probe_ms.append(st.ms_probe)
        assert st.n_devices == n and st.n_unhealthy == 0, st
    for _ in range(args.steps):
        for r in ctx.probe_health():
            per_gpu[r.device].append(r.gbs)
    out = {
        "n_gpus": n, "devices": list(devs), "open_s": round(open_s, 3), "steps": args.steps,
        "cycle_ms_median": round(statistics.median(cyc), 4), "cycle_ms_p99": round(sorted(cyc)[int(0.99 * (len(cyc) - 1))], 4),
        "cycle_ms_max": round(max(cyc), 4), "probe_fanout_ms_median": round(statistics.median(probe_ms), 4),
        "per_gpu_gbs_median": [round(statistics.median(g), 1) for g in per_gpu],
        "per_gpu_frac_of_peak": [round(statistics.median(g) / peak, 4) for g in per_gpu], "hbm_peak_gbs": peak,
        "aggregate_gbs_in_cycle": round(n * 2 * (1 << 30) / (statistics.median(cyc) * 1e-3) / 1e9, 1),
        "response_bytes": len(wire),
    }
    if n > 1 and not args.no_p2p:
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
            t0 = time.perf_counter()
            got = ctx.preferred_allocation(ids, [], size)
            lat.append((time.perf_counter() - t0) * 1e3)
            assert len(got) == size
        out["preferred_allocation_ms_by_size"] = [round(x, 4) for x in lat]
    s = json.dumps(out)
    print(s)
    if args.out:
        with open(args.out, "w") as f:
            f.write(s + "\n")
    ctx.close()


if __name__ == "__main__":
    main()
