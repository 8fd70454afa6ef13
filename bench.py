#!/usr/bin/env python3
"""bench.py -- the hot path of BASELINE.json on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, 1 rank/GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

metric  : health_probe_hbm_gbs -- bytes of HBM streamed + verified per second by the per-GPU health
          probe (BASELINE.json "per-GPU health-probe HBM GB/s vs peak"), whole job = sum over GPUs;
          `ms_per_step`/`cycle_ms` carry the second half of the metric ("ListAndWatch cycle ms").
step    : one heartbeat's GPU work on every GPU of the job: 1 GiB src -> 1 GiB dst per GPU, verified
          word by word (algorithmic bytes 2*S = 2 GiB per GPU per step).  Inputs (1 GiB) >> 126 MB L2.
value   : K x b2dp_probe_health() (C ABI; buffers resident in HBM), wall clock bracketed by
          barrier + synchronize, max over ranks.
e2e     : K x b2dp_list_and_watch(HEARTBEAT): enumerate -> node health -> probe fan-out -> health
          merge -> serialized v1beta1.ListAndWatchResponse in a host buffer -- the call a kubelet-facing
          host makes.  The path has no bulk host inputs: per step the host sends kernel arguments and
          receives the 48-byte result block per device through pinned mapped memory plus the response.
roofline: achieved = 2*S / mean CUDA-event time of the probe kernel over the timed steps; peak =
          MEASURED_PEAKS.json hbm_gbs (else the 6650 GB/s fallback of B200_PROFILING.md).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
PKG = "k8s-device-plugin_b200"
S_BYTES = 1 << 30
METRIC, UNIT = "health_probe_hbm_gbs", "GB/s"
WORKLOAD = "1xB200 enumerate + ListAndWatch with sm_100a HBM health-probe kernel (BASELINE.json configs[1]), per GPU"

SAMPLER = r"""
import sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
bad = {getattr(nv, n): n for n in dir(nv) if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason")}
mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
out = open(sys.argv[2], "w", buffering=1)
out.write("max,%d\n" % mx)
while True:
    try:
        sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        out.write("%d,%d,%d\n" % (time.monotonic_ns(), sm, r))
    except Exception as e:
        out.write("err,%s\n" % e)
    time.sleep(0.004)
"""

# NVML clocks-event-reason bits (nvml.h)
REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
           0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
           0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}


class ClockSampler:
    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="b2dp_clk_", suffix=".csv")
        try:
            self.p = subprocess.Popen([sys.executable, "-c", SAMPLER, str(gpu_index), self.path],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        time.sleep(0.3)

    def stop(self, windows):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, bits = [], 0
        try:
            for line in open(self.path):
                a = line.strip().split(",")
                if a[0] == "max":
                    out["sm_max_mhz"] = int(a[1])
                elif a[0].isdigit():
                    t = int(a[0])
                    if any(lo <= t <= hi for lo, hi in windows):
                        sm.append(int(a[1]))
                        bits |= int(a[2])
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["samples"] = len(sm)
        out["reasons"] = [n for b, n in REASONS.items() if bits & b and n != "gpu_idle"]
        return out


def peak_hbm():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_per_launch():
    try:
        return json.load(open(os.path.join(REPO, "profiles", "traffic.json")))["hbm_probe_tma_bytes_per_launch"]
    except Exception:
        return None


def cpu_probe_baseline(threads, budget_s=6.0, sample_bytes=256 << 20):
    """The oracle's C restatement of the probe pass streamed through host DRAM (bounded sample)."""
    import numpy as np
    from oracle import cbind
    import ctypes as C
    cbind.build()
    lib = cbind.probe_lib()
    n_words = sample_bytes // 4
    src = np.empty(n_words, dtype=np.uint32)
    dst = np.empty(n_words, dtype=np.uint32)
    seed = 0x5EED0000
    lib.oracle_fill(src.ctypes.data, n_words, seed, threads)
    out = (C.c_uint64 * 3)()
    lib.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, 0, threads, out)   # warm-up / page-in
    passes, t0 = 0, time.perf_counter()
    while True:
        lib.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, 0, threads, out)
        passes += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or passes >= 400:
            break
    assert out[1] == 0
    return 2.0 * sample_bytes * passes / dt / 1e9, passes, dt


def kfd_walk_baseline(n_devices, reps=200):
    """Reference-shaped CPU cycle (C port of the Go walk) on a synthetic N-device kfd tree in /dev/shm."""
    from oracle import cbind
    cbind.build()
    k = cbind.kfd_lib()
    pkg = importlib.import_module(PKG)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    root = tempfile.mkdtemp(prefix="b2dp_bench_kfd_", dir=base)
    pkg.synth.write_b200_tree(root, n_gpus=n_devices)
    r = root.encode()
    for _ in range(3):
        k.kfdwalk_cycle(r, 1)
    t0 = time.perf_counter()
    for _ in range(reps):
        k.kfdwalk_cycle(r, 1)
    start_ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    for _ in range(reps):
        k.kfdwalk_cycle(r, 0)
    beat_ms = (time.perf_counter() - t0) / reps * 1e3
    import shutil
    shutil.rmtree(root, ignore_errors=True)
    return start_ms, beat_ms


def run_reference(args, rank, world):
    """The reference's CPU path on the box's host cores: per step, the reference-shaped
    ListAndWatch cycle (kfd walk: enumerate x2 + text health check, C port of the Go code) followed
    by the CPU restatement of the probe over a bounded sample per device."""
    if rank != 0:
        return
    import numpy as np
    import ctypes as C
    from oracle import cbind
    cbind.build()
    lib, k = cbind.probe_lib(), cbind.kfd_lib()
    pkg = importlib.import_module(PKG)
    threads = os.cpu_count() or 1
    n = args.gpus
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    root = tempfile.mkdtemp(prefix="b2dp_ref_kfd_", dir=base)
    pkg.synth.write_b200_tree(root, n_gpus=n)
    r = root.encode()
    # bounded sample: keep K steps within ~60 s
    sample = 256 << 20
    n_words = sample // 4
    src = np.empty(n_words, dtype=np.uint32)
    dst = np.empty(n_words, dtype=np.uint32)
    seed = 0x5EED0000
    lib.oracle_fill(src.ctypes.data, n_words, seed, threads)
    out = (C.c_uint64 * 3)()
    t0 = time.perf_counter()
    lib.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, 0, threads, out)
    lib.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, 0, threads, out)
    per_pass = (time.perf_counter() - t0) / 2
    while sample > (16 << 20) and per_pass * n * (args.steps + args.warmup) > 60.0:
        sample //= 2
        per_pass /= 2
    n_words = sample // 4

    def step():
        k.kfdwalk_cycle(r, 1)
        for _ in range(n):
            lib.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, 0, threads, out)
    for _ in range(max(3, args.warmup)):
        step()
    walk = 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        w0 = time.perf_counter()
        k.kfdwalk_cycle(r, 1)
        walk += time.perf_counter() - w0
        for _ in range(n):
            lib.oracle_probe_pass(src.ctypes.data, dst.ctypes.data, n_words, seed, 0, threads, out)
    dt = time.perf_counter() - t0
    import shutil
    shutil.rmtree(root, ignore_errors=True)
    value = n * 2.0 * sample * args.steps / dt / 1e9
    sample_desc = "%d MiB of the 1 GiB per-device probe buffer per device per step, %d device(s), host DRAM" % (sample >> 20, n)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": n,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(dt / args.steps * 1e3, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_devices": n, "note": "CPU arm: C port of the reference's kfd walk (no Go "
                   "toolchain here) + CPU restatement of the probe on a bounded sample"},
        "cycle_ms_kfd_walk": round(walk / args.steps * 1e3, 4),
        "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_desc},
        "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--variant", type=int, default=0, help="0 = TMA-staged kernel (default), 1 = register path")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: libraries (NCCL prints its version banner there) are
    # diverted to stderr for the whole run and the line is written to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    warmup = max(3, args.warmup)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the probe has no CPU fallback")
    torch.cuda.set_device(local_rank)
    pkg = importlib.import_module(PKG)      # raises if libb200dp.so is missing
    N = pkg._native
    ranks = importlib.import_module(PKG + ".ranks")
    rg = ranks.RankGroup(backend="nccl")   # barrier + timing reductions only; no data-path collective
    ctx = pkg.Context("cuda:devices=%d,bytes=%d" % (local_rank, S_BYTES))
    barrier, max_over_ranks, sum_over_ranks = rg.barrier, rg.max, rg.sum

    sampler = ClockSampler(local_rank) if rank == 0 else None
    windows = []

    def warm(fn):
        """W untimed warm-up steps, and never less than 100 ms of them: the first NCCL barrier
        (communicator init) idles the GPU for seconds, so clocks and the driver must be warm
        again before a timed region starts."""
        t_w, i = time.perf_counter(), 0
        while i < warmup or time.perf_counter() - t_w < 0.1:
            fn()
            i += 1

    barrier()   # pays the NCCL communicator init before any warm-up
    # ---- value leg: the probe through the C ABI, buffers resident ------------------------------
    warm(lambda: ctx.probe_health(variant=args.variant))
    barrier()
    kernel_ms, unhealthy = [], 0
    w0 = time.monotonic_ns()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = ctx.probe_health(variant=args.variant)
        kernel_ms.append(res[0].ms_event)
        unhealthy += sum(0 if r.healthy else 1 for r in res)
    torch.cuda.synchronize()
    t_value = time.perf_counter() - t0
    windows.append((w0, time.monotonic_ns()))
    barrier()
    t_value = max_over_ranks(t_value)

    # ---- context for the roofline: the driver's own device-to-device copy of the same size, in this run, on this
    # GPU, timed the same way (CUDA events, after warm-up) -- what a plain copy achieves on this part today -------
    drv_copy_gbs = None
    try:
        a_t = torch.empty(S_BYTES, dtype=torch.uint8, device="cuda")
        b_t = torch.empty(S_BYTES, dtype=torch.uint8, device="cuda")
        a_t.zero_()
        for _ in range(5):
            b_t.copy_(a_t)
        torch.cuda.synchronize()
        best = None
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b_t.copy_(a_t)
            e1.record()
            e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None or ms < best else best
        drv_copy_gbs = 2.0 * S_BYTES / (best * 1e-3) / 1e9
        del a_t, b_t
        torch.cuda.empty_cache()
    except Exception as e:      # noqa: BLE001
        print("driver copy reference skipped: %r" % (e,), file=sys.stderr)
    barrier()

    # ---- e2e leg: the kubelet-facing call ----------------------------------------------------------
    warm(lambda: ctx.list_and_watch("gpu", N.LW_HEARTBEAT, variant=args.variant))
    barrier()
    enum_ms = enc_ms = 0.0
    w0 = time.monotonic_ns()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wire, st = ctx.list_and_watch("gpu", N.LW_HEARTBEAT, variant=args.variant)
        unhealthy += st.n_unhealthy
        enum_ms += st.ms_enumerate
        enc_ms += st.ms_encode
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    windows.append((w0, time.monotonic_ns()))
    barrier()
    t_e2e = max_over_ranks(t_e2e)
    unhealthy = sum_over_ranks(float(unhealthy))
    kernel_ms_mean = sum(kernel_ms) / len(kernel_ms)
    kernel_ms_max_rank = max_over_ranks(kernel_ms_mean)
    msg = pkg.v1beta1.ListAndWatchResponse.FromString(wire)
    assert len(msg.devices) == 1 and msg.devices[0].health == "Healthy", msg

    # ---- stream start: the initial ListAndWatch list (enumerate + homogeneity + encode; the reference does
    # two GetAMDGPUs() walks here, plugin.go:231,237) -- reported next to reference_cpu_path ------
    starts = []
    for i in range(60):
        t0 = time.perf_counter()
        ctx.list_and_watch("gpu", N.LW_INITIAL)
        if i >= 10:
            starts.append((time.perf_counter() - t0) * 1e3)
    starts.sort()
    stream_start_ms = max_over_ranks(starts[len(starts) // 2])

    # ---- the kubelet's view: heartbeat tick -> ListAndWatchResponse received over the plugin's unix
    # socket (grpcio server + client stream in this process); informative, not the timed `e2e` ------
    grpc_ms = None
    try:
        import grpc
        import statistics
        srv_mod = importlib.import_module(PKG + ".server")
        V = pkg.v1beta1
        d = tempfile.mkdtemp(prefix="b2b_", dir="/tmp")
        plugin = pkg.plugin.AMDGPUPlugin(ctx, "gpu")
        plugin.Start = lambda: None
        server = srv_mod.PluginServer(plugin, plugin_dir=d).start()
        with grpc.insecure_channel("unix://" + server.socket_path) as ch:
            stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                     response_deserializer=lambda b: b)(V.Empty())
            next(stream)
            lat = []
            for i in range(105):
                t0 = time.perf_counter()
                plugin.Heartbeat.put(True)
                next(stream)
                if i >= 5:
                    lat.append((time.perf_counter() - t0) * 1e3)
            stream.cancel()
        server.stop()
        grpc_ms = max_over_ranks(statistics.median(lat))
    except Exception as e:      # noqa: BLE001
        print("grpc leg skipped: %r" % (e,), file=sys.stderr)
        grpc_ms = max_over_ranks(-1.0)

    # ---- the same, with the native daemon (b200dp_plugind: C++ gRPC host, own process, own probe ring on this
    # GPU): SIGUSR1 "heartbeat now" -> probe -> ListAndWatchResponse received by a grpcio client; rank 0 only --
    native_ms = None
    if rank == 0:
        try:
            dp = importlib.import_module(PKG + ".daemon_probe")
            native_ms = dp.heartbeat_latency_ms("cuda:devices=%d,bytes=%d" % (local_rank, S_BYTES), iters=100)
        except Exception as e:      # noqa: BLE001
            print("native daemon leg skipped: %r" % (e,), file=sys.stderr)
    barrier()

    clocks = sampler.stop(windows) if sampler else None
    if rank != 0:
        ctx.close()
        rg.close()
        return

    bytes_per_step_per_gpu = 2.0 * S_BYTES
    n = world
    value = n * bytes_per_step_per_gpu * args.steps / t_value / 1e9
    e2e_value = n * bytes_per_step_per_gpu * args.steps / t_e2e / 1e9
    peak, peak_src = peak_hbm()
    achieved = bytes_per_step_per_gpu / (kernel_ms_mean * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": n, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(t_value / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "probe_bytes_per_buffer": S_BYTES, "algorithmic_bytes_per_gpu_step": int(bytes_per_step_per_gpu),
                   "kernel": "hbm_probe_tma<4,1024,3> grid=2xSMs" if args.variant == 0 else "hbm_probe_r128<512,2> grid=2xSMs",
                   "l2": "inputs (1 GiB per buffer) larger than the 126 MB L2; buffers ping-pong every step",
                   "parallelism": "1 process per GPU, no data-path collective"},
        "cycle_ms": round(t_e2e / args.steps * 1e3, 4),
        "e2e": {"value": round(e2e_value, 1), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 48 * 1 + len(wire),
                "cycle_ms": round(t_e2e / args.steps * 1e3, 4), "ms_enumerate": round(enum_ms / args.steps, 5),
                "ms_encode": round(enc_ms / args.steps, 5), "response_bytes": len(wire),
                "stream_start_ms": round(stream_start_ms, 4),
                "heartbeat_to_kubelet_grpc_ms": None if grpc_ms is None or grpc_ms < 0 else round(grpc_ms, 4),
                "heartbeat_to_kubelet_native_daemon_ms": native_ms,
                "note": "no bulk host buffers on this path: kernel arguments in, 48-byte result block (pinned mapped) + serialized response out; the value leg brackets every kernel with CUDA events (the roofline's clock), the kubelet-facing call completes on the published result block alone"},
        "gpu_launches": args.steps * n,
        "unhealthy_verdicts": int(unhealthy),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": traffic_per_launch(), "peak_source": peak_src,
                     "kernel_ms_mean": round(kernel_ms_mean, 5), "kernel_ms_mean_slowest_rank": round(kernel_ms_max_rank, 5),
                     "driver_d2d_copy_gbs_same_run": None if drv_copy_gbs is None else round(drv_copy_gbs, 1),
                     "note": "driver_d2d_copy = torch copy_ of the same 1 GiB (best of 20, CUDA events): the mixed read+write ceiling of this part; the probe verifies and re-keys every word at that rate"},
        "clocks": clocks,
    }
    # the reference's own sysfs/kfd CPU path on this box's host cores, in the same run, at every N
    # (C port of the Go walk on a synthetic N-device kfd tree in /dev/shm; single thread like the
    # reference's single goroutine): ~0.2 s of CPU work
    start_ms, beat_ms = kfd_walk_baseline(n)
    line["reference_cpu_path"] = {"kfd_walk_stream_start_ms": round(start_ms, 4), "kfd_walk_heartbeat_ms": round(beat_ms, 4),
                                  "n_devices": n, "threads": 1, "host_cores": os.cpu_count(), "kind": "port"}
    if n == 1:
        threads = os.cpu_count() or 1
        cpu_gbs, passes, dt = cpu_probe_baseline(threads)
        line["cpu_baseline"] = {"value": round(cpu_gbs, 2), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "256 MiB of the 1 GiB probe buffer, %d passes in %.1f s, host DRAM" % (passes, dt),
                                "kfd_walk_stream_start_ms": round(start_ms, 4), "kfd_walk_heartbeat_ms": round(beat_ms, 4)}
    print(json.dumps(line), flush=True)
    ctx.close()
    rg.close()


if __name__ == "__main__":
    main()
