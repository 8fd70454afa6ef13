#!/usr/bin/env python3
"""Does a pending prearm=1 doorbell wait slow down ANOTHER process on the same GPU (a tenant)?  The parent keeps a pass
armed (or, with `idle`, merely holds its context) while a child process times a small kernel + .item() every 10 ms for 5 s.
Result of round 2 (armed): 493 ops, median 0.040 ms, p99 0.128 ms -- unaffected.

    gpurun -- python tools/prearm_other_process_check.py armed|idle
"""
import importlib, sys, time, subprocess, os
sys.path.insert(0, "/root/repo")
if len(sys.argv) > 1 and sys.argv[1] == "tenant":
    import torch
    x = torch.ones(1 << 20, device="cuda:0")
    torch.cuda.synchronize()
    print("tenant ready", flush=True)
    lat = []
    t_end = time.time() + 5.0
    while time.time() < t_end:
        t0 = time.perf_counter()
        (x * 2).sum().item()
        lat.append((time.perf_counter() - t0) * 1e3)
        time.sleep(0.01)
    lat.sort()
    print("tenant ops=%d median_ms=%.3f p99_ms=%.3f max_ms=%.3f" % (len(lat), lat[len(lat)//2], lat[int(0.99*(len(lat)-1))], lat[-1]), flush=True)
    sys.exit(0)
P = importlib.import_module("k8s-device-plugin_b200")
mode = sys.argv[1]
a = P.Context("cuda:devices=0,bytes=%d,prearm=%s" % (64 << 20, "1" if mode == "armed" else "0"))
a.probe_health(timed=False, min_gbs=1e-3)          # armed mode: the next pass now waits behind its doorbell
ten = subprocess.Popen([sys.executable, __file__, "tenant"], stdout=subprocess.PIPE, text=True)
print(mode, ten.stdout.readline().strip(), flush=True)
out = ten.communicate(timeout=60)[0]
print(mode, out.strip(), flush=True)
r = a.probe_health(timed=False, min_gbs=1e-3)[0]
print(mode, "after", hex(r.flags), r.healthy, flush=True)
os._exit(0)
