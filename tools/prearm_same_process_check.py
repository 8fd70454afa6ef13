#!/usr/bin/env python3
"""What a pending prearm=1 doorbell wait does to OTHER work this process submits to the same GPU (the reason prearm is
opt-in): arms one pass, then tries torch work on the default and side streams and a second library context, each with a
time limit, then rings the doorbell.  Result of round 2: profiles/r02_prearm_same_process_constraint.txt.

    gpurun -- python tools/prearm_same_process_check.py [n_extra_streams]
"""
import importlib, sys, time, threading, os
sys.path.insert(0, "/root/repo")
import torch
P = importlib.import_module("k8s-device-plugin_b200")
def attempt(label, fn, limit=5.0):
    done = threading.Event()
    def run():
        fn(); done.set()
    th = threading.Thread(target=run, daemon=True); th.start()
    ok = done.wait(limit)
    print(label, "OK" if ok else "HUNG", flush=True)
    return ok
nstreams = int(sys.argv[1]) if len(sys.argv) > 1 else 0
extra = [torch.cuda.Stream(device=0) for _ in range(nstreams)]      # occupy stream-to-queue slots
for s in extra:
    with torch.cuda.stream(s):
        torch.zeros(1, device="cuda:0")
torch.cuda.synchronize()
a = P.Context("cuda:devices=0,bytes=%d,prearm=1" % (64 << 20))
a.probe_health(timed=False, min_gbs=1e-3)          # launches and arms the next pass
print("armed; now other work on the same GPU from this process", flush=True)
ok1 = attempt("torch default stream", lambda: (torch.zeros(4, device="cuda:0").sum().item()))
ok2 = all(attempt("torch side stream %d" % i, lambda s=s: (s.synchronize(), torch.cuda.current_stream(0).wait_stream(s), None)) for i, s in enumerate(extra[:3]))
def other_ctx():
    b = P.Context("cuda:devices=0,bytes=%d,prearm=0" % (64 << 20))
    b.probe_health(timed=False, min_gbs=1e-3)
    b.close()
ok3 = attempt("second library context", other_ctx, 20.0)
print("ringing", flush=True)
r = a.probe_health(timed=False, min_gbs=1e-3)[0]
print("after ring", hex(r.flags), r.healthy, flush=True)
os._exit(0)
