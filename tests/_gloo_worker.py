"""Worker for tests/test_ranks_gloo.py: the N>1 launch shape of bench.py on CPU (gloo, kfd: backend)."""
import importlib
import json
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
pkg = importlib.import_module("k8s-device-plugin_b200")
ranks = importlib.import_module("k8s-device-plugin_b200.ranks")

rg = ranks.RankGroup(backend="gloo")
root = tempfile.mkdtemp(prefix="b2dp_gloo_%d_" % rg.rank)
n_dev = 2 + rg.rank                       # uneven shards on purpose
pkg.synth.write_b200_tree(root, n_gpus=n_dev)
ctx = pkg.Context("kfd:" + root)
rg.barrier()
steps = 20
t0 = time.perf_counter()
units = 0
for _ in range(steps):
    wire, st = ctx.list_and_watch("gpu", pkg._native.LW_HEARTBEAT | pkg._native.LW_NO_PROBE)
    units += st.n_devices
if rg.rank == 1:
    time.sleep(0.2)                       # make the slow rank identifiable
dt = time.perf_counter() - t0
rg.barrier()
value, t = ranks.aggregate_throughput(rg, units, dt)
mx, sm = rg.max(rg.rank + 1), rg.sum(rg.rank + 1)
if rg.rank == 0:
    print(json.dumps({"value": value, "t": t, "my_dt": dt, "units_total": rg.world and value * t, "max": mx, "sum": sm,
                      "world": rg.world}))
ctx.close()
rg.close()
