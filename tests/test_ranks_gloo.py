"""The N>1 path of bench.py on CPU: world_size 2 over gloo, one shard of devices per rank, no
data-path collective -- only barrier + max/sum reductions of the timings."""
import json
import os
import socket
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_size_2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(REPO, "tests", "_gloo_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["max"] == 2.0 and out["sum"] == 3.0
    # 20 steps x (2 + 3) devices over the SLOWEST rank's time (rank 1 slept 0.2 s)
    assert round(out["units_total"]) == 20 * 5
    assert out["t"] >= 0.2 and out["t"] >= out["my_dt"]
