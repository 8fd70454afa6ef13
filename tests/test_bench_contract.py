"""bench.py's reference arm runs on CPU, so its JSON contract is checked here on every round: one line on
stdout, the keys the driver reads, and the product arm refusing to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "health_probe_hbm_gbs" and d["unit"] == "GB/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["n_gpus"] == 1
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] > 0
    e = d["e2e"]
    assert (e["value"], e["unit"], e["h2d_bytes_per_step"], e["d2h_bytes_per_step"]) == (d["value"], d["unit"], 0, 0)


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_product_never_imports_links_or_executes_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference legs may touch it.  The package, its native sources, the tools and the headers must not."""
    import re
    pat = re.compile(r"^\s*(from\s+oracle\b|import\s+oracle\b)|#include\s+[\"<][^\">]*oracle/|dlopen\([^)]*oracle|oracle/_build", re.M)
    offenders = []
    for top in ("k8s-device-plugin_b200", "tools", "include"):
        for dirpath, dirnames, files in os.walk(os.path.join(ROOT, top)):
            dirnames[:] = [d for d in dirnames if d not in ("build", "__pycache__")]
            for f in files:
                if not f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h", ".inc")):
                    continue
                p = os.path.join(dirpath, f)
                if pat.search(open(p, errors="replace").read()):
                    offenders.append(os.path.relpath(p, ROOT))
    assert offenders == []
    # bench.py: the oracle appears only inside the CPU-baseline / reference-arm functions
    src = open(os.path.join(ROOT, "bench.py")).read()
    for m in re.finditer(r"^\s*from oracle import (\w+)", src, re.M):
        head = src[:m.start()]
        func = re.findall(r"^def (\w+)\(", head, re.M)[-1]
        assert func in ("cpu_probe_baseline", "kfd_walk_baseline", "run_reference"), func
