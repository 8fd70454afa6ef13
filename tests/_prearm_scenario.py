"""The prearm=1 scenario of tests/test_gpu_probe.py::test_prearmed_passes_equal_launched_passes, run as its OWN process:
while a pass is armed, device-synchronising CUDA calls of the same process wait for the doorbell, so the scenario must not
share a process with pytest's torch context -- and a hang here becomes a clean test failure (subprocess timeout)."""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import probe as oprobe  # noqa: E402

MiB = 1 << 20


def main():
    P = importlib.import_module("k8s-device-plugin_b200")
    with P.Context("nvml:") as e:                       # device count and free memory without CUDA or torch
        n = len(e.enumerate())
        total0 = e.probe_describe(0)["total_memory"]
    nbytes = 96 * MiB + 16 * 13
    n_words = nbytes // 4
    PRE = P._native.RES_PREARMED
    # The two contexts run ONE AFTER THE OTHER: while a pass is armed, its stream wait stalls every other piece of GPU
    # work this process submits to that GPU (measured: other streams, torch, a second library context all wait for the
    # doorbell; other PROCESSES are not affected).  prearm is for a process whose only GPU user is this library.
    with P.Context("cuda:bytes=%d,prearm=0" % nbytes) as b:
        want = []
        for step in range(6):
            rb = b.probe_health(timed=False, min_gbs=1e-3)
            assert not any(r.flags & PRE for r in rb)
            want.append([(r.seed, r.checksum, r.expected_checksum, r.mismatches, r.healthy) for r in rb])
    with P.Context("cuda:bytes=%d,prearm=1" % nbytes) as a:
        seeds = [oprobe.initial_seed(i) for i in range(n)]
        for step in range(6):
            ra = a.probe_health(timed=False, min_gbs=1e-3)
            assert [(r.seed, r.checksum, r.expected_checksum, r.mismatches, r.healthy) for r in ra] == want[step]
            assert [r.seed for r in ra] == seeds and all(r.healthy for r in ra)
            assert all(bool(r.flags & PRE) == (step > 0) for r in ra)
            seeds = [oprobe.next_seed(s) for s in seeds]
        # peek discards the armed pass; the buffer is what the oracle says and the sequence goes on
        src = a.probe_peek(0, 0, n_words)
        assert np.array_equal(src, oprobe.pattern(n_words, seeds[0]))
        ra = a.probe_health(timed=False, min_gbs=1e-3)
        assert [r.seed for r in ra] == seeds and not (ra[0].flags & PRE) and all(r.healthy for r in ra)
        cs, bad, first, dst = oprobe.probe_pass(src, seeds[0], oprobe.next_seed(seeds[0]))
        assert ra[0].checksum == cs
        seeds = [oprobe.next_seed(s) for s in seeds]
        assert np.array_equal(a.probe_peek(0, 0, n_words), dst)
        # a fault poked in front of an armed pass is caught by the next pass, reported once, repaired
        a.probe_health(timed=False, min_gbs=1e-3)
        seeds = [oprobe.next_seed(s) for s in seeds]
        a.probe_inject_fault(n - 1, 4242, 0x2)
        ra = a.probe_health(timed=False, min_gbs=1e-3)
        assert (ra[-1].healthy, ra[-1].mismatches, ra[-1].first_bad_word) == (False, 1, 4242) and [r.seed for r in ra] == seeds
        seeds = [oprobe.next_seed(s) for s in seeds]
        for _ in range(3):
            ra = a.probe_health(timed=False, min_gbs=1e-3)
            assert all(r.healthy for r in ra) and [r.seed for r in ra] == seeds
            seeds = [oprobe.next_seed(s) for s in seeds]
        assert ra[0].flags & PRE
        # passes with other options interleave (each discards the armed pass, none loses a step)
        for kw in (dict(timed=True), dict(variant=1), dict(via_workers=True), dict(grid_ctas=64), dict()):
            ra = a.probe_health(min_gbs=1e-3, **({"timed": False} | kw))
            assert all(r.healthy for r in ra) and [r.seed for r in ra] == seeds, kw
            seeds = [oprobe.next_seed(s) for s in seeds]
        a.probe_reset(-1)
        ra = a.probe_health(timed=False, min_gbs=1e-3)
        assert all(r.healthy for r in ra) and [r.seed for r in ra] == seeds
        seeds = [oprobe.next_seed(s) for s in seeds]
        if n > 1:
            gbs, lt, mm = a.p2p_matrix(bytes_per_pair=32 * MiB)
            assert (mm == 0).all()
            ra = a.probe_health(timed=False, min_gbs=1e-3)
            assert all(r.healthy for r in ra) and [r.seed for r in ra] == seeds
            seeds = [oprobe.next_seed(s) for s in seeds]
        # ListAndWatch heartbeats ride on armed passes
        for _ in range(3):
            wire, st = a.list_and_watch("gpu", P._native.LW_HEARTBEAT, min_gbs=1e-3)
            assert st.n_unhealthy == 0 and st.n_devices == n
    # a pass that misses its deadline while the next one is already armed behind it
    big = 4 << 30
    if total0 > 2 * big + (8 << 30):
        with P.Context("cuda:devices=0,bytes=%d,prearm=1" % big) as c:
            assert c.probe_health(timed=False, min_gbs=1e-3)[0].healthy           # arms the next
            (r,) = c.probe_health(timed=False, timeout_ms=1, min_gbs=1e-3)       # rung; ~1.4 ms > 1 ms deadline
            assert r.err == P._native.E_TIMEOUT and not r.healthy
            time.sleep(0.3)
            for _ in range(20):
                (r,) = c.probe_health(timed=False, min_gbs=1e-3)
                if r.err == 0:
                    break
                time.sleep(0.05)
            assert r.err == 0 and r.healthy and r.checksum == r.expected_checksum
    print("PREARM_OK", n, flush=True)


if __name__ == "__main__":
    main()
