#!/usr/bin/env python3
"""BASELINE.json configs[0]: topology parse on the reference's fixtures, CPU only.

For each captured kfd tree (tests/golden/*.tar.gz unpacked to /dev/shm, page-cache-hot like sysfs)
times, single-threaded, median of K after 3 warm-ups:

  product   : libb200dp `kfd:` backend through the C ABI -- b2dp_enumerate (GetAMDGPUs),
              b2dp_list_and_watch(INITIAL) (stream start), heartbeat without probe, b2dp_start
              (pair weights + grouping), one b2dp_preferred_allocation
  port      : oracle/kfd_walk.c, the reference-shaped C port (same passes/opens as the Go code)
  oracle_py : the Python oracle (for scale only)

and reports algorithmic bytes (every node/link `properties` file read once) per tree.
    python tests/bench_kfd_parse.py [--out profiles/rNN_kfd_parse_cpu.json]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import statistics
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def med(fn, k):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(k):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return round(statistics.median(ts), 4)


def tree_bytes(nodes_dir):
    total = files = 0
    for dp, dn, fn in os.walk(nodes_dir, followlinks=True):
        if "properties" in fn and ("io_links" in dp or "p2p_links" in dp or os.path.dirname(dp).rstrip("/").endswith("nodes")
                                   or os.path.basename(os.path.dirname(dp)) == "nodes"):
            total += os.path.getsize(os.path.join(dp, "properties"))
            files += 1
    return total, files


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import fake_sysfs
    import kfd_fixtures
    from oracle import allocator as oalloc, amdgpu as oamd, cbind, plugin as oplug
    from test_oracle_golden import topo_dir
    pkg = importlib.import_module("k8s-device-plugin_b200")
    N = pkg._native
    cbind.build()
    k = cbind.kfd_lib()
    rows = []
    cases = [("topology-parsing (3 nodes)", "tp", {}), ("topo-mi210-xgmi-pcie (8 GPUs)", "mi210", {}),
             ("topology-parsing-mi308 (32 partitions)", "mi308", dict(compute="cpx", memory="nps1")),
             ("topo-mi300-cpx (63 partitions)", "cpx", dict(compute="cpx", memory="nps4"))]
    base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    for label, name, kw in cases:
        nodes = kfd_fixtures.root("topology-parsing") + "/topology/nodes" if name == "tp" else topo_dir(kfd_fixtures, name)
        root = fake_sysfs.build(os.path.join(base, "b2dp_parse_bench_" + name), nodes, **kw)
        rb = root.encode()
        nbytes, nfiles = tree_bytes(nodes)
        ctx = pkg.Context("kfd:" + root)
        ids = sorted(ctx.enumerate())
        buf = C.create_string_buffer(1 << 16)
        out3 = (C.c_longlong * 3)()
        row = {"fixture": label, "devices": len(ids), "property_files": nfiles, "algorithmic_bytes": nbytes}
        row["product_enumerate_ms"] = med(lambda: ctx.enumerate_raw(), args.reps)
        row["product_stream_start_ms"] = med(lambda: ctx.list_and_watch("gpu", N.LW_INITIAL), args.reps)
        row["product_heartbeat_ms"] = med(lambda: ctx.list_and_watch("gpu", N.LW_HEARTBEAT | N.LW_NO_PROBE), args.reps)
        row["port_stream_start_ms"] = med(lambda: k.kfdwalk_cycle(rb, 1), args.reps)
        row["port_heartbeat_ms"] = med(lambda: k.kfdwalk_cycle(rb, 0), args.reps)
        if len(ids) > 1:
            row["product_start_pair_weights_ms"] = med(lambda: ctx.start(), max(5, args.reps // 3))
            row["port_start_pair_weights_ms"] = med(lambda: k.kfdwalk_pair_weights(rb, out3), max(5, args.reps // 3))
            size = min(3, len(ids) - 1)
            row["product_preferred_allocation_ms"] = med(lambda: ctx.preferred_allocation(ids, [], size), args.reps)
            opol = oalloc.BestEffortPolicy()
            t0 = time.perf_counter()
            opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes")
            row["oracle_py_start_pair_weights_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        t0 = time.perf_counter()
        oamd.GetAMDGPUs(root)
        oamd.GetAMDGPUs(root)
        row["oracle_py_stream_start_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
        row["stream_start_speedup_vs_port"] = round(row["port_stream_start_ms"] / max(row["product_stream_start_ms"], 1e-6), 2)
        row["stream_start_MBps_product"] = round(nbytes / max(row["product_stream_start_ms"], 1e-6) / 1e3, 1)
        rows.append(row)
        ctx.close()
    out = {"host_cores": os.cpu_count(), "threads_used": 1, "rows": rows,
           "note": "port = C restatement of the Go walk (no Go toolchain in the image), single-threaded like the reference's single goroutine; the product is single-threaded too except pair-weight init, which parses a tree of >= 512 link files on up to 8 threads (records applied in the reference's order)"}
    s = json.dumps(out, indent=1)
    print(s)
    if args.out:
        open(args.out, "w").write(s + "\n")


if __name__ == "__main__":
    main()
