#!/usr/bin/env python3
"""Pack the reference's captured kfd sysfs trees into deterministic archives.

Run HERE (build container, where /root/reference is mounted):

    python tests/golden/make_fixtures.py

It reads /root/reference/testdata/{topology-parsing, topology-parsing-mi308,
topo-mi210-xgmi-pcie, topo-mi300-cpx, debugfs-parsing} and writes one
`tests/golden/<name>.tar.gz` per tree, holding only the files the hot path reads
(node / io_links / p2p_links / mem_banks `properties`, gpu_id, name; the
`caches/` sub-trees are never read by the reference path and are left out).
The archives are data fixtures (captured sysfs text), not reference source.
`/root/reference` does not exist on the GPU box, so tests unpack these archives
(tests/kfd_fixtures.py) instead of touching the reference tree.
"""
import gzip
import io
import os
import sys
import tarfile

REF = "/root/reference/testdata"
HERE = os.path.dirname(os.path.abspath(__file__))
TREES = ["topology-parsing", "topology-parsing-mi308", "topo-mi210-xgmi-pcie", "topo-mi300-cpx",
         "debugfs-parsing"]


def pack(name: str) -> str:
    root = os.path.join(REF, name)
    members = []
    for dirpath, dirnames, filenames in os.walk(root):
        dirnames.sort()
        if "/caches" in dirpath or dirpath.endswith("/caches"):
            dirnames[:] = []
            continue
        for fn in sorted(filenames):
            if fn.endswith(".md"):
                continue
            full = os.path.join(dirpath, fn)
            members.append((os.path.relpath(full, REF), full))
    members.sort()
    raw = io.BytesIO()
    with tarfile.open(fileobj=raw, mode="w", format=tarfile.USTAR_FORMAT) as tf:
        for arc, full in members:
            data = open(full, "rb").read()
            ti = tarfile.TarInfo(arc)
            ti.size = len(data)
            ti.mtime = 0
            ti.mode = 0o644
            ti.uid = ti.gid = 0
            ti.uname = ti.gname = ""
            tf.addfile(ti, io.BytesIO(data))
    out = os.path.join(HERE, name + ".tar.gz")
    with open(out, "wb") as f:
        with gzip.GzipFile(fileobj=f, mode="wb", mtime=0, compresslevel=9) as gz:
            gz.write(raw.getvalue())
    return f"{name}: {len(members)} files -> {os.path.getsize(out)} bytes"


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference testdata not mounted; fixtures are already committed")
    for t in TREES:
        print(pack(t))
