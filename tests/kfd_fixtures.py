"""Unpack the packed kfd fixture trees (tests/golden/*.tar.gz) into a scratch dir.

Paths mirror the reference's testdata layout, so tests read like the reference's:
    root("topology-parsing-mi308") + "/topology/nodes"   (device_test.go:83)
"""
import os
import tarfile
import tempfile
import threading

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_lock = threading.Lock()
_base = None

TREES = ["topology-parsing", "topology-parsing-mi308", "topo-mi210-xgmi-pcie", "topo-mi300-cpx",
         "debugfs-parsing"]


def base_dir() -> str:
    """Directory holding every unpacked tree (prefers /dev/shm: page-cache-hot like sysfs)."""
    global _base
    with _lock:
        if _base is None:
            parent = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
            d = tempfile.mkdtemp(prefix="b2dp_kfd_", dir=parent)
            for t in TREES:
                with tarfile.open(os.path.join(_GOLDEN, t + ".tar.gz"), "r:gz") as tf:
                    tf.extractall(d, filter="data")
            _base = d
        return _base


def root(name: str) -> str:
    return os.path.join(base_dir(), name)


def cleanup() -> None:
    global _base
    import shutil
    with _lock:
        if _base:
            shutil.rmtree(_base, ignore_errors=True)
            _base = None
