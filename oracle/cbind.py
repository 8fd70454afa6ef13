"""ctypes loaders for the C restatements (oracle/probe_oracle.c, oracle/kfd_walk.c).
Test infrastructure / bench CPU baseline only."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")


def build():
    r = subprocess.run(["make", "-C", _HERE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the C oracle failed:\n" + r.stdout)


def probe_lib():
    lib = C.CDLL(os.path.join(_BUILD, "libprobe_oracle.so"))
    lib.oracle_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int]
    lib.oracle_fill.restype = None
    lib.oracle_probe_pass.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int,
                                      C.POINTER(C.c_uint64)]
    lib.oracle_probe_pass.restype = None
    lib.oracle_expected_checksum.argtypes = [C.c_uint64, C.c_uint32, C.c_int]
    lib.oracle_expected_checksum.restype = C.c_uint64
    return lib


def kfd_lib():
    lib = C.CDLL(os.path.join(_BUILD, "libkfd_walk.so"))
    lib.kfdwalk_enumerate.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    lib.kfdwalk_health.argtypes = [C.c_char_p]
    lib.kfdwalk_pair_weights.argtypes = [C.c_char_p, C.POINTER(C.c_longlong)]
    lib.kfdwalk_cycle.argtypes = [C.c_char_p, C.c_int]
    return lib
