"""Oracle for the GPU health probe arithmetic (numpy; test infrastructure only).

The reference has no GPU probe: its health verdict is the text check restated in
oracle/plugin.py:simpleHealthCheck plus an external exporter's string.  The B200
probe replaces that evidence with a full HBM stream-verify; THIS file is the
specification of its integer arithmetic, against which the CUDA kernels
(k8s-device-plugin_b200/csrc/hbm_probe.cuh) must be bit-exact:

    word(i, seed)  = (uint32(i) * 2654435761) ^ seed              i = 32-bit word index
    probe pass     : checksum   = sum(src[i])            mod 2^64
                     mismatches = #{ i : src[i] != word(i, seed) }
                     first_bad  = min such i, else 2^64-1
                     dst[i]     = src[i] ^ (seed ^ next_seed)
    seed schedule  : next_seed  = seed * 1664525 + 1013904223     mod 2^32
    initial seed   : 0x5EED0000 | device_index   

Link classification for the P2P matrix (replaces the kfd link `type` read at
internal/pkg/allocator/device.go:143-149; type codes are the reference's):
    peer access + >= 25 % of 770 GB/s  -> 11 (XGMI-class: NVLink)
    peer access below that             ->  2 (PCIe-class)
    no peer access                     ->  0 (other)
"""
import numpy as np

PATTERN_MUL = 2654435761
SEED_BASE = 0x5EED0000
LCG_A, LCG_C = 1664525, 1013904223
NO_BAD = (1 << 64) - 1

LINK_NVLINK, LINK_PCIE, LINK_OTHER = 11, 2, 0
NVLINK_REF_GBS = 770.0
NVLINK_CLASS_FRACTION = 0.25


def initial_seed(device_index: int) -> int:
    return (SEED_BASE | (device_index & 0xFFFF)) & 0xFFFFFFFF


def next_seed(seed: int) -> int:
    return (seed * LCG_A + LCG_C) & 0xFFFFFFFF


def pattern(n_words: int, seed: int, start: int = 0) -> np.ndarray:
    idx = (np.arange(start, start + n_words, dtype=np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    with np.errstate(over="ignore"):
        return (idx * np.uint32(PATTERN_MUL)) ^ np.uint32(seed)


def checksum(buf: np.ndarray) -> int:
    return int(buf.astype(np.uint64).sum(dtype=np.uint64))


def expected_checksum(n_words: int, seed: int, chunk: int = 1 << 24) -> int:
    """Checksum of a clean buffer, computed in chunks so 2^28 words never sit in memory."""
    total = 0
    for start in range(0, n_words, chunk):
        n = min(chunk, n_words - start)
        total = (total + checksum(pattern(n, seed, start))) & NO_BAD
    return total


def probe_pass(src: np.ndarray, seed: int, nxt: int):
    """One probe launch over a uint32 buffer -> (checksum, mismatches, first_bad, dst)."""
    exp = pattern(src.size, seed)
    bad = np.nonzero(src != exp)[0]
    dst = src ^ np.uint32(seed ^ nxt)
    return checksum(src), int(bad.size), (int(bad[0]) if bad.size else NO_BAD), dst


def classify_link(can_access_peer: bool, gbs: float) -> int:
    if not can_access_peer:
        return LINK_OTHER
    return LINK_NVLINK if gbs >= NVLINK_CLASS_FRACTION * NVLINK_REF_GBS else LINK_PCIE


MIN_FRAC = 0.8                 # BASELINE.json: "each per-GPU probe >= 80 % of HBM peak GB/s"
FLOOR_MIN_BYTES = 128 << 20    # the fractional floor needs a ring slot above the 126 MB L2 (a pass that measures HBM)


def health_floor(slot_bytes: int, gbs_ref: float, min_frac: float = MIN_FRAC, abs_min_gbs: float = 0.0) -> float:
    """The GB/s floor of the product's verdict: an absolute `min_gbs` if one is given, else min_frac x the device's
    calibrated ceiling when the ring streams from HBM, else none."""
    if abs_min_gbs > 0:
        return abs_min_gbs
    return min_frac * gbs_ref if slot_bytes >= FLOOR_MIN_BYTES and gbs_ref > 0 else 0.0


def probe_healthy(launch_ok: bool, got_checksum: int, mismatches: int, n_words: int, seed: int,
                  gbs: float, min_gbs: float) -> bool:
    """Verdict rule of the product: the launch completed, every word verified,
    the checksum matches the closed form, and the stream ran at >= min_gbs (= health_floor(...))."""
    return bool(launch_ok and mismatches == 0 and got_checksum == expected_checksum(n_words, seed)
                and gbs >= min_gbs)
