"""Mirror of internal/pkg/exporter/health.go above the C ABI."""
import ctypes as C
from typing import Dict, List, Optional

from . import _native as N

Healthy, Unhealthy = "Healthy", "Unhealthy"     # v1beta1/constants.go:21-23


def gpu_states_to_map(gpu_states) -> Dict[str, str]:
    """health.go:74-80: [(Device, Health)] -> map; only exactly "healthy" is Healthy."""
    return {dev: (Healthy if health == Healthy.lower() else Unhealthy) for dev, health in gpu_states}


def PopulatePerGPUDHealth(dev_ids: List[str], defaultHealth: str, hMap: Optional[Dict[str, str]]) -> List[str]:
    """health.go:86-106.  `hMap` None = getGPUHealth() failed (socket absent / RPC error)."""
    n = len(dev_ids)
    ids = (N.Id64 * max(1, n))()
    for i, d in enumerate(dev_ids):
        ids[i].value = d.encode()
    m = len(hMap) if hMap is not None else 0
    sids = (N.Id64 * max(1, m))()
    sh = (C.c_int32 * max(1, m))()
    if hMap is not None:
        for j, (k, v) in enumerate(hMap.items()):
            sids[j].value = k.encode()
            sh[j] = 1 if v == Healthy else 0
    out = (C.c_int32 * max(1, n))()
    N.check(N.lib.b2dp_merge_health(ids, n, 1 if defaultHealth == Healthy else 0, 0 if hMap is None else 1,
                                    sids, sh, m, out))
    return [Healthy if out[i] else Unhealthy for i in range(n)]
