"""Mirror of internal/pkg/allocator (allocator.go, device.go, besteffort_policy.go) above the C ABI."""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

from . import _native as N


@dataclass
class Device:                       # device.go:56-65
    Id: str = ""
    NodeId: int = 0
    NumaNode: int = 0
    DevId: str = ""
    Card: int = 0
    RenderD: int = 0
    ComputePartitionType: str = ""
    MemoryPartitionType: str = ""


class AllocError(Exception):
    """Go `error` from the allocator; str(e) is the reference's message."""

    def __init__(self, code):
        super().__init__(N.lib.b2dp_strerror(code).decode())
        self.code = code


def _to_abi(devs: List[Device]):
    arr = (N.Device * max(1, len(devs)))()
    for i, d in enumerate(devs):
        arr[i].id = d.Id.encode()
        arr[i].dev_id = d.DevId.encode()
        arr[i].card, arr[i].render_d, arr[i].node_id, arr[i].numa_node = d.Card, d.RenderD, d.NodeId, d.NumaNode
        arr[i].compute_partition = d.ComputePartitionType.encode()
        arr[i].memory_partition = d.MemoryPartitionType.encode()
    return arr


class BestEffortPolicy:
    """besteffort_policy.go:45-151.  Init/Allocate return Go-style (value, err)."""

    def __init__(self):
        self._h = C.c_void_p()
        N.check(N.lib.b2dp_allocator_new(C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None):
            N.lib.b2dp_allocator_free(self._h)
            self._h = None

    def Init(self, devs: List[Device], topoDir: str) -> Optional[AllocError]:
        rc = N.lib.b2dp_allocator_init(self._h, _to_abi(devs), len(devs), topoDir.encode())
        return None if rc == N.OK else AllocError(rc)

    def InitLinks(self, devs: List[Device], links) -> Optional[AllocError]:
        """Init from a measured link list [(node_from, node_to, type)] (the P2P matrix)."""
        arr = (N.Link * max(1, len(links)))()
        for i, (a, b, t) in enumerate(links):
            arr[i].node_from, arr[i].node_to, arr[i].type = a, b, t
        rc = N.lib.b2dp_allocator_init_links(self._h, _to_abi(devs), len(devs), arr, len(links))
        return None if rc == N.OK else AllocError(rc)

    def Allocate(self, availableIds: List[str], requiredIds: Optional[List[str]], size: int):
        requiredIds = requiredIds or []
        out = (N.Id64 * max(1, len(availableIds), len(requiredIds)))()
        n = C.c_int(0)
        rc = N.lib.b2dp_allocator_allocate(self._h, N.str_array(availableIds), len(availableIds),
                                           N.str_array(requiredIds), len(requiredIds), size, out, len(out), C.byref(n))
        if rc != N.OK:
            return [], AllocError(rc)
        return [N.s(out[i].value) for i in range(n.value)], None

    # ---- introspection used by the parity tests (device_test.go) --------------------------
    def pair_weights(self):
        """p2pWeights as {from: {to: weight}} (device.go:220-252)."""
        cap = 4096
        while True:
            arr = (N.PairWeight * cap)()
            n, rows = C.c_int(0), C.c_int(0)
            rc = N.lib.b2dp_allocator_pair_weights(self._h, arr, cap, C.byref(n), C.byref(rows))
            if rc == N.E_NOSPC:
                cap = n.value
                continue
            N.check(rc)
            break
        w = {}
        for e in arr[:n.value]:
            w.setdefault(e.node_from, {})[e.node_to] = e.weight
        assert len(w) == rows.value
        return w

    def group_count(self) -> int:
        v = C.c_int32()
        N.check(N.lib.b2dp_allocator_group_count(self._h, C.byref(v)))
        return v.value

    def candidates(self, availableIds, requiredIds, size):
        """device.go:353-442 -> ((n_candidates, best_weight), err)."""
        requiredIds = requiredIds or []
        nc, bw = C.c_int32(), C.c_int32()
        rc = N.lib.b2dp_allocator_candidates(self._h, N.str_array(availableIds), len(availableIds),
                                             N.str_array(requiredIds), len(requiredIds), size, C.byref(nc), C.byref(bw))
        if rc != N.OK:
            return (0, 0), AllocError(rc)
        return (nc.value, bw.value), None


def NewBestEffortPolicy() -> BestEffortPolicy:
    """besteffort_policy.go:52-59."""
    return BestEffortPolicy()
