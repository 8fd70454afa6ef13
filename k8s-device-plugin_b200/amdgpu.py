"""Mirror of internal/pkg/amdgpu (amdgpu.go) above the C ABI."""
import ctypes as C

from . import _native as N
from .context import Context

KFD_ROOT = "/sys/class/kfd/kfd"   # amdgpu.go:102


class TopologyPropertyError(Exception):
    def __init__(self, code, value):
        super().__init__(N.lib.b2dp_strerror(code).decode())
        self.code, self.value = code, value


def ParseTopologyProperties(path: str, key: str):
    """amdgpu.go:442-463 with the regexp `<key>\\s(\\d+)`.  Returns (value, err) like Go."""
    v = C.c_int64(0)
    rc = N.lib.b2dp_parse_topology_property(path.encode(), key.encode(), C.byref(v))
    return v.value, (None if rc == N.OK else TopologyPropertyError(rc, v.value))


def GetDevIdsFromTopology(topoRoot: str = KFD_ROOT):
    """amdgpu.go:101-146 -> {render_minor: devID}."""
    rc, (mn, ids), n = _two(lambda cap: ((C.c_int32 * cap)(), ((C.c_char * 24) * cap)()),
                            lambda a, cap, pn: N.lib.b2dp_dev_ids_from_topology(topoRoot.encode(), a[0], a[1], cap, pn))
    N.check(rc)
    return {mn[i]: N.s(ids[i].value) for i in range(n)}


def GetNodeIdsFromTopology(topoRoot: str = KFD_ROOT):
    """amdgpu.go:496-538 -> {render_minor: node_id}."""
    rc, (mn, nd), n = _two(lambda cap: ((C.c_int32 * cap)(), (C.c_int32 * cap)()),
                           lambda a, cap, pn: N.lib.b2dp_node_ids_from_topology(topoRoot.encode(), a[0], a[1], cap, pn))
    N.check(rc)
    return {mn[i]: nd[i] for i in range(n)}


def _two(make, call):
    cap = 128
    while True:
        arrs = make(cap)
        n = C.c_int(0)
        rc = call(arrs, cap, C.byref(n))
        if rc == N.E_NOSPC:
            cap = max(n.value, cap * 2)
            continue
        return rc, arrs, n.value


def parseDebugFSFirmwareInfo(path: str):
    """amdgpu.go:467-490 -> (feat, fw)."""
    rc, arr, n = N.grow_call(lambda cap: (N.FwEntry * cap)(),
                             lambda a, cap, pn: N.lib.b2dp_parse_debugfs_firmware_info(path.encode(), a, cap, pn))
    N.check(rc)
    return ({N.s(e.name): e.feature for e in arr[:n]}, {N.s(e.name): e.firmware for e in arr[:n]})


def GetAMDGPUs(ctx: Context):
    """amdgpu.go:149-268 on the context's backend."""
    return ctx.enumerate()


def UniquePartitionConfigCount(devices: dict):
    """amdgpu.go:270-285 (pure function of the device map)."""
    out = {}
    for d in devices.values():
        c, m = d["computePartitionType"], d["memoryPartitionType"]
        if c != "" and m != "":
            out[c + "_" + m] = out.get(c + "_" + m, 0) + 1
    return out


def IsHomogeneous(ctx: Context) -> bool:
    """amdgpu.go:287-293."""
    return ctx.is_homogeneous()


def IsComputePartitionSupported(ctx: Context) -> bool:
    """amdgpu.go:295-311."""
    return ctx.partition_supported(0)


def IsMemoryPartitionSupported(ctx: Context) -> bool:
    """amdgpu.go:313-328."""
    return ctx.partition_supported(1)
