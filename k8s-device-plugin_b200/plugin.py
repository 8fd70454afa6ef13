"""Mirror of internal/pkg/plugin/plugin.go and cmd/k8s-device-plugin/main.go above the C ABI.

`AMDGPUPlugin` keeps the reference's method names and semantics (DevicePluginServer interface,
api.pb.go:1419-1441); responses are produced as serialized v1beta1 protobuf bytes by the C++
core, so a gRPC server can pass them straight through (see `server.py`).
"""
import ctypes as C
import queue
import threading
from typing import Iterable, List, Optional

from . import _native as N
from . import v1beta1
from .context import Context, CycleStats

RESOURCE_NAMESPACE = "amd.com"     # plugin.go:406-408


def countGPUDevFromTopology(topoRoot: str = "/sys/class/kfd/kfd") -> int:
    """plugin.go:123-159."""
    v = C.c_int32()
    N.check(N.lib.b2dp_count_gpu_dev_from_topology(topoRoot.encode(), C.byref(v)))
    return v.value


def simpleHealthCheck(topoRoot: str = "/sys/class/kfd/kfd") -> bool:
    """plugin.go:161-206 (kfd root injectable)."""
    v = C.c_int32()
    N.check(N.lib.b2dp_simple_health_check(topoRoot.encode(), C.byref(v)))
    return bool(v.value)


class PluginError(Exception):
    """Go `error` returned from an RPC handler (becomes gRPC status Unknown + message)."""


class AMDGPUPlugin:
    """plugin.go:41-48.  One instance per resource name ("gpu" or "<compute>_<memory>")."""

    def __init__(self, ctx: Context, resource: str = "gpu", heartbeat: Optional["queue.Queue"] = None,
                 probe_timeout_ms: int = 0, probe_min_gbs: float = 0.0):
        self.ctx = ctx
        self.Resource = resource
        self.Heartbeat = heartbeat if heartbeat is not None else queue.Queue()
        self.signal = threading.Event()        # plugin.go:83-84 (SIGINT/SIGQUIT/SIGTERM)
        self.allocatorInitError = False
        self.probe_timeout_ms = probe_timeout_ms
        self.probe_min_gbs = probe_min_gbs
        self.last_stats: Optional[CycleStats] = None

    # plugin.go:82-91
    def Start(self):
        rc = self.ctx.start()
        if rc != N.OK:
            self.allocatorInitError = True     # "Falling back to kubelet default allocation"
        return None

    def Stop(self):                            # plugin.go:117-119
        return None

    # plugin.go:210-217
    def GetDevicePluginOptions(self) -> bytes:
        return v1beta1.DevicePluginOptions(
            get_preferred_allocation_available=not self.allocatorInitError).SerializeToString()

    # plugin.go:222-224
    def PreStartContainer(self, device_ids: List[str]) -> bytes:
        return b""

    # plugin.go:229-330
    def ListAndWatch(self) -> Iterable[bytes]:
        """Generator of serialized ListAndWatchResponse messages: the initial list, then one per
        heartbeat tick until `signal` is set.  A heterogeneous node with no devices for this
        resource sends nothing (plugin.go:296-298)."""
        wire, st = self.ctx.list_and_watch(self.Resource, N.LW_INITIAL)
        self.last_stats = st
        if st.n_devices or st.homogeneous:
            yield wire
        while not self.signal.is_set():
            try:
                tick = self.Heartbeat.get(timeout=0.05)
            except queue.Empty:
                continue
            if tick is None:                   # test hook: end of stream
                return
            wire, st = self.heartbeat_cycle()
            if st.n_devices or st.homogeneous:
                yield wire

    def heartbeat_cycle(self, external=None):
        """One heartbeat: node health + per-device health (GPU probe, or `external` map) + re-send
        (plugin.go:304-320)."""
        wire, st = self.ctx.list_and_watch(self.Resource, N.LW_HEARTBEAT, external=external,
                                           timeout_ms=self.probe_timeout_ms, min_gbs=self.probe_min_gbs)
        self.last_stats = st
        return wire, st

    # plugin.go:337-351
    def GetPreferredAllocation(self, container_requests) -> bytes:
        """container_requests: [(available_ids, must_include_ids, allocation_size)]."""
        out = b""
        for available, must_include, size in container_requests:
            try:
                ids = self.ctx.preferred_allocation(list(available), list(must_include), int(size))
            except N.B2dpError as e:
                raise PluginError("unable to get preferred allocation list. Error:%s" % e.message)
            out += v1beta1.wrap_bytes_field(
                1, v1beta1.ContainerPreferredAllocationResponse(deviceIDs=ids).SerializeToString())
        return out

    # plugin.go:356-393
    def Allocate(self, container_requests: List[List[str]]) -> bytes:
        out = b""
        for ids in container_requests:
            out += v1beta1.wrap_bytes_field(1, self.ctx.allocate_response(list(ids)))
        return out


class AMDGPULister:
    """plugin.go:398-438."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.ResUpdateChan: "queue.Queue" = queue.Queue()
        self.Heartbeat: "queue.Queue" = queue.Queue()

    def GetResourceNamespace(self) -> str:
        import sys
        return sys.modules[__name__].RESOURCE_NAMESPACE

    def NewPlugin(self, resourceLastName: str) -> AMDGPUPlugin:
        return AMDGPUPlugin(self.ctx, resourceLastName, self.Heartbeat)


# ---- cmd/k8s-device-plugin/main.go ---------------------------------------------------------
def ParseStrategy(s: str) -> str:
    """main.go:42-51."""
    if s in ("single", "mixed"):
        return s
    raise ValueError("invalid resource naming strategy: %s" % s)


def getResourceList(ctx: Context, strategy: str) -> List[str]:
    """main.go:53-91; raises B2dpError(E_HETEROGENEOUS) like the reference's error return."""
    return ctx.resource_list(strategy)
