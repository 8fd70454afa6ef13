"""gRPC host for the plugin: the kubelet-facing v1beta1.DevicePlugin
server, the Registration client, and a metricssvc.MetricsService server for the exporter socket.

Mirrors what the vendored dpm does around the reference plugin
(vendor/github.com/kubevirt/device-plugin-manager/pkg/dpm/plugin.go:63-191): listen on
`<plugin dir>/amd.com_<name>`, serve DevicePlugin, dial `kubelet.sock` and Register
{version, endpoint, resource_name, options}.  Responses for ListAndWatch / Allocate /
GetPreferredAllocation are the C++-encoded protobuf bytes passed straight through (identity
serializers), so the per-heartbeat host work is one C-ABI call and one socket write.
"""
import os
import threading
from concurrent import futures

import grpc
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

from . import v1beta1
from . import plugin as _plugin
from .plugin import AMDGPUPlugin, PluginError

_ident = lambda b: b  # noqa: E731


class DevicePluginServicer(grpc.GenericRpcHandler):
    """Generic handler for /v1beta1.DevicePlugin/* backed by an AMDGPUPlugin."""

    def __init__(self, plugin: AMDGPUPlugin):
        self.plugin = plugin
        self._methods = {
            v1beta1.GET_OPTIONS: grpc.unary_unary_rpc_method_handler(self._options, _ident, _ident),
            v1beta1.LIST_AND_WATCH: grpc.unary_stream_rpc_method_handler(self._list_and_watch, _ident, _ident),
            v1beta1.GET_PREFERRED_ALLOCATION: grpc.unary_unary_rpc_method_handler(self._preferred, _ident, _ident),
            v1beta1.ALLOCATE: grpc.unary_unary_rpc_method_handler(self._allocate, _ident, _ident),
            v1beta1.PRE_START_CONTAINER: grpc.unary_unary_rpc_method_handler(self._prestart, _ident, _ident),
        }

    def service(self, handler_call_details):
        return self._methods.get(handler_call_details.method)

    def _options(self, request, context):
        return self.plugin.GetDevicePluginOptions()

    def _list_and_watch(self, request, context):
        for wire in self.plugin.ListAndWatch():
            if not context.is_active():
                return
            yield wire

    def _preferred(self, request, context):
        req = v1beta1.PreferredAllocationRequest.FromString(request)
        try:
            return self.plugin.GetPreferredAllocation(
                [(list(c.available_deviceIDs), list(c.must_include_deviceIDs), c.allocation_size)
                 for c in req.container_requests])
        except PluginError as e:       # Go: return nil, fmt.Errorf(...) => status Unknown + message
            context.abort(grpc.StatusCode.UNKNOWN, str(e))

    def _allocate(self, request, context):
        req = v1beta1.AllocateRequest.FromString(request)
        return self.plugin.Allocate([list(c.devices_ids) for c in req.container_requests])

    def _prestart(self, request, context):
        return self.plugin.PreStartContainer([])


class PluginServer:
    """One resource's plugin server: dpm/plugin.go StartServer / register / StopServer."""

    def __init__(self, plugin: AMDGPUPlugin, plugin_dir: str = v1beta1.DevicePluginPath,
                 kubelet_socket: str = None):
        self.plugin = plugin
        self.plugin_dir = plugin_dir
        self.endpoint = "%s_%s" % (_plugin.RESOURCE_NAMESPACE, plugin.Resource)           # dpm/plugin.go:51-59
        self.socket_path = os.path.join(plugin_dir, self.endpoint)
        self.kubelet_socket = kubelet_socket or os.path.join(plugin_dir, "kubelet.sock")
        self.resource_name = "%s/%s" % (_plugin.RESOURCE_NAMESPACE, plugin.Resource)
        self._server = None

    def start(self):
        self.plugin.Start()                                   # dpm/manager.go:182-193: Start() before serve
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)                       # dpm/plugin.go cleanup()
        self._server = grpc.server(futures.ThreadPoolExecutor(max_workers=8))
        self._server.add_generic_rpc_handlers((DevicePluginServicer(self.plugin),))
        self._server.add_insecure_port("unix://" + self.socket_path)
        self._server.start()
        return self

    def register(self, timeout=5.0):
        """dpm/plugin.go:127-162: Registration.Register on kubelet.sock."""
        opts = v1beta1.DevicePluginOptions.FromString(self.plugin.GetDevicePluginOptions())
        req = v1beta1.RegisterRequest(version=v1beta1.Version, endpoint=self.endpoint,
                                      resource_name=self.resource_name, options=opts)
        with grpc.insecure_channel("unix://" + self.kubelet_socket) as ch:
            call = ch.unary_unary(v1beta1.REGISTER, request_serializer=lambda m: m.SerializeToString(),
                                  response_deserializer=v1beta1.Empty.FromString)
            call(req, timeout=timeout)

    def stop(self):
        self.plugin.signal.set()
        if self._server:
            self._server.stop(grace=0.2)
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)


# ---- metricssvc (internal/pkg/exporter/metricssvc/metricssvc.pb.go:95-110,179-186,227-236,284-291) ----
def _build_metricssvc():
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name, fdp.package, fdp.syntax = "b200dp/metricssvc.proto", "metricssvc", "proto3"
    T = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fdp.message_type.add()
        m.name = name
        for fname, num, rep, tname in fields:
            f = m.field.add()
            f.name, f.number = fname, num
            f.label = T.LABEL_REPEATED if rep else T.LABEL_OPTIONAL
            if tname:
                f.type, f.type_name = T.TYPE_MESSAGE, ".metricssvc." + tname
            else:
                f.type = T.TYPE_STRING
    msg("GPUState", [("ID", 1, 0, None), ("UUID", 2, 0, None), ("Health", 3, 0, None),
                     ("AssociatedWorkload", 4, 1, None), ("Device", 5, 0, None)])
    msg("GPUGetRequest", [("ID", 1, 1, None)])
    msg("GPUUpdateRequest", [("ID", 1, 1, None), ("Health", 2, 1, None)])
    msg("GPUStateResponse", [("GPUState", 1, 1, "GPUState")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return {m.name: message_factory.GetMessageClass(pool.FindMessageTypeByName("metricssvc." + m.name))
            for m in fdp.message_type}


_MS = _build_metricssvc()
GPUState, GPUGetRequest, GPUStateResponse = _MS["GPUState"], _MS["GPUGetRequest"], _MS["GPUStateResponse"]
METRICS_LIST = "/metricssvc.MetricsService/List"                    # metricssvc_grpc.pb.go:46
METRICS_GET = "/metricssvc.MetricsService/GetGPUState"              # metricssvc_grpc.pb.go:45
EXPORTER_SOCKET = "/var/lib/amd-metrics-exporter/amdgpu_device_metrics_exporter_grpc.socket"   # health.go:36


class MetricsServer:
    """Serves the B200 probe's verdicts in the exporter's wire format, so an UNMODIFIED reference
    plugin (exporter/health.go:42-82) consumes them: Health is "healthy"/"unhealthy" (lower case,
    health.go:75), Device is the kubelet device id (health.go:98)."""

    def __init__(self, ctx, socket_path=EXPORTER_SOCKET, min_gbs=0.0):
        self.ctx, self.socket_path, self.min_gbs = ctx, socket_path, min_gbs
        self._server = None
        self._lock = threading.Lock()

    def _states(self):
        with self._lock:
            ids = sorted(self.ctx.enumerate())
            res = self.ctx.probe_health(min_gbs=self.min_gbs)
        return [GPUState(ID=str(r.device), UUID=ids[r.device], Health="healthy" if r.healthy else "unhealthy",
                         Device=ids[r.device]) for r in res]

    def start(self):
        os.makedirs(os.path.dirname(self.socket_path), exist_ok=True)
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
        outer = self

        class H(grpc.GenericRpcHandler):
            def service(self, d):
                if d.method == METRICS_LIST:
                    return grpc.unary_unary_rpc_method_handler(
                        lambda req, c: GPUStateResponse(GPUState=outer._states()).SerializeToString(), _ident, _ident)
                if d.method == METRICS_GET:
                    def get(req, c):
                        want = set(GPUGetRequest.FromString(req).ID)
                        return GPUStateResponse(
                            GPUState=[s for s in outer._states() if not want or s.ID in want]).SerializeToString()
                    return grpc.unary_unary_rpc_method_handler(get, _ident, _ident)
                return None
        self._server = grpc.server(futures.ThreadPoolExecutor(max_workers=4))
        self._server.add_generic_rpc_handlers((H(),))
        self._server.add_insecure_port("unix://" + self.socket_path)
        self._server.start()
        return self

    def stop(self):
        if self._server:
            self._server.stop(grace=0.2)
        if os.path.exists(self.socket_path):
            os.unlink(self.socket_path)
