"""The kubelet device-plugin v1beta1 wire contract, built programmatically.

Message and field numbers restate vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/api.proto
(services Registration :24-26 and DevicePlugin :51-77; messages :28-223).  There is no protoc /
grpc_tools in the image, so the FileDescriptorProto is assembled by hand and turned into message
classes with the protobuf runtime; gRPC method handlers/stubs are wired with grpcio's generic API.
The hot path itself does not use these classes: ListAndWatch / Allocate responses are serialized
in C++ (csrc/pbwire.hpp) and pass through gRPC as raw bytes; the classes here are for requests,
for the fake kubelet in the tests, and to check the C++ encoder byte-for-byte.
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

Version = "v1beta1"                                        # constants.go:26
Healthy, Unhealthy = "Healthy", "Unhealthy"                # constants.go:21-23
DevicePluginPath = "/var/lib/kubelet/device-plugins/"      # constants.go:30
KubeletSocket = DevicePluginPath + "kubelet.sock"          # constants.go:32

_T = descriptor_pb2.FieldDescriptorProto
_STR, _BOOL, _I32, _I64, _MSG = _T.TYPE_STRING, _T.TYPE_BOOL, _T.TYPE_INT32, _T.TYPE_INT64, _T.TYPE_MESSAGE


def _build():
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name = "b200dp/v1beta1/api.proto"
    fdp.package = "v1beta1"
    fdp.syntax = "proto3"

    def msg(name, fields, maps=()):
        m = fdp.message_type.add()
        m.name = name
        for fname, num, ftype, rep, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type = fname, num, ftype
            f.label = _T.LABEL_REPEATED if rep else _T.LABEL_OPTIONAL
            if tname:
                f.type_name = ".v1beta1." + tname
        for fname, num in maps:   # map<string,string>
            entry = m.nested_type.add()
            entry.name = "".join(p.capitalize() for p in fname.split("_")) + "Entry"
            entry.options.map_entry = True
            for en, n in (("key", 1), ("value", 2)):
                ef = entry.field.add()
                ef.name, ef.number, ef.type, ef.label = en, n, _STR, _T.LABEL_OPTIONAL
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, _MSG, _T.LABEL_REPEATED
            f.type_name = ".v1beta1.%s.%s" % (name, entry.name)
        return m

    msg("DevicePluginOptions", [("pre_start_required", 1, _BOOL, 0, None),
                                ("get_preferred_allocation_available", 2, _BOOL, 0, None)])
    msg("RegisterRequest", [("version", 1, _STR, 0, None), ("endpoint", 2, _STR, 0, None),
                            ("resource_name", 3, _STR, 0, None), ("options", 4, _MSG, 0, "DevicePluginOptions")])
    msg("Empty", [])
    msg("ListAndWatchResponse", [("devices", 1, _MSG, 1, "Device")])
    msg("TopologyInfo", [("nodes", 1, _MSG, 1, "NUMANode")])
    msg("NUMANode", [("ID", 1, _I64, 0, None)])
    msg("Device", [("ID", 1, _STR, 0, None), ("health", 2, _STR, 0, None), ("topology", 3, _MSG, 0, "TopologyInfo")])
    msg("PreStartContainerRequest", [("devices_ids", 1, _STR, 1, None)])
    msg("PreStartContainerResponse", [])
    msg("PreferredAllocationRequest", [("container_requests", 1, _MSG, 1, "ContainerPreferredAllocationRequest")])
    msg("ContainerPreferredAllocationRequest", [("available_deviceIDs", 1, _STR, 1, None),
                                                ("must_include_deviceIDs", 2, _STR, 1, None),
                                                ("allocation_size", 3, _I32, 0, None)])
    msg("PreferredAllocationResponse", [("container_responses", 1, _MSG, 1, "ContainerPreferredAllocationResponse")])
    msg("ContainerPreferredAllocationResponse", [("deviceIDs", 1, _STR, 1, None)])
    msg("AllocateRequest", [("container_requests", 1, _MSG, 1, "ContainerAllocateRequest")])
    msg("ContainerAllocateRequest", [("devices_ids", 1, _STR, 1, None)])
    msg("CDIDevice", [("name", 1, _STR, 0, None)])
    msg("AllocateResponse", [("container_responses", 1, _MSG, 1, "ContainerAllocateResponse")])
    msg("ContainerAllocateResponse", [("mounts", 2, _MSG, 1, "Mount"), ("devices", 3, _MSG, 1, "DeviceSpec"),
                                      ("cdi_devices", 5, _MSG, 1, "CDIDevice")],
        maps=[("envs", 1), ("annotations", 4)])
    msg("Mount", [("container_path", 1, _STR, 0, None), ("host_path", 2, _STR, 0, None),
                  ("read_only", 3, _BOOL, 0, None)])
    msg("DeviceSpec", [("container_path", 1, _STR, 0, None), ("host_path", 2, _STR, 0, None),
                       ("permissions", 3, _STR, 0, None)])
    pool = descriptor_pool.DescriptorPool()
    fd = pool.Add(fdp)
    del fd
    return {m.name: message_factory.GetMessageClass(pool.FindMessageTypeByName("v1beta1." + m.name))
            for m in fdp.message_type}


_M = _build()
DevicePluginOptions = _M["DevicePluginOptions"]
RegisterRequest = _M["RegisterRequest"]
Empty = _M["Empty"]
ListAndWatchResponse = _M["ListAndWatchResponse"]
TopologyInfo = _M["TopologyInfo"]
NUMANode = _M["NUMANode"]
Device = _M["Device"]
PreStartContainerRequest = _M["PreStartContainerRequest"]
PreStartContainerResponse = _M["PreStartContainerResponse"]
PreferredAllocationRequest = _M["PreferredAllocationRequest"]
ContainerPreferredAllocationRequest = _M["ContainerPreferredAllocationRequest"]
PreferredAllocationResponse = _M["PreferredAllocationResponse"]
ContainerPreferredAllocationResponse = _M["ContainerPreferredAllocationResponse"]
AllocateRequest = _M["AllocateRequest"]
ContainerAllocateRequest = _M["ContainerAllocateRequest"]
CDIDevice = _M["CDIDevice"]
AllocateResponse = _M["AllocateResponse"]
ContainerAllocateResponse = _M["ContainerAllocateResponse"]
Mount = _M["Mount"]
DeviceSpec = _M["DeviceSpec"]

# full method names (api.pb.go:1419-1441, 1560-1589)
REGISTER = "/v1beta1.Registration/Register"
GET_OPTIONS = "/v1beta1.DevicePlugin/GetDevicePluginOptions"
LIST_AND_WATCH = "/v1beta1.DevicePlugin/ListAndWatch"
GET_PREFERRED_ALLOCATION = "/v1beta1.DevicePlugin/GetPreferredAllocation"
ALLOCATE = "/v1beta1.DevicePlugin/Allocate"
PRE_START_CONTAINER = "/v1beta1.DevicePlugin/PreStartContainer"


def wrap_bytes_field(field: int, payload: bytes) -> bytes:
    """Length-delimited field: tag, varint length, payload (used to wrap C++-encoded sub-messages)."""
    out = bytearray()
    v = (field << 3) | 2
    for x in (v, len(payload)):
        while x >= 0x80:
            out.append((x & 0x7F) | 0x80)
            x >>= 7
        out.append(x)
    return bytes(out) + payload
