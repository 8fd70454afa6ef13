"""Oracle restatement of internal/pkg/amdgpu/amdgpu.go (test infrastructure only).

Every root the reference hard-codes under / is injectable through `sys_root`
(the directory that plays the role of "/"), so fixture trees can stand in for
live sysfs.  `topo_root` mirrors the reference's own variadic `topoRootParam`.
"""
import os

from . import gosem
from .gosem import GoFatal, GoPanic, ParseError

# amdgpu.go:492-494
topoDrmRenderMinorRe = gosem.compile_re2(r"drm_render_minor\s(\d+)")
topoLocationIdRe = gosem.compile_re2(r"location_id\s(\d+)")
topoDomainRe = gosem.compile_re2(r"domain\s(\d+)")
# amdgpu.go:465
fwVersionRe = gosem.compile_re2(r"(\w+) feature version: (\d+), firmware version: (0x[0-9a-fA-F]+)")

KFD_ROOT = "/sys/class/kfd/kfd"                      # amdgpu.go:102,497
DRIVER_DIR = "/sys/module/amdgpu/drivers/"           # amdgpu.go:150
PCI_GLOB = "/sys/module/amdgpu/drivers/pci:amdgpu/[0-9a-fA-F][0-9a-fA-F][0-9a-fA-F][0-9a-fA-F]:*"  # :155
XCP_GLOB = "/sys/devices/platform/amdgpu_xcp_*"      # amdgpu.go:221


class TopologyPropertyError(Exception):
    pass


def _under(sys_root, path):
    return os.path.join(sys_root, path.lstrip("/")) if sys_root else path


def ParseTopologyProperties(path, regex):
    """amdgpu.go:442-463.  First matching line wins; returns (value, err) like Go:
    err is None on success, an Exception instance otherwise (value is what Go returns)."""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        return 0, e
    err = TopologyPropertyError("Topology property not found.  Regex: " + regex.pattern.decode())
    v = 0
    for line in gosem.scanner_lines(data):
        m = regex.search(line)
        if m is None:
            continue
        try:
            v = gosem.parse_int(m.group(1), 0, 64)
            err = None
        except ParseError as pe:
            v, err = pe.value, pe
        break
    return v, err


def GetDevIdsFromTopology(topo_root=KFD_ROOT):
    """amdgpu.go:101-146 -> {render_minor: "dddd:bb:dd:0"}."""
    out = {}
    for node_file in gosem.glob(topo_root + "/topology/nodes/*/properties"):
        v, e = ParseTopologyProperties(node_file, topoDrmRenderMinorRe)
        if e is not None:
            continue
        if v <= 0:
            continue
        location_id, e = ParseTopologyProperties(node_file, topoLocationIdRe)
        if e is not None:
            continue
        domain, e = ParseTopologyProperties(node_file, topoDomainRe)
        if e is not None:
            continue
        dev = (location_id >> 3) & 0x1F
        bus = (location_id >> 8) & 0xFF
        out[int(v)] = "%04x:%02x:%02x:0" % (domain, bus, dev)
    return out


def GetNodeIdsFromTopology(topo_root=KFD_ROOT):
    """amdgpu.go:496-538 -> {render_minor: node_index}."""
    out = {}
    for node_file in gosem.glob(topo_root + "/topology/nodes/*/properties"):
        v, e = ParseTopologyProperties(node_file, topoDrmRenderMinorRe)
        if e is not None:
            continue
        if v <= 0:
            continue
        node_index = os.path.basename(os.path.dirname(node_file))
        try:
            node_id = gosem.atoi(node_index)
        except ParseError:
            continue
        out[int(v)] = node_id
    return out


def _read_trim_lower(path):
    with open(path, "rb") as f:
        return f.read().decode("utf-8", "replace").strip().lower()


def _drm_children(path):
    return gosem.glob(path + "/drm/*")


def GetAMDGPUs(sys_root=""):
    """amdgpu.go:149-268.  Returns {id: {card, renderD, devID, computePartitionType,
    memoryPartitionType, numaNode, nodeId}}.  `card, renderD, nodeId, devID` are declared
    outside the loops (amdgpu.go:157-159) and therefore carry over between devices."""
    if not os.path.exists(_under(sys_root, DRIVER_DIR)):
        raise GoFatal("amdgpu driver unavailable. exiting with exit code 2.")
    matches = gosem.glob(_under(sys_root, PCI_GLOB))

    devID = ""
    devices = {}
    card, renderD, nodeId = 0, 128, 0
    topo_root = _under(sys_root, KFD_ROOT)
    renderDevIds = GetDevIdsFromTopology(topo_root)
    renderNodeIds = GetNodeIdsFromTopology(topo_root)

    for path in matches:
        compute, memory = "", ""
        try:
            compute = _read_trim_lower(os.path.join(path, "current_compute_partition"))
        except OSError:
            pass
        try:
            memory = _read_trim_lower(os.path.join(path, "current_memory_partition"))
        except OSError:
            pass
        try:
            with open(os.path.join(path, "numa_node"), "rb") as f:
                numa_str = f.read().decode("utf-8", "replace").strip()
        except OSError:
            continue
        try:
            numaNode = gosem.atoi(numa_str)
        except ParseError:
            continue

        for dev_path in _drm_children(path):
            name = os.path.basename(dev_path)
            if len(name) < 4:
                raise GoPanic("slice bounds out of range name[0:4]")      # amdgpu.go:202
            if name[0:4] == "card":
                card = gosem.atoi_ignore_err(name[4:])
            else:
                if len(name) < 7:
                    raise GoPanic("slice bounds out of range name[0:7]")  # amdgpu.go:204
                if name[0:7] == "renderD":
                    renderD = gosem.atoi_ignore_err(name[7:])
                    if renderD in renderDevIds:
                        devID = renderDevIds[renderD]
                    if renderD in renderNodeIds:
                        nodeId = renderNodeIds[renderD]
        devices[os.path.basename(path)] = {
            "card": card, "renderD": renderD, "devID": devID, "computePartitionType": compute,
            "memoryPartitionType": memory, "numaNode": numaNode, "nodeId": nodeId}

    for path in gosem.glob(_under(sys_root, XCP_GLOB)):
        compute, memory = "", ""
        numaNode = -1
        for dev_path in _drm_children(path):
            name = os.path.basename(dev_path)
            if len(name) < 4:
                raise GoPanic("slice bounds out of range name[0:4]")
            if name[0:4] == "card":
                card = gosem.atoi_ignore_err(name[4:])
            else:
                if len(name) < 7:
                    raise GoPanic("slice bounds out of range name[0:7]")
                if name[0:7] == "renderD":
                    renderD = gosem.atoi_ignore_err(name[7:])
                    if renderD in renderDevIds:
                        devID = renderDevIds[renderD]
                    # amdgpu.go:240-249: Go ranges over the map in random order and takes the
                    # first hit; canonical order here = sorted by id.
                    for key in sorted(devices):
                        device = devices[key]
                        if device["devID"] == devID:
                            if device["computePartitionType"] != "" and device["memoryPartitionType"] != "":
                                compute = device["computePartitionType"]
                                memory = device["memoryPartitionType"]
                                numaNode = device["numaNode"]
                                break
                    if renderD in renderNodeIds:
                        nodeId = renderNodeIds[renderD]
        if renderD not in renderDevIds:
            continue
        if numaNode == -1:
            continue
        devices[os.path.basename(path)] = {
            "card": card, "renderD": renderD, "devID": devID, "computePartitionType": compute,
            "memoryPartitionType": memory, "numaNode": numaNode, "nodeId": nodeId}
    return devices


def UniquePartitionConfigCount(devices):
    """amdgpu.go:270-285."""
    out = {}
    for device in devices.values():
        c, m = device["computePartitionType"], device["memoryPartitionType"]
        if c != "" and m != "":
            key = c + "_" + m
            out[key] = out.get(key, 0) + 1
    return out


def IsHomogeneous(sys_root=""):
    """amdgpu.go:287-293 (re-enumerates, like the reference)."""
    return len(UniquePartitionConfigCount(GetAMDGPUs(sys_root))) <= 1


def _partition_supported(sys_root, fname):
    matches = gosem.glob(_under(sys_root, PCI_GLOB))
    if not matches:
        return False
    return os.path.exists(os.path.join(matches[0], fname))


def IsComputePartitionSupported(sys_root=""):
    """amdgpu.go:295-311."""
    return _partition_supported(sys_root, "available_compute_partition")


def IsMemoryPartitionSupported(sys_root=""):
    """amdgpu.go:313-328."""
    return _partition_supported(sys_root, "available_memory_partition")


def parseDebugFSFirmwareInfo(path):
    """amdgpu.go:467-490: ParseInt(..., 0, 32) errors are ignored and the (clamped)
    value is cast to uint32."""
    feat, fw = {}, {}
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError:
        return feat, fw
    for line in gosem.scanner_lines(data):
        m = fwVersionRe.search(line)
        if m is None:
            continue
        name = m.group(1).decode()

        def p32(tok):
            try:
                return gosem.parse_int(tok, 0, 32)
            except ParseError as pe:
                return pe.value
        feat[name] = gosem.to_uint32(p32(m.group(2)))
        fw[name] = gosem.to_uint32(p32(m.group(3)))
    return feat, fw
