"""Oracle restatement of internal/pkg/plugin/plugin.go, internal/pkg/exporter/health.go
and cmd/k8s-device-plugin/main.go:42-91 (test infrastructure only)."""

from . import amdgpu, gosem
from .allocator import Device


Healthy = "Healthy"        # vendor/k8s.io/kubelet/pkg/apis/deviceplugin/v1beta1/constants.go:21
Unhealthy = "Unhealthy"    # constants.go:23

topoSIMDre = gosem.compile_re2(r"simd_count\s(\d+)")     # plugin.go:121


def getDevices(sys_root=""):
    """plugin.go:93-111 (canonical order: sorted by id; Go's map order is random)."""
    devs = amdgpu.GetAMDGPUs(sys_root)
    return [Device(Id=k, Card=v["card"], RenderD=v["renderD"], DevId=v["devID"],
                   ComputePartitionType=v["computePartitionType"],
                   MemoryPartitionType=v["memoryPartitionType"], NodeId=v["nodeId"],
                   NumaNode=v["numaNode"]) for k, v in sorted(devs.items())]


def countGPUDevFromTopology(topo_root=amdgpu.KFD_ROOT):
    """plugin.go:123-159."""
    count = 0
    for node_file in gosem.glob(topo_root + "/topology/nodes/*/properties"):
        try:
            with open(node_file, "rb") as f:
                data = f.read()
        except OSError:
            continue
        for line in gosem.scanner_lines(data):
            m = topoSIMDre.search(line)
            if m is None:
                continue
            if gosem.atoi_ignore_err(m.group(1)) > 0:
                count += 1
                break
    return count


def simpleHealthCheck(topo_root=amdgpu.KFD_ROOT):
    """plugin.go:161-206 with the kfd root made injectable: true iff some node has
    cpu_cores_count == 0 (or absent) and gfx_target_version > 0; first hit wins."""
    for prop_file in gosem.glob(topo_root + "/topology/nodes/*/properties"):
        try:
            with open(prop_file, "rb") as f:
                data = f.read()
        except OSError:
            continue
        cpuCores, gfxVersion = 0, 0
        lines, scan_err = gosem.scan(data)
        for line in lines:
            if line.startswith(b"cpu_cores_count"):
                parts = gosem.fields(line)
                if len(parts) == 2:
                    cpuCores = gosem.atoi_ignore_err(parts[1])
            elif line.startswith(b"gfx_target_version"):
                parts = gosem.fields(line)
                if len(parts) == 2:
                    gfxVersion = gosem.atoi_ignore_err(parts[1])
        if scan_err:        # plugin.go:193-196: scanner.Err() != nil => skip this file
            continue
        if cpuCores == 0 and gfxVersion > 0:
            return True
    return False


def merge_health(dev_ids, default_health, exporter_map):
    """exporter/health.go:86-106.  `exporter_map` is None when getGPUHealth() returned an
    error (socket absent / RPC failed), else {Device: "Healthy"|"Unhealthy"}."""
    out = []
    for d in dev_ids:
        if exporter_map is None:
            out.append(default_health)
        elif d in exporter_map:
            out.append(exporter_map[d])
        else:
            out.append(default_health)
    return out


def exporter_states_to_map(gpu_states):
    """exporter/health.go:74-80: [(Device, Health)] -> map; only the exact lower-case
    string "healthy" counts as healthy."""
    return {dev: (Healthy if health == Healthy.lower() else Unhealthy) for dev, health in gpu_states}


def list_and_watch_devices(gpus, resource):
    """plugin.go:235-299: the device list ListAndWatch sends first.  Returns
    (isHomogeneous, [(ID, Health, numa)]) in canonical (sorted-by-ID) order, or None when a
    heterogeneous node has no devices for `resource` (nothing is sent)."""
    homogeneous = len(amdgpu.UniquePartitionConfigCount(gpus)) <= 1
    if homogeneous:
        return True, [(k, Healthy, int(v["numaNode"])) for k, v in sorted(gpus.items())]
    by_type = {}
    for k, v in sorted(gpus.items()):
        pt = v["computePartitionType"] + "_" + v["memoryPartitionType"]
        by_type.setdefault(pt, []).append((k, Healthy, int(v["numaNode"])))
    if resource in by_type:
        return False, by_type[resource]
    return False, None


def allocate_device_specs(gpus, device_ids):
    """plugin.go:356-393 for one container request: [(host_path, container_path, perms)].
    /dev/kfd first, then card/renderD paths per requested id (Go emits the two in random
    map order; canonical here: card then renderD).  Unknown ids add nothing."""
    specs = [("/dev/kfd", "/dev/kfd", "rw")]
    for id_ in device_ids:
        g = gpus.get(id_)
        if g is None:
            continue
        for k in ("card", "renderD"):
            p = "/dev/dri/%s%d" % (k, g[k])
            specs.append((p, p, "rw"))
    return specs


# ---- cmd/k8s-device-plugin/main.go ---------------------------------------------------
class StrategyError(Exception):
    pass


def ParseStrategy(s):
    """main.go:42-51."""
    if s in ("single", "mixed"):
        return s
    raise StrategyError("invalid resource naming strategy: %s" % s)


def getResourceList(strategy, sys_root=""):
    """main.go:53-91.  Returns (resources, err); resources sorted (Go map order is random)."""
    gpus = amdgpu.GetAMDGPUs(sys_root)
    counts = amdgpu.UniquePartitionConfigCount(gpus)
    homogeneous = len(counts) <= 1
    if len(gpus) == 0:
        return [], None
    if homogeneous:
        if strategy == "single":
            return ["gpu"], None
        if len(counts) == 0:
            return ["gpu"], None
        return sorted(k for k, c in counts.items() if c > 0), None
    if strategy == "single":
        return [], StrategyError("Partitions of different styles across GPUs in a node is not supported with "
                                 "single strategy. Please start device plugin with mixed strategy")
    return sorted(k for k, c in counts.items() if c > 0), None
