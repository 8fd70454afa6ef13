"""Oracle restatement of internal/pkg/allocator/{device.go,besteffort_policy.go}
(test infrastructure only)."""
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from . import gosem
from .gosem import GoPanic, ParseError

topoRootPath = "/sys/class/kfd/kfd/topology/nodes"   # device.go:34

# device.go:38-54
sameDevIdWeight = 10
xgmiLinkWeight = 10
sameNumaNodeWeight = 10
differentDevIdWeight = 20
differentNumaNodeWeight = 20
pcieLinkWeight = 40
otherLinkWeight = 50

# besteffort_policy.go:36-43
invalidSize = "allocation size can not be negative"
invalidAvailable = "available devices count less than allocation size"
invalidRequired = "must_include devices size is more than allocation size"
invalidReqAvailable = "must_include length should be less than or equal to avilable device size"
invalidInit = "Init method must be called before Allocate"
noCandidateFound = "No candidate subset found with matching criteria"

MAX_INT32 = (1 << 31) - 1


class AllocError(Exception):
    """A Go `error` returned by the allocator; str(e) is the reference message."""


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


@dataclass
class DeviceSet:                    # device.go:67-73
    Ids: List[int]
    TotalWeight: int
    LastIdx: int
    Size: int
    ParentIds: List[int]


@dataclass
class DevicePartitions:             # device.go:75-80
    ParentId: str = ""
    DevId: str = ""
    Ids: List[int] = field(default_factory=list)
    Devs: List[str] = field(default_factory=list)


def setContainsAll(s, subset):
    """device.go:88-105."""
    if len(subset) > len(s):
        return False
    return all(any(x == d for x in s) for d in subset)


def fetchTopoProperties(path, regexes):
    """device.go:107-133: every line is tried against every regex; the LAST match per
    regex wins; ParseInt(base 0, 32 bits) failure aborts the file."""
    try:
        with open(path, "rb") as f:
            data = f.read()
    except OSError as e:
        return [0], e
    res = [0] * len(regexes)
    for line in gosem.scanner_lines(data):
        for idx, rx in enumerate(regexes):
            m = rx.search(line)
            if m is None:
                continue
            try:
                res[idx] = gosem.parse_int(m.group(1), 0, 32)
            except ParseError as pe:
                return None, pe
    return res, None


def calculatePairWeight(frm: Device, to: Device, linkType: int) -> int:
    """device.go:135-157."""
    w = sameDevIdWeight if frm.DevId == to.DevId else differentDevIdWeight
    if linkType == 11:
        w += xgmiLinkWeight
    elif linkType == 2:
        w += pcieLinkWeight
    else:
        w += otherLinkWeight
    w += sameNumaNodeWeight if frm.NumaNode == to.NumaNode else differentNumaNodeWeight
    return w


_linkRes = [gosem.compile_re2(r"node_from\s(\d+)"), gosem.compile_re2(r"node_to\s(\d+)"),
            gosem.compile_re2(r"type\s(\d+)")]                     # device.go:169-173
_drmRenderMinor = [gosem.compile_re2(r"drm_render_minor\s(\d+)")]  # device.go:237


def scanAndPopulatePeerWeights(fromPath, devices, lookupNodes, p2pWeights):
    """device.go:159-218."""
    paths = gosem.glob(os.path.join(fromPath, "io_links", "[0-9]*"))
    p2pPaths = gosem.glob(os.path.join(fromPath, "p2p_links", "[0-9]*"))
    paths = paths + p2pPaths
    for topath in paths:
        vals, err = fetchTopoProperties(os.path.join(topath, "properties"), _linkRes)
        if err is not None:
            continue
        if vals[0] < vals[1]:
            frm, to = vals[0], vals[1]
        else:
            frm, to = vals[1], vals[0]
        if frm not in lookupNodes or to not in lookupNodes:
            continue
        fromDev = toDev = None
        found = False
        for d in devices:                       # device.go:198-209 (linear search, last hit kept)
            if d.NodeId == frm:
                fromDev = d
            if d.NodeId == to:
                toDev = d
            if fromDev is not None and toDev is not None:
                found = True
                break
        if found:
            p2pWeights.setdefault(frm, {})[to] = calculatePairWeight(fromDev, toDev, int(vals[2]))
    return None


def fetchAllPairWeights(devices, p2pWeights, folderPath):
    """device.go:220-252.  Returns a Go-style error (Exception) or None."""
    if len(devices) == 0:
        return AllocError("Devices list is empty. Unable to calculate pair wise weights")
    if folderPath == "":
        folderPath = topoRootPath
    paths = gosem.glob(os.path.join(folderPath, "[0-9]*"))
    nodeIds = {d.NodeId for d in devices}
    for path in paths:
        vals, err = fetchTopoProperties(os.path.join(path, "properties"), _drmRenderMinor)
        if err is not None or vals[0] <= 0:
            continue
        scanAndPopulatePeerWeights(path, devices, nodeIds, p2pWeights)
    return None


def NewDeviceSet(nodeIds, parentIds, weight, lastIdx):
    return DeviceSet(Ids=nodeIds, TotalWeight=weight, LastIdx=lastIdx, Size=len(nodeIds), ParentIds=parentIds)


def addDeviceToSubsetAndUpdateWeight(subset, devId, devIdx, p2pWeights):
    """device.go:254-273: missing p2pWeights entries read as 0."""
    w = subset.TotalWeight
    for d in subset.Ids:
        frm, to = (d, devId) if d < devId else (devId, d)
        w += p2pWeights.get(frm, {}).get(to, 0)
    return NewDeviceSet(list(subset.Ids) + [devId], subset.ParentIds, w, devIdx)


def groupPartitionsByDevId(devs):
    """device.go:287-304."""
    partitions: Dict[str, DevicePartitions] = {}
    for dev in devs:
        if dev.DevId not in partitions:
            partitions[dev.DevId] = DevicePartitions(DevId=dev.DevId)
        if "amdgpu_xcp" not in dev.Id:
            partitions[dev.DevId].ParentId = dev.Id
        partitions[dev.DevId].Ids.append(dev.NodeId)
        partitions[dev.DevId].Devs.append(dev.Id)
    return partitions


def filterPartitions(partitions, available, required):
    """device.go:310-351.  Go iterates the map in random order and then sort.Slice()s by
    (len, ParentId); with unique ParentIds the result is deterministic.  Ties (equal len and
    equal ParentId, only possible with ParentId == "") are broken here by DevId."""
    availableIds = {a.NodeId for a in available}
    requiredIds = {r.NodeId for r in required}
    outset = []
    for key in sorted(partitions):
        ps = partitions[key]
        filtered = [i for i in ps.Ids if i not in requiredIds and i in availableIds]
        if filtered:
            outset.append(DevicePartitions(DevId=ps.DevId, Ids=sorted(filtered), ParentId=ps.ParentId))
    outset.sort(key=lambda p: (len(p.Ids), p.ParentId, p.DevId))
    return outset


def getCandidateDeviceSubsets(allDevPartitions, total, available, required, size, p2pWeights):
    """device.go:353-442.  Returns (subsets, err)."""
    if size <= 0:
        return [], AllocError("subset size should be positive integer")
    if len(available) < size:
        return [], AllocError("subset size is more than available devices")
    if any(a is None for a in available):
        raise GoPanic("nil *Device in available (unknown device id)")   # device.go:362-364
    available.sort(key=lambda d: d.NodeId)
    devPartitions = filterPartitions(allDevPartitions, available, required)
    newSize = size - len(required)
    subsetsTemp: List[DeviceSet] = []
    subsetsFinal: List[DeviceSet] = []
    for idx, partition in enumerate(devPartitions):
        devset = NewDeviceSet([partition.Ids[0]], [idx], 0, idx)
        if newSize == 1:
            for req in required:
                devset = addDeviceToSubsetAndUpdateWeight(devset, req.NodeId, idx, p2pWeights)
            subsetsFinal.append(devset)
            continue
        sizeFulfilled = False
        for i in range(1, len(partition.Ids)):
            devset = addDeviceToSubsetAndUpdateWeight(devset, partition.Ids[i], idx, p2pWeights)
            if i == newSize - 1:
                sizeFulfilled = True
                break
        if sizeFulfilled:
            for req in required:
                devset = addDeviceToSubsetAndUpdateWeight(devset, req.NodeId, idx, p2pWeights)
            subsetsFinal.append(devset)
        else:
            subsetsTemp.append(devset)
    head = 0
    while head < len(subsetsTemp):
        current = subsetsTemp[head]
        head += 1
        if len(current.ParentIds) == len(devPartitions):
            continue
        for idx in range(len(devPartitions)):
            if idx in current.ParentIds:
                continue
            parentIds = list(current.ParentIds) + [idx]
            devset = NewDeviceSet(current.Ids, parentIds, current.TotalWeight, current.LastIdx)
            for id_ in devPartitions[idx].Ids:
                devset = addDeviceToSubsetAndUpdateWeight(devset, id_, idx, p2pWeights)
                if devset.Size == newSize:
                    for req in required:
                        devset = addDeviceToSubsetAndUpdateWeight(devset, req.NodeId, idx, p2pWeights)
                    subsetsFinal.append(devset)
                    break
            if devset.Size < newSize:
                subsetsTemp.append(devset)
    return subsetsFinal, None


class BestEffortPolicy:
    """besteffort_policy.go:45-151."""

    def __init__(self):
        self.devices: List[Device] = []
        self.devicesMap: Dict[str, Device] = {}
        self.devicePartitions: Dict[str, DevicePartitions] = {}
        self.p2pWeights: Dict[int, Dict[int, int]] = {}
        self.last_score: Optional[int] = None
        self.last_candidates: Optional[int] = None

    def Init(self, devs, topoDir):
        err = fetchAllPairWeights(devs, self.p2pWeights, topoDir)
        if len(self.p2pWeights) == 0:
            return AllocError("Besteffort Policy init failed to initialize p2pWeights")
        if err is None:
            self.devices = devs
            for d in devs:
                self.devicesMap[d.Id] = d
            self.devicePartitions = groupPartitionsByDevId(devs)
        return err

    def getDevicesFromIds(self, ids):
        return [self.devicesMap.get(i) for i in ids]

    def Allocate(self, availableIds, requiredIds, size):
        """Returns (ids, err) exactly like besteffort_policy.go:88-151."""
        requiredIds = requiredIds or []
        if size <= 0:
            return [], AllocError(invalidSize)
        if len(availableIds) < size:
            return [], AllocError(invalidAvailable)
        if len(requiredIds) > size:
            return [], AllocError(invalidRequired)
        if len(requiredIds) > len(availableIds):
            return [], AllocError(invalidReqAvailable)
        if len(self.devices) == 0:
            return [], AllocError(invalidInit)
        if len(availableIds) == size:
            return availableIds, None
        if len(requiredIds) == size:
            return requiredIds, None
        if len(self.p2pWeights) == 0:
            return [], AllocError(invalidInit)
        if not setContainsAll(availableIds, requiredIds):
            return [], AllocError(noCandidateFound)
        available = self.getDevicesFromIds(availableIds)
        required = self.getDevicesFromIds(requiredIds)
        allSubsets, err = getCandidateDeviceSubsets(self.devicePartitions, self.devices, available, required,
                                                    size, self.p2pWeights)
        if err is not None:
            return [], err
        bestScore = MAX_INT32
        candidate = None
        for s in allSubsets:
            if s.TotalWeight < bestScore:
                candidate, bestScore = s, s.TotalWeight
        if candidate is None:
            raise GoPanic("nil candidate DeviceSet")                     # besteffort_policy.go:141
        outset = []
        for id_ in candidate.Ids:
            for d in available:
                if d.NodeId == id_:
                    outset.append(d.Id)
                    break
        self.last_score, self.last_candidates = candidate.TotalWeight, len(allSubsets)
        return outset, None


def getTestDevices(devCount, partitionCountPerDev, numanodeCount, startNodeId, endNodeId):
    """The reference tests' synthetic-device convention (device_test.go:43-67)."""
    res = []
    nodeId = startNodeId
    for i in range(devCount):
        numa = devCount // numanodeCount
        for j in range(partitionCountPerDev):
            id_ = "amdgpu_xcp_%d" % (i * 8 + j)
            if j == 0:
                id_ = "test%d" % (i + 1)
            if nodeId > endNodeId:
                break
            res.append(Device(Id=id_, NodeId=nodeId, NumaNode=i // numa, DevId=str(i)))
            nodeId += 1
    return res
