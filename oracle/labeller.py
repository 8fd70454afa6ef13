"""Oracle restatement of cmd/k8s-node-labeller/main.go:37-397 (test infrastructure only).

libdrm-backed values (family, firmware, marketing name; amdgpu.go:84-99,392-437,540-552)
cannot be restated -- they are ioctl results -- so the generators that need them take a
`drm` provider: {"card<N>": {"family": str, "feat": {blk: u32}, "fw": {blk: u32},
"marketing_name": str}}.  Parity for those three generators is unpinned (the reference
only exercises them on live hardware)."""
import os

from . import amdgpu, gosem

experimentalAMDPrefix = "beta.amd.com"   # main.go:38
amdPrefix = "amd.com"                    # main.go:39

GENERATOR_NAMES = ["firmware", "family", "driver-version", "driver-src-version", "device-id",
                   "product-name", "vram", "simd-count", "cu-count", "compute-memory-partition",
                   "compute-partitioning-supported", "memory-partitioning-supported"]   # main.go:115-379

reSizeInBytes = gosem.compile_re2(r"size_in_bytes\s(\d+)")       # main.go:110
reSimdCount = gosem.compile_re2(r"simd_count\s(\d+)")            # main.go:111
reSimdPerCu = gosem.compile_re2(r"simd_per_cu\s(\d+)")           # main.go:112
reDrmRenderMinor = gosem.compile_re2(r"drm_render_minor\s(\d+)")  # main.go:113


def createLabelPrefix(name, experimental):
    """main.go:76-85."""
    return "%s/gpu.%s" % (experimentalAMDPrefix if experimental else amdPrefix, name)


def initLabelLists():
    """main.go:46-53 -> (allLabelKeys, allExperimentalLabelKeys), sorted."""
    return (sorted(createLabelPrefix(n, False) for n in GENERATOR_NAMES),
            sorted(createLabelPrefix(n, True) for n in GENERATOR_NAMES))


def removeOldNodeLabels(labels):
    """main.go:55-74 on a plain dict (node.Labels); mutates and returns it."""
    if labels is None:
        return labels
    all_keys, all_exp = initLabelLists()
    for k in all_keys:
        labels.pop(k, None)
    for k in all_exp:
        if k in labels:
            val = labels.pop(k)
            labels.pop("%s.%s" % (k, val), None)
    return labels


def createLabels(kind, entries):
    """main.go:87-108."""
    labels = {}
    prefix = createLabelPrefix(kind, True)
    for k, v in entries.items():
        labels["%s.%s" % (prefix, k)] = str(v)
        if len(entries) == 1:
            labels[prefix] = k
    prefix = createLabelPrefix(kind, False)
    for k, v in entries.items():
        if len(entries) == 1:
            labels[prefix] = k
        else:
            labels["%s.%s" % (prefix, k)] = str(v)
    return labels


def go_round_half_away(x: float) -> int:
    """math.Round: half away from zero."""
    import math
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def vram_label_value(size_in_bytes: int) -> str:
    """main.go:268-272: int(math.Round(float64(size/1MiB)/1024)) + "G"."""
    tmp = size_in_bytes // (1024 * 1024)
    return "%dG" % go_round_half_away(float(tmp) / 1024)


def _under(sys_root, path):
    return os.path.join(sys_root, path.lstrip("/")) if sys_root else path


def _read_first(gpus, sys_root, fmt):
    version = ""
    for _, v in sorted(gpus.items()):
        p = _under(sys_root, fmt % v["card"])
        try:
            with open(p, "rb") as f:
                version = f.read().decode("utf-8", "replace").strip()
        except OSError:
            continue
        break
    return version


def _node_files(sys_root):
    return gosem.glob(_under(sys_root, "/sys/class/kfd/kfd/topology/nodes/*/properties"))


def generate(name, gpus, sys_root="", drm=None):
    """One label generator (main.go:115-379).  Go ranges over `gpus` in random order; where
    that matters (driver-version picks the first readable card) canonical order = sorted id."""
    drm = drm or {}
    if name == "firmware":                                   # main.go:116-144
        counts = {}
        for _, v in sorted(gpus.items()):
            d = drm.get("card%d" % v["card"])
            if d is None:
                continue
            for fw, ver in d["feat"].items():
                k = "%s.feat.%s" % (fw, ver)          # %d in Go (uint32); %s here so an exported tree may carry NVML's dotted versions
                counts[k] = counts.get(k, 0) + 1
            for fw, ver in d["fw"].items():
                k = "%s.fw.%s" % (fw, ver)
                counts[k] = counts.get(k, 0) + 1
        pfx = createLabelPrefix("firmware", True)
        return {"%s.%s" % (pfx, k): str(v) for k, v in counts.items()}
    if name == "family":                                     # main.go:145-158
        counts = {}
        for _, v in sorted(gpus.items()):
            d = drm.get("card%d" % v["card"])
            if d is None or not d.get("family"):
                continue
            counts[d["family"]] = counts.get(d["family"], 0) + 1
        return createLabels("family", counts)
    if name == "driver-version":                             # main.go:159-174
        return {createLabelPrefix(name, False):
                _read_first(gpus, sys_root, "/sys/class/drm/card%d/device/driver/module/version")}
    if name == "driver-src-version":                         # main.go:175-190
        return {createLabelPrefix(name, False):
                _read_first(gpus, sys_root, "/sys/class/drm/card%d/device/driver/module/srcversion")}
    if name == "device-id":                                  # main.go:191-209
        counts = {}
        for _, v in sorted(gpus.items()):
            p = _under(sys_root, "/sys/class/drm/card%d/device/device" % v["card"])
            try:
                with open(p, "rb") as f:
                    devid = f.read().decode("utf-8", "replace").strip()
            except OSError:
                continue
            if len(devid) < 2:
                raise gosem.GoPanic("slice bounds out of range devid[0:2]")   # main.go:200
            if devid[0:2] == "0x":
                devid = devid[2:]
            counts[devid] = counts.get(devid, 0) + 1
        return createLabels("device-id", counts)
    if name == "product-name":                               # main.go:210-238
        counts = {}

        def repl(s):
            return s.replace(" ", "_").replace("(", "").replace(")", "")
        for _, v in sorted(gpus.items()):
            p = _under(sys_root, "/sys/class/drm/card%d/device/product_name" % v["card"])
            try:
                with open(p, "rb") as f:
                    b = f.read().decode("utf-8", "replace")
            except OSError:
                b = ""
            prod = repl(b.strip())
            if prod == "":
                d = drm.get("card%d" % v["card"])
                if d is not None and d.get("marketing_name") is not None:
                    prod = repl(d["marketing_name"].strip())
            if prod == "":
                continue
            counts[prod] = counts.get(prod, 0) + 1
        return createLabels("product-name", counts)
    if name in ("vram", "simd-count", "cu-count"):           # main.go:239-354
        files = _node_files(sys_root)
        if not files:
            return {}
        counts = {}
        for _, gpu in sorted(gpus.items()):
            for file in files:
                render_minor, _ = amdgpu.ParseTopologyProperties(file, reDrmRenderMinor)
                if int(render_minor) != gpu["renderD"]:
                    continue
                if name == "vram":
                    node_number = file.split("/")[-2]
                    vpath = _under(sys_root, "/sys/class/kfd/kfd/topology/nodes/%s/mem_banks/0/properties" % node_number)
                    vsize, err = amdgpu.ParseTopologyProperties(vpath, reSizeInBytes)
                    if err is not None:
                        continue
                    key = vram_label_value(vsize)
                elif name == "simd-count":
                    s, e = amdgpu.ParseTopologyProperties(file, reSimdCount)
                    if e is not None:
                        continue
                    key = "%d" % s
                else:
                    s, e = amdgpu.ParseTopologyProperties(file, reSimdCount)
                    if e is not None:
                        continue
                    c, e = amdgpu.ParseTopologyProperties(file, reSimdPerCu)
                    if e is not None or c == 0:
                        continue
                    key = "%d" % _go_div(s, c)
                counts[key] = counts.get(key, 0) + 1
                break
        return createLabels(name, counts)
    if name == "compute-memory-partition":                   # main.go:355-368
        counts = amdgpu.UniquePartitionConfigCount(gpus)
        if len(amdgpu.UniquePartitionConfigCount(amdgpu.GetAMDGPUs(sys_root))) <= 1:
            for pt in sorted(counts):
                if counts[pt] > 0:
                    return {createLabelPrefix(name, False): pt}
        return {}
    if name == "compute-partitioning-supported":             # main.go:369-373
        return {createLabelPrefix(name, False): "true" if amdgpu.IsComputePartitionSupported(sys_root) else "false"}
    if name == "memory-partitioning-supported":              # main.go:374-378
        return {createLabelPrefix(name, False): "true" if amdgpu.IsMemoryPartitionSupported(sys_root) else "false"}
    raise KeyError(name)


def _go_div(a, b):
    """Go integer division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def generateLabels(enabled, sys_root="", drm=None):
    """main.go:383-397: union of the enabled generators' labels."""
    results = {}
    gpus = amdgpu.GetAMDGPUs(sys_root)
    for name in GENERATOR_NAMES:
        if not enabled.get(name, False):
            continue
        results.update(generate(name, gpus, sys_root, drm))
    return results
