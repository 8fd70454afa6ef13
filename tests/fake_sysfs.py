"""Build a fake "/" around one of the reference's captured kfd topologies so that
GetAMDGPUs-level code (reference: hard-coded /sys paths, amdgpu.go:150-155,221) can run on it.

The reference ships no fixture for /sys/module/amdgpu or /sys/devices/platform, so
those are derived from the topology: render minors are grouped by devID; the lowest minor of
each group becomes the PCI function (driver dir named after its BDF), the others become
amdgpu_xcp_<k> platform devices -- the CPX layout of docs/user-guide/resource-allocation.md.
"""
import os
import shutil

from oracle import amdgpu as oamd


def build(dst, topo_src_nodes, compute="", memory="", numa_of_group=None, drop_numa_for=(), odd_drm=None,
          hetero_second=None):
    """topo_src_nodes: directory holding the node dirs (".../topology/nodes" or ".../nodes")."""
    if os.path.exists(dst):
        shutil.rmtree(dst)
    kfd = os.path.join(dst, "sys/class/kfd/kfd/topology")
    os.makedirs(kfd)
    os.symlink(topo_src_nodes, os.path.join(kfd, "nodes"))
    os.makedirs(os.path.join(dst, "sys/devices/platform"))
    drv = os.path.join(dst, "sys/module/amdgpu/drivers/pci:amdgpu")
    os.makedirs(drv)
    dev_ids = oamd.GetDevIdsFromTopology(os.path.join(dst, "sys/class/kfd/kfd"))
    groups = {}
    for minor, dev in sorted(dev_ids.items()):
        groups.setdefault(dev, []).append(minor)
    xcp = 0
    for gi, (dev, minors) in enumerate(sorted(groups.items())):
        dom, bus, dv, _ = dev.split(":")
        bdf = "%s:%s:%s.0" % (dom, bus, dv)
        numa = numa_of_group(gi) if numa_of_group else (0 if gi < (len(groups) + 1) // 2 else 1)
        for k, minor in enumerate(minors):
            if k == 0:
                base = os.path.join(drv, bdf)
                os.makedirs(base)
                if bdf not in drop_numa_for:
                    with open(os.path.join(base, "numa_node"), "w") as f:
                        f.write("%d\n" % numa)
                c, m = compute, memory
                if hetero_second and gi % 2 == 1:
                    c, m = hetero_second
                if c:
                    with open(os.path.join(base, "current_compute_partition"), "w") as f:
                        f.write(c.upper() + "\n")
                    with open(os.path.join(base, "available_compute_partition"), "w") as f:
                        f.write("SPX, DPX, QPX, CPX\n")
                if m:
                    with open(os.path.join(base, "current_memory_partition"), "w") as f:
                        f.write(m.upper() + "\n")
            else:
                xcp += 1
                base = os.path.join(dst, "sys/devices/platform", "amdgpu_xcp_%d" % (gi * 8 + k))
                os.makedirs(base)
            os.makedirs(os.path.join(base, "drm", "card%d" % (minor - 127)))
            os.makedirs(os.path.join(base, "drm", "renderD%d" % minor))
            cls = os.path.join(dst, "sys/class/drm/card%d/device" % (minor - 127))
            os.makedirs(os.path.join(cls, "driver/module"))
            for name, text in (("device", "0x74a1\n"), ("product_name", "AMD Instinct MI300X (test)\n"),
                               ("driver/module/version", "6.8.5\n"), ("driver/module/srcversion", "ABCDEF0123\n")):
                with open(os.path.join(cls, name), "w") as f:
                    f.write(text)
    if odd_drm:   # e.g. a platform dir whose render minor kfd does not know (amdgpu.go:258-260)
        base = os.path.join(dst, "sys/devices/platform", odd_drm[0])
        os.makedirs(os.path.join(base, "drm", "card%d" % odd_drm[1]))
        os.makedirs(os.path.join(base, "drm", "renderD%d" % odd_drm[2]))
    return dst
