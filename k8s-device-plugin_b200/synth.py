"""Synthetic kfd-shaped sysfs trees (fixture generator; not on the hot path).

Writes what the reference's walk reads (amdgpu.go:149-268, plugin.go:161-206,
allocator/device.go:159-252): `sys/module/amdgpu/drivers/pci:amdgpu/<bdf>/{numa_node,
current_*_partition, drm/}`, `sys/devices/platform/amdgpu_xcp_*`, and
`sys/class/kfd/kfd/topology/nodes/<k>/{properties, mem_banks/0/properties, io_links/}`.
Used to give the reference algorithm (the oracle) and the `kfd:` backend the *same* N-device
node shape as an N x B200 NVSwitch box: N GPUs, every pair linked with type 11, GPUs split
evenly over two NUMA nodes, optional `partitions` > 1 for the MIG/CPX-style layout.
"""
import os


def _w(path, text):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        f.write(text)


def link_props(ltype, frm, to, weight):
    return ("type %d\nversion_major 0\nversion_minor 0\nnode_from %d\nnode_to %d\nweight %d\nmin_latency 0\n"
            "max_latency 0\nmin_bandwidth 0\nmax_bandwidth 0\nrecommended_transfer_size 0\nflags 1\n"
            % (ltype, frm, to, weight))


def write_b200_tree(root, n_gpus=8, partitions=1, n_cpu_nodes=2, vram_bytes=192265846784, sm_count=148,
                    compute_partition="", memory_partition="", link_type=11):
    """Returns the list of device ids the tree describes (sorted)."""
    ids = []
    nodes = os.path.join(root, "sys/class/kfd/kfd/topology/nodes")
    os.makedirs(os.path.join(root, "sys/devices/platform"), exist_ok=True)
    for k in range(n_cpu_nodes):
        _w(os.path.join(nodes, str(k), "properties"),
           "cpu_cores_count 64\nsimd_count 0\nmem_banks_count 1\ncaches_count 0\nio_links_count 0\n"
           "cpu_core_id_base %d\nsimd_id_base 0\nvendor_id 0\ndevice_id 0\nlocation_id 0\ndomain 0\n"
           "drm_render_minor 0\n" % (k * 64))
    gpu_nodes = []   # (node_id, gpu_index)
    node_id = n_cpu_nodes
    minor = 128
    card = 0
    for g in range(n_gpus):
        bus = 0x19 + 0x10 * g
        bdf = "0000:%02x:00.0" % bus
        numa = 0 if g < (n_gpus + 1) // 2 else 1
        pci = os.path.join(root, "sys/module/amdgpu/drivers/pci:amdgpu", bdf)
        _w(os.path.join(pci, "numa_node"), "%d\n" % numa)
        if compute_partition:
            _w(os.path.join(pci, "current_compute_partition"), compute_partition.upper() + "\n")
            _w(os.path.join(pci, "available_compute_partition"), "SPX, " + compute_partition.upper() + "\n")
        if memory_partition:
            _w(os.path.join(pci, "current_memory_partition"), memory_partition.upper() + "\n")
            _w(os.path.join(pci, "available_memory_partition"), memory_partition.upper() + "\n")
        for p in range(partitions):
            if p == 0:
                base = pci
                ids.append(bdf)
            else:
                name = "amdgpu_xcp_%d" % (g * 8 + p)
                base = os.path.join(root, "sys/devices/platform", name)
                ids.append(name)
            os.makedirs(os.path.join(base, "drm", "card%d" % card), exist_ok=True)
            os.makedirs(os.path.join(base, "drm", "renderD%d" % minor), exist_ok=True)
            drm = os.path.join(root, "sys/class/drm/card%d/device" % card)
            _w(os.path.join(drm, "device"), "0x2901\n")
            _w(os.path.join(drm, "product_name"), "NVIDIA B200\n")
            _w(os.path.join(drm, "driver/module/version"), "580.159.03\n")
            _w(os.path.join(drm, "driver/module/srcversion"), "SYNTHETIC0000000000000000\n")
            nd = os.path.join(nodes, str(node_id))
            _w(os.path.join(nd, "properties"),
               "cpu_cores_count 0\nsimd_count %d\nmem_banks_count 1\ncaches_count 0\nio_links_count %d\n"
               "cpu_core_id_base 0\nsimd_id_base 0\nmax_waves_per_simd 16\nwave_front_size 32\nsimd_per_cu 4\n"
               "gfx_target_version 100000\nvendor_id 4318\ndevice_id 10497\nlocation_id %d\ndomain 0\n"
               "drm_render_minor %d\nlocal_mem_size %d\n"
               % (sm_count * 4 // partitions, n_gpus * partitions - 1, bus << 8, minor, vram_bytes // partitions))
            _w(os.path.join(nd, "mem_banks/0/properties"),
               "heap_type 1\nsize_in_bytes %d\nflags 0\nwidth 8192\nmem_clk_max 3996\n" % (vram_bytes // partitions))
            gpu_nodes.append((node_id, g))
            node_id += 1
            minor += 1
            card += 1
    for nid, g in gpu_nodes:
        li = 0
        for other, og in gpu_nodes:
            if other == nid:
                continue
            _w(os.path.join(nodes, str(nid), "io_links", str(li), "properties"),
               link_props(link_type, nid, other, 13 if og == g else 15))
            li += 1
    return sorted(ids)
