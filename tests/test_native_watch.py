"""The library-owned ListAndWatch loop (b2dp_watch_*): plugin.go:229-330 + the -pulse ticker of
cmd/k8s-device-plugin/main.go:129-137 as a native thread.  kfd: backend, CPU only."""
import queue
import time

import fake_sysfs
from oracle import amdgpu as oamd
from oracle import plugin as oplug
from test_oracle_golden import topo_dir


def test_watch_initial_beats_and_stop(pkg, kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    V = pkg.v1beta1
    _, want = oplug.list_and_watch_devices(oamd.GetAMDGPUs(root), "gpu")
    got = queue.Queue()
    with pkg.Context("kfd:" + root) as ctx:
        w = ctx.watch(lambda rc, wire, st: got.put((rc, wire, st)), resource="gpu", flags=pkg._native.LW_NO_PROBE)
        rc, wire, st = got.get(timeout=30)                    # stream start: the full list, all Healthy
        assert rc == 0 and st.n_devices == 63
        assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in V.ListAndWatchResponse.FromString(wire).devices] == want
        assert got.empty()                                    # nothing more until a tick
        w.beat()
        w.beat()
        for _ in range(2):
            rc, wire2, st = got.get(timeout=30)
            assert rc == 0 and wire2 == wire and st.node_healthy
        w.stop()                                              # p.signal: the loop ends, the thread is joined
        assert got.empty()


def test_watch_ticker(pkg, tmp_path):
    root = str(tmp_path / "t")
    pkg.synth.write_b200_tree(root, n_gpus=4)
    got = queue.Queue()
    with pkg.Context("kfd:" + root) as ctx:
        t0 = time.monotonic()
        w = ctx.watch(lambda rc, wire, st: got.put(time.monotonic()), pulse_ms=50, flags=pkg._native.LW_NO_PROBE)
        stamps = [got.get(timeout=30) for _ in range(4)]       # initial + 3 ticks
        w.stop()
    assert stamps[0] - t0 < 3.0
    gaps = [b - a for a, b in zip(stamps, stamps[1:])]
    assert all(0.03 < g < 1.5 for g in gaps), gaps              # 50 ms ticks; generous upper bound for loaded hosts


def test_watch_heterogeneous_resource_without_devices_sends_nothing(pkg, kfd, tmp_path):
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi308"), compute="cpx", memory="nps1",
                            hetero_second=("spx", "nps1"))
    got = queue.Queue()
    with pkg.Context("kfd:" + root) as ctx:
        w = ctx.watch(lambda rc, wire, st: got.put((rc, st.n_devices)), resource="gpu", flags=pkg._native.LW_NO_PROBE)
        w.beat()
        time.sleep(0.3)
        w.stop()
        assert got.empty()                                    # plugin.go:296-298
        w = ctx.watch(lambda rc, wire, st: got.put((rc, st.n_devices)), resource="spx_nps1", flags=pkg._native.LW_NO_PROBE)
        assert got.get(timeout=30) == (0, 16)
        w.stop()


def test_watch_can_be_stopped_from_its_own_callback(pkg, tmp_path):
    """b2dp_watch_stop() called on the loop's own thread (from the callback) must not try to join itself:
    the loop ends after the current send and frees the handle."""
    root = str(tmp_path / "t")
    pkg.synth.write_b200_tree(root, n_gpus=2)
    got = queue.Queue()
    holder = {}
    with pkg.Context("kfd:" + root) as ctx:
        sends = [0]

        def cb(rc, wire, st):
            sends[0] += 1
            got.put(len(wire))
            if sends[0] == 2:
                while "w" not in holder:                    # ctx.watch() may not have returned yet
                    time.sleep(0.001)
                holder["w"].stop()                          # second send: stop from inside
        holder["w"] = ctx.watch(cb, pulse_ms=20, flags=pkg._native.LW_NO_PROBE)
        assert got.get(timeout=30) > 0 and got.get(timeout=30) > 0
        time.sleep(0.3)
        assert got.empty()                                  # no third send: the loop is gone
