"""The two `main`s (cmd/k8s-device-plugin, cmd/k8s-node-labeller) on the kfd: backend.  CPU only."""
import importlib
import io
import json
import os
import sys
import threading
import time

import grpc

import fake_sysfs
from oracle import labeller as olab
from test_grpc_host import FakeKubelet
from test_oracle_golden import topo_dir


def test_node_labeller_main(pkg, kfd, tmp_path, capsys, monkeypatch):
    lab = importlib.import_module("k8s-device-plugin_b200.cmd_node_labeller")
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    assert lab.main(["-vram", "-cu-count", "-compute-memory-partition", "-backend", "kfd:" + root]) == 0
    got = json.loads(capsys.readouterr().out)
    want = olab.generateLabels({"vram": True, "cu-count": True, "compute-memory-partition": True}, root)
    assert got == want and got["amd.com/gpu.compute-memory-partition"] == "cpx_nps4"
    # no flags => no labels (every generator defaults to false, main.go:407-409)
    assert lab.main(["-backend", "kfd:" + root]) == 0
    assert json.loads(capsys.readouterr().out) == {}
    monkeypatch.setattr(sys, "stdin", io.StringIO(json.dumps({"beta.amd.com/gpu.vram": "1G",
                                                               "beta.amd.com/gpu.vram.1G": "8", "keep": "me"})))
    assert lab.main(["-vram", "-reconcile", "-backend", "kfd:" + root]) == 0
    out = json.loads(capsys.readouterr().out)
    assert out == {"keep": "me", **olab.generateLabels({"vram": True}, root)}


def test_device_plugin_main_bad_strategy(pkg, capsys):
    dp = importlib.import_module("k8s-device-plugin_b200.cmd_device_plugin")
    assert dp.main(["-resource_naming_strategy", "both"]) == 1            # main.go:113-117: exit(1)
    assert "invalid resource naming strategy: both" in capsys.readouterr().err


def test_device_plugin_main_registers_and_beats(pkg, kfd, tmp_path, short_dir):
    dp = importlib.import_module("k8s-device-plugin_b200.cmd_device_plugin")
    V = pkg.v1beta1
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "mi308"), compute="cpx", memory="nps1",
                            hetero_second=("spx", "nps1"))
    plug_dir = short_dir
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    rc = {}

    def run():
        # signal handlers only install on the main thread; patch them out for the test thread
        import signal as _s
        orig = _s.signal
        _s.signal = lambda *a, **k: None
        try:
            rc["v"] = dp.main(["-pulse", "1", "-resource_naming_strategy", "mixed", "-backend", "kfd:" + root,
                               "-plugin_dir", plug_dir])
        finally:
            _s.signal = orig
    th = threading.Thread(target=run, daemon=True)
    th.start()
    regs = sorted((kubelet.requests.get(timeout=30) for _ in range(2)), key=lambda r: r.resource_name)
    assert [r.resource_name for r in regs] == ["amd.com/cpx_nps1", "amd.com/spx_nps1"]      # heterogeneous + mixed
    assert [r.endpoint for r in regs] == ["amd.com_cpx_nps1", "amd.com_spx_nps1"]
    with grpc.insecure_channel("unix://" + os.path.join(plug_dir, "amd.com_cpx_nps1")) as ch:
        stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                 response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
        first = next(stream)
        assert len(first.devices) == 16
        second = next(stream)                                           # arrives with the -pulse ticker
        assert [d.ID for d in second.devices] == [d.ID for d in first.devices]
        stream.cancel()
    # kubelet restart: kubelet.sock is re-created => the plugin serves again and re-registers
    # (dpm/manager.go:73-84)
    kubelet.server.stop(0).wait()
    sock = os.path.join(plug_dir, "kubelet.sock")
    if os.path.exists(sock):
        os.unlink(sock)
    time.sleep(0.3)
    kubelet2 = FakeKubelet(sock, V)
    regs = sorted((kubelet2.requests.get(timeout=30) for _ in range(2)), key=lambda r: r.resource_name)
    assert [r.resource_name for r in regs] == ["amd.com/cpx_nps1", "amd.com/spx_nps1"]
    with grpc.insecure_channel("unix://" + os.path.join(plug_dir, "amd.com_spx_nps1")) as ch:
        stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                 response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
        assert len(next(stream).devices) == 16
        stream.cancel()
    kubelet2.server.stop(0)
