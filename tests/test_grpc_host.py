"""A fake kubelet against the gRPC host (server.py) on the kfd: backend -- what the reference never
tested: Registration.Register on kubelet.sock, then the kubelet drives
GetDevicePluginOptions / ListAndWatch (stream + heartbeat) / GetPreferredAllocation / Allocate over
the plugin's unix socket.  CPU only."""
import importlib
import os
import queue
from concurrent import futures

import grpc
import pytest

import fake_sysfs
from oracle import allocator as oalloc
from oracle import amdgpu as oamd
from oracle import plugin as oplug
from test_oracle_golden import topo_dir


class FakeKubelet:
    def __init__(self, sock, V):
        self.requests = queue.Queue()
        self.V = V
        outer = self

        class H(grpc.GenericRpcHandler):
            def service(self, d):
                if d.method != V.REGISTER:
                    return None

                def reg(req, ctx):
                    outer.requests.put(V.RegisterRequest.FromString(req))
                    return V.Empty().SerializeToString()
                return grpc.unary_unary_rpc_method_handler(reg, lambda b: b, lambda b: b)
        self.server = grpc.server(futures.ThreadPoolExecutor(max_workers=2))
        self.server.add_generic_rpc_handlers((H(),))
        self.server.add_insecure_port("unix://" + sock)
        self.server.start()


def _call(ch, method, req, resp_cls):
    return ch.unary_unary(method, request_serializer=lambda m: m.SerializeToString(),
                          response_deserializer=resp_cls.FromString)(req, timeout=30)


def test_kubelet_round_trip(pkg, kfd, tmp_path, short_dir):
    srv_mod = importlib.import_module("k8s-device-plugin_b200.server")
    V = pkg.v1beta1
    root = fake_sysfs.build(str(tmp_path / "r"), topo_dir(kfd, "cpx"), compute="cpx", memory="nps4")
    plug_dir = short_dir
    kubelet = FakeKubelet(os.path.join(plug_dir, "kubelet.sock"), V)
    gpus = oamd.GetAMDGPUs(root)
    ids = sorted(gpus)
    with pkg.Context("kfd:" + root) as ctx:
        lister = pkg.plugin.AMDGPULister(ctx)
        assert lister.GetResourceNamespace() == "amd.com"
        resources = pkg.plugin.getResourceList(ctx, pkg.plugin.ParseStrategy("single"))
        assert resources == ["gpu"]
        plugin = lister.NewPlugin(resources[0])
        server = srv_mod.PluginServer(plugin, plugin_dir=plug_dir).start()
        try:
            server.register()
            reg = kubelet.requests.get(timeout=30)
            assert (reg.version, reg.endpoint, reg.resource_name) == ("v1beta1", "amd.com_gpu", "amd.com/gpu")
            assert reg.options.get_preferred_allocation_available and not reg.options.pre_start_required
            assert os.path.exists(os.path.join(plug_dir, "amd.com_gpu"))

            with grpc.insecure_channel("unix://" + server.socket_path) as ch:
                opts = _call(ch, V.GET_OPTIONS, V.Empty(), V.DevicePluginOptions)
                assert opts.get_preferred_allocation_available
                stream = ch.unary_stream(V.LIST_AND_WATCH, request_serializer=lambda m: m.SerializeToString(),
                                         response_deserializer=V.ListAndWatchResponse.FromString)(V.Empty())
                first = next(stream)
                _, want = oplug.list_and_watch_devices(gpus, "gpu")
                assert [(d.ID, d.health, d.topology.nodes[0].ID) for d in first.devices] == want
                lister.Heartbeat.put(True)                      # main.go:129-137 ticker
                second = next(stream)
                assert [d.ID for d in second.devices] == ids and all(d.health == "Healthy" for d in second.devices)

                opol = oalloc.BestEffortPolicy()
                opol.Init(oplug.getDevices(root), root + "/sys/class/kfd/kfd/topology/nodes")
                req = V.PreferredAllocationRequest(container_requests=[
                    V.ContainerPreferredAllocationRequest(available_deviceIDs=ids, must_include_deviceIDs=[ids[3]],
                                                          allocation_size=4)])
                resp = _call(ch, V.GET_PREFERRED_ALLOCATION, req, V.PreferredAllocationResponse)
                assert list(resp.container_responses[0].deviceIDs) == opol.Allocate(list(ids), [ids[3]], 4)[0]
                bad = V.PreferredAllocationRequest(container_requests=[
                    V.ContainerPreferredAllocationRequest(available_deviceIDs=ids[:2], allocation_size=5)])
                with pytest.raises(grpc.RpcError) as ei:
                    _call(ch, V.GET_PREFERRED_ALLOCATION, bad, V.PreferredAllocationResponse)
                assert ei.value.code() == grpc.StatusCode.UNKNOWN
                assert ei.value.details() == "unable to get preferred allocation list. Error:" + oalloc.invalidAvailable

                areq = V.AllocateRequest(container_requests=[V.ContainerAllocateRequest(devices_ids=[ids[0], ids[1]])])
                aresp = _call(ch, V.ALLOCATE, areq, V.AllocateResponse)
                got = [(d.host_path, d.container_path, d.permissions) for d in aresp.container_responses[0].devices]
                assert got == oplug.allocate_device_specs(gpus, [ids[0], ids[1]])
                assert _call(ch, V.PRE_START_CONTAINER, V.PreStartContainerRequest(), V.PreStartContainerResponse) is not None
                stream.cancel()
        finally:
            server.stop()
            kubelet.server.stop(0)
    assert not os.path.exists(os.path.join(plug_dir, "amd.com_gpu"))
