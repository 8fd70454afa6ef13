"""Mirror of cmd/k8s-device-plugin/main.go:93-155: flags, resource list, heartbeat ticker, one
gRPC plugin server per resource, registration with the kubelet.

    python -m k8s-device-plugin_b200.cmd_device_plugin -pulse=10 -resource_naming_strategy=single \
           [-backend cuda:] [-plugin_dir /var/lib/kubelet/device-plugins/]

From dpm (vendor/github.com/kubevirt/device-plugin-manager/pkg/dpm/manager.go) it keeps: plugin
server start retried 3 times 3 s apart (:16-20,205-219) and re-serve + re-register when kubelet.sock
is re-created (kubelet restart, :73-84; polled once a second instead of fsnotify).  The process
exits on SIGINT/SIGTERM like the reference.
"""
import argparse
import os
import signal
import sys
import threading
import time

from . import v1beta1
from .context import Context
from .plugin import AMDGPULister, ParseStrategy, getResourceList
from .server import PluginServer


def main(argv=None):
    ap = argparse.ArgumentParser(prog="k8s-device-plugin", description="B200 GPU device plugin for Kubernetes")
    ap.add_argument("-pulse", type=int, default=0,
                    help="time between health check polling in seconds.  Set to 0 to disable.")        # main.go:109
    ap.add_argument("-resource_naming_strategy", default="single",
                    help="Resource strategy to be used: single or mixed")                               # main.go:110
    ap.add_argument("-backend", default="cuda:", help="libb200dp backend uri (cuda:[opts] | kfd:<sysroot>)")
    ap.add_argument("-resource_namespace", default="amd.com",
                    help="vendor domain of the resources (amd.com/gpu keeps the reference's names; e.g. nvidia.com)")
    ap.add_argument("-plugin_dir", default=v1beta1.DevicePluginPath)
    ap.add_argument("-start_retry_wait", type=float, default=3.0, help="seconds between plugin start attempts")
    # glog's flags (the reference image runs `-logtostderr=true -stderrthreshold=INFO -v=5`, Dockerfile:33): accepted
    # and ignored so existing DaemonSet args keep working
    for g in ("-logtostderr", "-alsologtostderr"):
        ap.add_argument(g, nargs="?", const="true", default="true", help=argparse.SUPPRESS)
    for g in ("-stderrthreshold", "-v", "-log_dir", "-vmodule", "-log_backtrace_at"):
        ap.add_argument(g, default="", help=argparse.SUPPRESS)
    args = ap.parse_args(argv)
    try:
        strategy = ParseStrategy(args.resource_naming_strategy)                                         # main.go:113-117
    except ValueError as e:
        print(e, file=sys.stderr)
        return 1
    from . import plugin as _plugin
    _plugin.RESOURCE_NAMESPACE = args.resource_namespace        # plugin.go:406-408 returns "amd.com"
    ctx = Context(args.backend)
    lister = AMDGPULister(ctx)
    try:
        resources = getResourceList(ctx, strategy)                                                      # main.go:141-146
    except Exception as e:
        print("Error occured: %s" % e, file=sys.stderr)
        return 1
    servers = []
    stop = threading.Event()
    kubelet_sock = os.path.join(args.plugin_dir, "kubelet.sock")

    def sock_identity():
        try:
            st = os.stat(kubelet_sock)
            return (st.st_ino, st.st_ctime_ns)
        except OSError:
            return None

    def start_all():
        for srv in servers:
            srv.stop()
        del servers[:]
        for name in resources:                                                                          # dpm handleNewPlugins
            last = None
            for attempt in range(3):                                                                    # dpm/manager.go:16-20,205-219
                try:
                    plugin = lister.NewPlugin(name)
                    srv = PluginServer(plugin, plugin_dir=args.plugin_dir).start()
                    srv.register()
                    servers.append(srv)
                    last = None
                    break
                except Exception as e:          # noqa: BLE001
                    last = e
                    if stop.wait(args.start_retry_wait):
                        return
            if last is not None:
                print("Failed to start plugin %s: %s" % (name, last), file=sys.stderr)

    start_all()
    seen = sock_identity()
    for s in (signal.SIGINT, signal.SIGTERM):
        signal.signal(s, lambda *_: stop.set())
    if args.pulse > 0:                                                                                  # main.go:129-137
        def beat():
            while not stop.wait(args.pulse):
                lister.Heartbeat.put(True)
        threading.Thread(target=beat, daemon=True).start()
    last_check = time.monotonic()
    while not stop.is_set():
        time.sleep(0.2)
        if time.monotonic() - last_check >= 1.0:                                                        # dpm/manager.go:73-84
            last_check = time.monotonic()
            now = sock_identity()
            if now is not None and now != seen:
                start_all()                      # kubelet restarted: serve again and re-register
            if now is not None:
                seen = now
    for srv in servers:
        srv.stop()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
