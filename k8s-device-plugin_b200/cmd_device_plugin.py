"""Mirror of cmd/k8s-device-plugin/main.go:93-155: flags, resource list, heartbeat ticker, one
gRPC plugin server per resource, registration with the kubelet.

    python -m k8s-device-plugin_b200.cmd_device_plugin -pulse=10 -resource_naming_strategy=single \
           [-backend cuda:] [-plugin_dir /var/lib/kubelet/device-plugins/]

Not reimplemented from dpm (out of scope): fsnotify re-registration on kubelet
restart and the 3x3 s start retry; the process exits on SIGINT/SIGTERM like the reference.
"""
import argparse
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
    ap.add_argument("-plugin_dir", default=v1beta1.DevicePluginPath)
    args = ap.parse_args(argv)
    try:
        strategy = ParseStrategy(args.resource_naming_strategy)                                         # main.go:113-117
    except ValueError as e:
        print(e, file=sys.stderr)
        return 1
    ctx = Context(args.backend)
    lister = AMDGPULister(ctx)
    try:
        resources = getResourceList(ctx, strategy)                                                      # main.go:141-146
    except Exception as e:
        print("Error occured: %s" % e, file=sys.stderr)
        return 1
    servers = []
    for name in resources:                                                                              # dpm handleNewPlugins
        srv = PluginServer(lister.NewPlugin(name), plugin_dir=args.plugin_dir).start()
        srv.register()
        servers.append(srv)
    stop = threading.Event()
    for s in (signal.SIGINT, signal.SIGTERM):
        signal.signal(s, lambda *_: stop.set())
    if args.pulse > 0:                                                                                  # main.go:129-137
        def beat():
            while not stop.wait(args.pulse):
                lister.Heartbeat.put(True)
        threading.Thread(target=beat, daemon=True).start()
    while not stop.is_set():
        time.sleep(0.2)
    for srv in servers:
        srv.stop()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
