"""Mirror of cmd/k8s-node-labeller/main.go:399-479 minus the controller-runtime client: one bool
flag per label generator (main.go:407-409), labels generated once at start-up (main.go:430-432) and
printed as JSON -- or merged into a label map read from stdin with the reference's Reconcile
semantics (controller.go:23-58).  Talking to the kube-apiserver is the caller's job (out of scope).

    python -m k8s-device-plugin_b200.cmd_node_labeller -vram -cu-count -product-name [-backend cuda:] [-reconcile]
"""
import argparse
import json
import sys

from .context import Context
from .labeller import Reconcile, generateLabels, labelGeneratorNames, node_label_merge_patch


def main(argv=None):
    ap = argparse.ArgumentParser(prog="k8s-node-labeller", description="B200 GPU Node Labeller for Kubernetes")
    names = labelGeneratorNames()
    for k in names:
        ap.add_argument("-" + k, action="store_true", default=False,
                        help="Set this to label nodes with " + k + " properties")
    ap.add_argument("-backend", default="cuda:")
    ap.add_argument("-reconcile", action="store_true",
                    help="read the node's current labels (JSON object) from stdin and print the reconciled map")
    ap.add_argument("-patch", action="store_true",
                    help="like -reconcile, but print the JSON merge patch (kubectl patch node $DS_NODE_NAME --type merge -p ...)")
    args = ap.parse_args(argv)
    enabled = {k: bool(getattr(args, k.replace("-", "_"))) for k in names}
    with Context(args.backend) as ctx:
        labels = generateLabels(ctx, enabled)
    if args.patch:
        node = json.load(sys.stdin)
        sys.stdout.write(node_label_merge_patch(node, Reconcile(dict(node), labels)) + "\n")
        return 0
    if args.reconcile:
        node = json.load(sys.stdin)
        labels = Reconcile(node, labels)
    json.dump(labels, sys.stdout, indent=1, sort_keys=True)
    sys.stdout.write("\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
