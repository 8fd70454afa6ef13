"""Mirror of cmd/k8s-node-labeller/main.go label arithmetic above the C ABI."""
from typing import Dict, List, Optional

from . import _native as N
from .context import Context


def labelGeneratorNames() -> List[str]:
    """The 12 generators (main.go:115-379), sorted."""
    rc, arr, n = N.grow_call(lambda cap: (N.Id64 * cap)(), lambda a, cap, pn: N.lib.b2dp_label_generator_names(a, cap, pn))
    N.check(rc)
    return [N.s(arr[i].value) for i in range(n)]


_domain = "amd.com"      # main.go:39 amdPrefix; "beta." + it is experimentalAMDPrefix (main.go:38)


def setVendorDomain(domain: str) -> None:
    """Publish labels (and, via plugin.RESOURCE_NAMESPACE, resources) under another vendor domain;
    the default "amd.com" keeps the reference's keys."""
    global _domain
    N.check(N.lib.b2dp_set_vendor_domain(domain.encode()))
    _domain = domain


def createLabelPrefix(name: str, experimental: bool) -> str:
    """main.go:76-85."""
    return "%s/gpu.%s" % ("beta." + _domain if experimental else _domain, name)


def createLabels(kind: str, entries: Dict[str, int]) -> Dict[str, str]:
    """main.go:87-108."""
    arr = (N.KvCount * max(1, len(entries)))()
    for i, (k, v) in enumerate(entries.items()):
        arr[i].key, arr[i].count = k.encode(), v
    rc, out, n = N.grow_call(lambda cap: (N.Label * cap)(),
                             lambda a, cap, pn: N.lib.b2dp_create_labels(kind.encode(), arr, len(entries), a, cap, pn))
    N.check(rc)
    return {N.s(x.key): N.s(x.value) for x in out[:n]}


def removeOldNodeLabels(labels: Optional[Dict[str, str]]) -> Optional[Dict[str, str]]:
    """main.go:55-74 on a plain dict (node.Labels); mutates and returns it."""
    if labels is None:
        return None
    import ctypes as C
    arr = (N.Label * max(1, len(labels)))()
    for i, (k, v) in enumerate(labels.items()):
        arr[i].key, arr[i].value = k.encode(), v.encode()
    n = C.c_int(0)
    N.check(N.lib.b2dp_remove_old_node_labels(arr, len(labels), C.byref(n)))
    kept = {N.s(x.key): N.s(x.value) for x in arr[:n.value]}
    labels.clear()
    labels.update(kept)
    return labels


def generateLabels(ctx: Context, lblProps: Dict[str, bool]) -> Dict[str, str]:
    """main.go:383-397: `lblProps` = the labeller's bool flags."""
    return ctx.generate_labels([k for k, v in lblProps.items() if v])


def Reconcile(node_labels: Optional[Dict[str, str]], labels: Dict[str, str]) -> Dict[str, str]:
    """controller.go:23-58 on the node's label map (the K8s client Get/Update around it is the
    caller's job): nil map -> {}, remove this labeller's old labels, set the generated ones."""
    if node_labels is None:
        node_labels = {}
    removeOldNodeLabels(node_labels)
    node_labels.update(labels)
    return node_labels


def node_label_merge_patch(before: Optional[Dict[str, str]], after: Dict[str, str]) -> str:
    """The JSON merge patch (RFC 7386, `kubectl patch node --type merge`) that takes a node's label
    map from `before` to `after`: changed/new labels carry their value, removed ones null.  This is
    the wire body a K8s client sends for the reference's `client.Update` (controller.go:47-55)
    when only labels change."""
    import json
    before = before or {}
    labels = {k: v for k, v in after.items() if before.get(k) != v}
    labels.update({k: None for k in before if k not in after})
    return json.dumps({"metadata": {"labels": labels}}, sort_keys=True, separators=(",", ":"))
