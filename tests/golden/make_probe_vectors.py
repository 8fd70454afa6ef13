#!/usr/bin/env python3
"""Generate tests/golden/probe_vectors.json from the probe oracle (oracle/probe.py).

The probe has no reference counterpart; these vectors freeze its specification so that neither
the oracle nor the kernels can drift silently.  Re-run only when the pattern definition changes
on purpose:  python tests/golden/make_probe_vectors.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import probe as P  # noqa: E402

cases = []
for n_words, dev in [(1, 0), (4, 0), (1024, 1), (4096 + 4, 2), ((1 << 20) + 148, 7), (1 << 24, 3)]:
    seed = P.initial_seed(dev)
    nxt = P.next_seed(seed)
    src = P.pattern(n_words, seed)
    cs, bad, first, dst = P.probe_pass(src, seed, nxt)
    cases.append({"n_words": n_words, "device": dev, "seed": seed, "next_seed": nxt, "checksum": cs,
                  "head": [int(x) for x in src[:4]], "tail": [int(x) for x in src[-2:]],
                  "dst_head": [int(x) for x in dst[:4]], "dst_checksum": P.checksum(dst)})
full = {"n_words": 1 << 28, "seeds": {}}
s = P.initial_seed(0)
for _ in range(3):
    full["seeds"][str(s)] = P.expected_checksum(1 << 28, s)
    s = P.next_seed(s)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe_vectors.json")
json.dump({"cases": cases, "full_size": full}, open(out, "w"), indent=1)
print("wrote", out)
