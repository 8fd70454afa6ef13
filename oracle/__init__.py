"""CPU oracle for the b200 device-plugin hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-Python / numpy restatement of the reference's algorithms
(ROCm/k8s-device-plugin @ dea1db13) for the path named in BASELINE.json:
enumerate -> health-probe -> property-read (ListAndWatch + heartbeat) and the
pair-weight / best-effort allocation path behind GetPreferredAllocation, plus the
labeller's label arithmetic.  Every function cites the reference file:line it
follows; Go standard-library semantics that affect results (filepath.Glob
ordering, unanchored RE2 matches, strconv.ParseInt base-0 rules, bufio.Scanner
line splitting, map-iteration nondeterminism) are restated in `oracle.gosem`.

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` plus its untimed `parity`
checker leg (after every timed region: product answers vs this oracle on the tree
the product exports) -- as the checker or as the timed CPU baseline, never as the
product.  The product
(`k8s-device-plugin_b200/` + `libb200dp.so`) must not import, link or execute
anything under `oracle/`, and it has no CPU fallback for the GPU probe.

Parity pinning: the reference cannot be compiled or run in the build container
(no Go toolchain; cgo needs libdrm/hwloc headers), so there is no `oracle/_ref`.
The oracle is pinned instead against every known-answer value the reference's own
tests hold for this path (amdgpu_test.go, plugin_test.go, device_test.go,
besteffort_policy_test.go, main_test.go) on the reference's own captured kfd
trees (packed under tests/golden/ by tests/golden/make_fixtures.py) -- see
tests/test_oracle_golden.py.  Functions the reference never tests (GetAMDGPUs,
simpleHealthCheck, PopulatePerGPUDHealth, label generators) are "parity
unpinned": they follow the code line by line but have no reference-side vector.
The GPU probe (hbm_probe / p2p_probe) has no reference counterpart at all; its
oracle (oracle/probe.py, oracle/probe_oracle.c) is the closed-form definition of
the pattern arithmetic and is the specification, pinned by its own fixed vectors.
"""
