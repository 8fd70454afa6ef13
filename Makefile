# Convenience targets; the driver uses __graft_entry__.build(), pytest and bench.py directly.
PY ?= python
.PHONY: build test test-gpu bench clean
build:
	$(PY) -c "import __graft_entry__ as g; g.build()"
test: build
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu: build          # needs a B200 (gpurun -- make test-gpu)
	$(PY) -m pytest tests -x -q -m gpu
bench: build             # needs a B200
	$(PY) bench.py --gpus 1 --steps 200 --warmup 5
clean:
	$(MAKE) -C k8s-device-plugin_b200/csrc clean
	$(MAKE) -C oracle clean
	rm -f tools/probe_sweep tools/p2p_sweep tools/doorbell_probe tools/b200dp_cli tools/b200dp_kubelet_sim k8s-device-plugin_b200/b200dp_plugind k8s-device-plugin_b200/b200dp_probe_helper
