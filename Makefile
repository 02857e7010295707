# Convenience targets; the contract entry points are __graft_entry__.py (build / smoke), tests/ and bench.py.
PY ?= python

.PHONY: build test test-gpu smoke bench bench-reference clean

build:            ## compile libgradus_b200.so (sm_100a) and the CPU oracle (test infrastructure)
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## oracle vs the reference's golden values, host logic, C ABI, gloo sharding (no GPU needed)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## CUDA path vs the oracle through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu

smoke: build
	$(PY) -c "import __graft_entry__ as g; g.smoke()"

bench: build      ## one JSON line: value, e2e, roofline, cpu_baseline, clocks
	$(PY) bench.py

bench-reference: build
	$(PY) bench.py --impl reference

clean:
	$(MAKE) -C gradus.jl_b200/csrc clean
	$(MAKE) -C oracle clean
