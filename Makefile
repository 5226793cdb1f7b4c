# Convenience targets; the driver uses __graft_entry__.build() / pytest / bench.py directly.
PY ?= python

all:            ## CUDA library (sm_100a), CPU oracle, native host library + `raingun` CLI
	$(MAKE) -C raingun_b200/csrc -j4
	$(MAKE) -C oracle
	$(MAKE) -C raingun_b200/host

test:           ## everything that needs no GPU (oracle vs goldens, host library, ABI, gloo sharding)
	$(PY) -m pytest tests -q -m "not gpu"

test-gpu:       ## the parity tests proper (needs a B200)
	$(PY) -m pytest tests -q -m gpu

bench:          ## headline benchmark, one GPU
	$(PY) bench.py

clean:
	$(MAKE) -C raingun_b200/csrc clean
	$(MAKE) -C oracle clean
	$(MAKE) -C raingun_b200/host clean

.PHONY: all test test-gpu bench clean
