# Convenience targets (the driver uses __graft_entry__.build() / pytest / bench.py directly).
PY ?= python

build:            ## nvcc sm_100a -> adpres_b200/libadpres_b200.so, gcc -> oracle/liboracle.so
	$(PY) -c "import __graft_entry__ as g; g.build()"

test: build       ## CPU tests: oracle vs golden values, host logic, ABI, host check of the %XTAB device code
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu: build   ## parity tests on a B200
	$(PY) -m pytest tests -x -q -m gpu

bench: build
	$(PY) bench.py

fixtures:         ## regenerate tests/golden from the reference tree (needs /root/reference)
	$(PY) tests/golden/make_fixtures.py

clean:
	rm -rf adpres_b200/build adpres_b200/libadpres_b200.so oracle/liboracle.so oracle/_ref

.PHONY: build test test-gpu bench fixtures clean
