# Convenience targets; everything is also reachable from Python (__graft_entry__.py, parallelfdtd_b200/build.py).
PY ?= python

.PHONY: build test test-gpu smoke bench bench-reference sanitize clean
build:            ## nvcc (sm_100a) + g++: C-ABI library, host layer, Python module, MEX driver, oracle, reference build
	$(PY) __graft_entry__.py
test: build       ## CPU suite (oracle vs golden vectors, host logic, ABI, gloo slab tests)
	$(PY) -m pytest tests -x -q -m "not gpu"
test-gpu:         ## parity suite through the C ABI (needs a B200)
	$(PY) -m pytest tests -x -q -m gpu
smoke:
	$(PY) -c "import __graft_entry__ as g; g.smoke()"
bench:            ## one JSON line on stdout
	$(PY) bench.py
bench-reference:
	$(PY) bench.py --impl reference
sanitize:         ## compute-sanitizer over every kernel family (needs a GPU)
	for t in memcheck racecheck synccheck; do compute-sanitizer --tool $$t --error-exitcode 3 $(PY) tools/sanitize_case.py 6 || exit 1; done
clean:
	find . -name "*.o" -not -path "./.git/*" -delete; rm -f parallelfdtd_b200/*.so oracle/*.so tests/cpp/host_tests tests/cpp/mex_tests; rm -rf oracle/_ref
