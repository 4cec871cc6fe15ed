#!/bin/bash
# The product's host shim (swarmkit_b200/csrc/scheduler_host.cpp) under AddressSanitizer + UBSan: the test-only library
# that links it against the oracle's ABI is rebuilt with the sanitizers, the shim's CPU suites run against it, and the
# regular library is put back.  CPU only (the same source drives the CUDA engine on the GPU box).
set -e
cd "$(dirname "$0")/.."
ASAN=$(g++ -print-file-name=libasan.so)
(cd oracle && g++ -O1 -g -std=c++17 -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -shared shim_on_oracle.cpp flat_oracle.cpp -o _build/libshim_on_oracle.so)
trap '(cd oracle && rm -f _build/libshim_on_oracle.so && make -s)' EXIT
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python -m pytest tests/test_shim_volumes_cpu.py tests/test_shim_cpu.py -x -q -p no:cacheprovider
