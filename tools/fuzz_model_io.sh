#!/usr/bin/env bash
# Builds csrc/model_io.cpp with AddressSanitizer + UBSan and runs the damaged-file test of tests/test_model_io.py against it.
set -eu
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -ffp-contract=off -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined -I include \
    divshot_b200/csrc/model_io.cpp -o /tmp/dvs_model_io_asan.so -lz
ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1 DVS_MODEL_IO_LIB=/tmp/dvs_model_io_asan.so \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    python -m pytest tests/test_model_io.py -q -k "forged or round_trip or lossy" -p no:cacheprovider
