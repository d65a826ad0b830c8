#!/bin/bash
# AddressSanitizer run of the native CPU runtime (the reference ran its C++ under ASAN in CI):
# builds csrc/rt with -fsanitize=address into /tmp/dlcfn_asan and runs the parameter-server / engine / storage /
# recordio tests against it.  Usage: bash tools/sanitize_rt.sh [pytest -k expression]
set -e
cd "$(dirname "$0")/.."
out=/tmp/dlcfn_asan
mkdir -p $out
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address -fno-omit-frame-pointer -pthread csrc/rt/*.cpp -o $out/libdlcfn_rt.so -lrt
export DLCFN_RT_LIB=$out/libdlcfn_rt.so
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=0:log_path=$out/asan
rm -f $out/asan.*
python -m pytest tests/test_kvstore.py tests/test_parallel_cpu.py -x -q -k "${1:-dist or engine or planner or storage or recordio or reducer or sync_replicas or profiling or garbage or fuzz}" 2>&1 | tail -5
echo "ASAN reports: $(ls $out/asan.* 2>/dev/null | wc -l)"
for f in $out/asan.*; do [ -f "$f" ] && head -30 "$f"; done 2>/dev/null | head -120
