#!/bin/bash
# UndefinedBehaviorSanitizer run of the native CPU runtime (companion of tools/sanitize_rt.sh and tools/tsan_rt.sh): any report
# aborts the process that hit it, so a finding shows up as a failed test with the report in its output.
#   bash tools/ubsan_rt.sh [pytest -k expression]
set -e
cd "$(dirname "$0")/.."
out=/tmp/dlcfn_ubsan
mkdir -p $out
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -pthread csrc/rt/*.cpp -o $out/libdlcfn_rt.so -lrt
export DLCFN_RT_LIB=$out/libdlcfn_rt.so
export LD_PRELOAD=$(gcc -print-file-name=libubsan.so)
export UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1
python -m pytest tests/test_kvstore.py tests/test_parallel_cpu.py -x -q -p no:cacheprovider \
  -k "${1:-dist or engine or planner or storage or recordio or reducer or sync_replicas or profiling or garbage or fuzz or c_level}" 2>&1 | tail -6
