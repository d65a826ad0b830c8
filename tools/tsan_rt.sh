#!/bin/bash
# ThreadSanitizer run of the native CPU runtime (van / heartbeat / resend threads, server executor, engine workers, host
# reducer, record-decoder ring).  The reference has no TSAN job; this is the data-race counterpart of
# tools/sanitize_rt.sh.  Python itself is not instrumented, so only races between the runtime's own threads are reported.
#   bash tools/tsan_rt.sh [pytest -k expression]
set -e
cd "$(dirname "$0")/.."
out=/tmp/dlcfn_tsan
mkdir -p $out
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=thread -fno-omit-frame-pointer -pthread csrc/rt/*.cpp -o $out/libdlcfn_rt.so -lrt
export DLCFN_RT_LIB=$out/libdlcfn_rt.so
export LD_PRELOAD=$(gcc -print-file-name=libtsan.so)
export OMP_NUM_THREADS=1 MKL_NUM_THREADS=1   # libgomp is not instrumented: its barriers look like races
export TSAN_OPTIONS="halt_on_error=0:exitcode=0:report_signal_unsafe=0:log_path=$out/tsan:history_size=4:second_deadlock_stack=1"
rm -f $out/tsan.*
python -m pytest tests/test_kvstore.py tests/test_parallel_cpu.py -x -q -p no:cacheprovider -k "${1:-engine or storage or recordio or reducer or dist_sync}" 2>&1 | tail -5
n=$(ls $out/tsan.* 2>/dev/null | wc -l)
echo "TSAN report files: $n"
grep -h "SUMMARY: ThreadSanitizer" $out/tsan.* 2>/dev/null | sort | uniq -c | sort -rn | head -30

# ---- a whole parameter-server job with every role under TSAN.  The launchers start their children through `sh -c`, and a
# shell with libtsan preloaded crashes at start-up, so the five processes are started by hand here.
echo "== parameter-server job under TSAN (scheduler + 2 servers + 2 workers, dist_sync then dist_async)"
for mode in dist_sync dist_async; do
  port=$((9600 + RANDOM % 300))
  export DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=$port DMLC_NUM_WORKER=2 DMLC_NUM_SERVER=2 PYTHONPATH=$PWD
  export TSAN_OPTIONS="halt_on_error=0:exitcode=0:report_signal_unsafe=0:log_path=$out/job_$mode:history_size=4"
  rm -f $out/job_$mode.*
  pids=""
  DMLC_ROLE=scheduler timeout 300 python -c "import deeplearning_cfn_b200.kv" > $out/sched_$mode.log 2>&1 & pids="$pids $!"
  for i in 0 1; do DMLC_ROLE=server timeout 300 python -c "import deeplearning_cfn_b200.kv" > $out/server${i}_$mode.log 2>&1 & pids="$pids $!"; done
  for i in 0 1; do DMLC_ROLE=worker timeout 300 python tests/mp/dist_kvstore_worker.py $mode > $out/worker${i}_$mode.log 2>&1 & pids="$pids $!"; done
  rc=0
  for p in $pids; do wait $p || rc=$?; done
  echo "$mode: exit $rc; $(grep -h "_OK" $out/worker*_$mode.log | tr '\n' ' ')"
  grep -h "SUMMARY: ThreadSanitizer" $out/job_$mode.* 2>/dev/null | grep -v "libtorch\|python3\|libgomp\|memintrinsics" | sort | uniq -c | sort -rn | head -20
done
