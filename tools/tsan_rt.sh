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
python -m pytest tests/test_kvstore.py tests/test_parallel_cpu.py -x -q -p no:cacheprovider -k "${1:-engine or storage or recordio or reducer or planner}" 2>&1 | tail -5
n=$(ls $out/tsan.* 2>/dev/null | wc -l)
echo "TSAN report files: $n"
grep -h "SUMMARY: ThreadSanitizer" $out/tsan.* 2>/dev/null | sort | uniq -c | sort -rn | head -30

# ---- whole parameter-server jobs with every role under TSAN.  The launchers start their children through `sh -c`, and a
# shell with libtsan preloaded crashes at start-up, so the processes are started by hand here.
run_job() { # tag, workers, script [args...]   (environment of the caller is inherited: PS_RESEND, PS_DROP_MSG, ...)
  local tag=$1 nw=$2 script=$3; shift 3
  local port=$((9600 + RANDOM % 300)) pids="" rc=0
  export DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=$port DMLC_NUM_WORKER=$nw DMLC_NUM_SERVER=2 PYTHONPATH=$PWD
  export TSAN_OPTIONS="halt_on_error=0:exitcode=0:report_signal_unsafe=0:log_path=$out/job_$tag:history_size=4"
  rm -f $out/job_$tag.*
  DMLC_ROLE=scheduler timeout 300 python -c "import deeplearning_cfn_b200.kv" > $out/sched_$tag.log 2>&1 & pids="$pids $!"
  for i in 0 1; do DMLC_ROLE=server timeout 300 python -c "import deeplearning_cfn_b200.kv" > $out/server${i}_$tag.log 2>&1 & pids="$pids $!"; done
  for i in $(seq 1 $nw); do DMLC_ROLE=worker timeout 300 python $script "$@" > $out/worker${i}_$tag.log 2>&1 & pids="$pids $!"; done
  for p in $pids; do wait $p || rc=$?; done
  echo "$tag: exit $rc; $(grep -h "_OK\|PROFILED" $out/worker*_$tag.log | tr '\n' ' ')"
  grep -h "SUMMARY: ThreadSanitizer" $out/job_$tag.* 2>/dev/null | grep -v "libtorch\|python3\|libgomp\|memintrinsics" | sort | uniq -c | sort -rn | head -20
}
set +e
echo "== parameter-server jobs under TSAN (scheduler + 2 servers + workers)"
run_job dist_sync 2 tests/mp/dist_kvstore_worker.py dist_sync
run_job dist_async 2 tests/mp/dist_kvstore_worker.py dist_async
PS_RESEND=1 PS_RESEND_TIMEOUT=200 PS_DROP_MSG=10 PS_HEARTBEAT_INTERVAL=1 run_job resend_drop_heartbeat 2 tests/mp/dist_kvstore_worker.py dist_sync
MXNET_KVSTORE_BIGARRAY_BOUND=40 run_job sync_replicas_python_callbacks 3 tests/mp/sync_replicas_worker.py
run_job server_profiling 2 tests/mp/server_profile_worker.py $out/prof
# round 2: shard boundaries on key units (2-bit words, row-sparse rows), dtype conversion, staging-buffer reuse under queued pushes
run_job store_regressions_big_keys 2 tests/mp/dist_regress_worker.py store
