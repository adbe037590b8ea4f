#!/bin/bash
# How the runs of profiles/r2_emulator.md were made (CPU only).  Usage: profiles/r2_emulator_runs.sh plain|asan|tsan|late [test files...]
# Builds tests/cpp/libb2rpc_emul*.so from the product sources (tests/cpp/gen_emul_lib.py + cuda_emul.h) and runs the GPU test files against it.
set -u
cd "$(dirname "$0")/.."
MODE=${1:-plain}; shift || true
FILES=${*:-$(ls tests/test_gpu_*.py)}
OUT=/tmp/emul_$MODE; mkdir -p $OUT
python tests/cpp/gen_emul_lib.py
SAN=""; PRE=""
case $MODE in
  asan) SAN="-fsanitize=address -fno-omit-frame-pointer"; PRE=$(gcc -print-file-name=libasan.so); export ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 ;;
  tsan) SAN="-fsanitize=thread"; PRE=$(gcc -print-file-name=libtsan.so); export TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 history_size=4" ;;
  late) export B2_EMUL_ASYNC=late ;;
esac
g++ -O1 -g -std=c++17 -fPIC -shared -w -pthread $SAN -I include -I brpc_b200/csrc -o $OUT/libb2rpc_emul.so tests/cpp/emul_lib.cc -ldl || exit 1
export B2_EMUL_LIB=$OUT/libb2rpc_emul.so B2_FUZZ_SECONDS=150
for f in $FILES; do
  b=$(basename $f .py); s=$(date +%s)
  LD_PRELOAD=$PRE timeout 3400 python tests/emul_runner.py $f -m gpu -q -x -p no:faulthandler > $OUT/$b.log 2>&1
  echo "$b rc=$? $(( $(date +%s) - s ))s sanitizer_reports=$(grep -c 'WARNING: ThreadSanitizer\|ERROR: AddressSanitizer' $OUT/$b.log) $(grep -E ' passed| failed' $OUT/$b.log | tail -1)" | tee -a $OUT/summary.txt
done
