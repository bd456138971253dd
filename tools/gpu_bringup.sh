#!/usr/bin/env bash
# First-contact GPU run: stages in increasing order of risk, each in its own process under a
# timeout; the script stops at the first failing stage so a faulting kernel is never re-run.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n "${TAIL_LINES:-30}" "gpurun_out/$name.log" | tee -a gpurun_out/summary.txt
  if [ $rc -ne 0 ]; then echo "STOP after $name" | tee -a gpurun_out/summary.txt; exit 1; fi
}
[ -n "${SKIP_GEMM_STORE:-}" ] || run gemm_tiny 240 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s -k "gemm_store_bias and 16-128-64"
[ -n "${SKIP_GEMM_STORE:-}" ] || run gemm_store 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s -k "gemm_store"
run gemm_rest 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s -k "gemm and not gemm_store"
run attention 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -s -k "attention"
TAIL_LINES=120 run model_shrunk 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -s
echo "ALL STAGES PASSED" | tee -a gpurun_out/summary.txt
