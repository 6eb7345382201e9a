#!/usr/bin/env bash
# compute-sanitizer over small instances of every hand-written kernel family (run on a B200 under gpurun).
# Each (tool, case) pair runs in its own process under a timeout; one summary line per pair goes to
# gpurun_out/sanitizer_summary.txt and the full logs next to it.  Copy the summary into profiles/ to have it judged.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitizer
mkdir -p "$OUT"
CASES=${CASES:-"attn_fwd_2tile attn_fwd_3tile attn_fwd_doc attn_bwd_2tile attn_bwd_doc_rope gemm_kk_multi gemm_mm_tails gemm_resid gemm_rope gemm_wgrad_atomic_split bandwidth"}
TOOLS=${TOOLS:-"memcheck racecheck synccheck"}
: > gpurun_out/sanitizer_summary.txt
for tool in $TOOLS; do
  for c in $CASES; do
    log="$OUT/${tool}_${c}.log"
    timeout "${SAN_TIMEOUT:-240}" compute-sanitizer --tool "$tool" --print-limit 20 \
      python tools/gpu_kernel_check.py --case "$c" > "$log" 2>&1
    rc=$?
    errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
    res=$(grep -c '^RESULT ' "$log")
    echo "$tool $c rc=$rc result_lines=$res :: ${errs:-no summary line (timeout or crash)}" | tee -a gpurun_out/sanitizer_summary.txt
  done
done
