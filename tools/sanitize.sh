#!/usr/bin/env bash
# compute-sanitizer over small instances of every hand-written kernel family (run on a B200 under gpurun).
# One process per tool runs all cases (importing torch under the sanitizer costs about a minute); the per-tool log and a
# one-line summary per tool go to gpurun_out/sanitizer/.  Copy them into profiles/ to have them judged.
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out/sanitizer
mkdir -p "$OUT"
CASES=${CASES:-"attn_fwd_2tile,attn_fwd_3tile,attn_fwd_doc,attn_bwd_2tile,attn_bwd_doc_rope,gemm_kk_multi,gemm_mm_tails,gemm_resid,gemm_rope,gemm_wgrad_atomic_split,lmhead_ce,bandwidth"}
TOOLS=${TOOLS:-"memcheck synccheck racecheck"}
: > "$OUT/summary.txt"
for tool in $TOOLS; do
  log="$OUT/${tool}.log"
  timeout "${SAN_TIMEOUT:-600}" compute-sanitizer --tool "$tool" --print-limit 30 \
    python tools/gpu_kernel_check.py --cases "$CASES" > "$log" 2>&1
  rc=$?
  errs=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" "$log" | tail -1)
  res=$(grep -c '^RESULT ' "$log")
  n=$(echo "$CASES" | tr ',' '\n' | wc -l)
  echo "$tool rc=$rc cases_completed=$res/$n :: ${errs:-no summary line (timeout or crash)}" | tee -a "$OUT/summary.txt"
done
