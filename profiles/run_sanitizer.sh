#!/bin/bash
# compute-sanitizer passes over the smallest parity tests of each kernel family (SURVEY §5.2 race / sync tier).
# Usage (on the GPU box): bash profiles/run_sanitizer.sh <out_dir>
# memcheck / synccheck / racecheck slow a kernel down 10-100x, so only the small shapes run here.
OUT=${1:-gpurun_out/sanitizer}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='test_gemm_fwd_matches_oracle and auto and 130-100-256 or test_dgrad_and_wgrad_match_oracle and auto and 130-100-256 or test_agg_fwd_matches_oracle and 100 or test_agg_bwd_matches_oracle and 100 or test_sampler_bit_exact_vs_oracle and fan1 or test_fused_step_matches_oracle_and_autograd and fan1'
for tool in memcheck synccheck racecheck; do
  timeout 1500 $CS --tool $tool --print-limit 20 --log-file "$OUT/$tool.log" \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -x -q -m gpu -k "$SEL" > "$OUT/$tool.pytest.txt" 2>&1
  echo "$tool rc=$?" >> "$OUT/summary.txt"
  tail -n 3 "$OUT/$tool.log" >> "$OUT/summary.txt"
done
cat "$OUT/summary.txt"
