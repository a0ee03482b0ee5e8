#!/bin/bash
# tools/sanitize.sh TOOL  -- compute-sanitizer over the small-shape parity suite (mbarrier / TMEM / shared-memory
# races are silent in a timing run).  TOOL = memcheck | racecheck | synccheck | initcheck; ONE tool per GPU call (the
# profiling guide: several tools in one call have left a B200 unusable).  Output: gpurun_out/sanitize_TOOL.log; the
# summary line ("ERROR SUMMARY: 0 errors") is what profiles/r2_sanitize.txt records.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/sanitize.sh memcheck'
TOOL=${1:-memcheck}
OUT=gpurun_out/sanitize_${TOOL}.log
mkdir -p gpurun_out
SEL='attention_matches_reference_kernel_golden or quant_matches_reference_kernel_golden or tail_masking or tiny_sequences or kivi_pack or per_warp_quant or mixed_k_quantizer_bit_exact'
# the plain run first: a faulting program under a sanitizer is what the guide warns about
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitize_plain.log 2>&1 || { tail -5 gpurun_out/sanitize_plain.log; echo "plain run failed: not sanitizing"; exit 1; }
timeout 1400 compute-sanitizer --tool $TOOL --print-limit 20 --error-exitcode 3 \
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $OUT 2>&1
RC=$?
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $OUT | tail -5
echo "compute-sanitizer --tool $TOOL rc=$RC"
exit $RC
