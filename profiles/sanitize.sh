#!/bin/bash
# compute-sanitizer pass over one small invocation of every kernel family (memcheck, racecheck, synccheck).
# Usage (on the GPU box): bash profiles/sanitize.sh <out_dir> [tool ...]
OUT=${1:-gpurun_out/sanitize}; shift
TOOLS=${@:-memcheck racecheck synccheck}
mkdir -p $OUT
SEL='tests/test_gpu_snv_tc.py::test_bf16_forward_matches_reference[hs_AT]
tests/test_gpu_snv_tc.py::test_dense_lattice_equals_per_site_stages[0]
tests/test_gpu_snv_tc.py::test_bf16_tiny_batches[129]
tests/test_gpu_snv_tc.py::test_local_branch_tensor_core_equals_fp32_kernel
tests/test_gpu_snv_forward.py::test_fp32_forward_matches_reference[hs_AT]
tests/test_gpu_snv_train.py::test_train_forward_backward_vs_autograd[ex_ckpt6]
tests/test_gpu_snv_train.py::test_fused_optimizer_matches_torch[AdamW]
tests/test_gpu_indel.py::test_indel_forward_matches_reference[ex_indel9]
tests/test_gpu_indel_train.py::test_indel_train_forward_backward_vs_autograd[hs_ins-500-6]
tests/test_gpu_evaluation.py::test_group_tables_exact_and_reproducible
tests/test_gpu_evaluation.py::test_window_runs_vs_oracle_large_and_edges
tests/test_gpu_encode.py::test_encode_edge_cases
tests/test_gpu_predict_pipeline.py::test_tsv_matches_oracle_pipeline[True]'
for tool in $TOOLS; do
  i=0
  for t in $SEL; do
    i=$((i+1))
    log=$OUT/${tool}_$i.log
    echo "== $tool :: $t" > $log
    timeout 600 compute-sanitizer --tool $tool --target-processes all --print-limit 20 \
      python -m pytest "$t" -x -q -p no:cacheprovider >> $log 2>&1
    echo "rc=$?" >> $log
    echo "$tool $t: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=' $log | tr '\n' ' ')"
  done
done 2>&1 | tee $OUT/summary.txt
