#!/bin/bash
# compute-sanitizer pass over one small invocation of every kernel family (memcheck, racecheck, synccheck).
# Usage (on the GPU box): bash profiles/sanitize.sh <out_dir> [tool ...]
# SET=a (default): the round-1 kernel families; SET=b: kernels added or changed in round 2.
OUT=${1:-gpurun_out/sanitize}; shift
TOOLS=${@:-memcheck racecheck synccheck}
mkdir -p $OUT
SEL_a='tests/test_gpu_snv_tc.py::test_bf16_forward_matches_reference[hs_AT]
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
# b: fused tensor-core U-Net level kernels (k_unet_level: shipped shape, site groups + strides, 16-warp CTAs), split-bf16 mma.sync
# conv / weight-gradient kernels, the two-warpgroup local MLP, the bulk-copy weight staging of the stage kernels, the auto-mode
# recompute on its side stream, the tiled MuRaL-indel training kernels, the graph-replayed training step
SEL_b='tests/test_gpu_indel.py::test_indel_forward_matches_reference[hs_ins]
tests/test_gpu_indel.py::test_indel_level_kernels_generic_shapes[8-5-down2-544-True]
tests/test_gpu_indel.py::test_indel_level_kernels_generic_shapes[8-9-down4-3000-True]
tests/test_gpu_conv_mma.py
tests/test_gpu_snv_tc.py::test_local_branch_tensor_core_equals_fp32_kernel
tests/test_gpu_snv_tc.py::test_bf16_forward_matches_reference[hs_AT]
tests/test_gpu_predict_pipeline.py::test_auto_mode_routes_exception_windows_to_fp32
tests/test_gpu_indel_train.py::test_indel_fused_step_and_dropin_loop
tests/test_gpu_snv_train.py::test_graph_step_equals_eager_step'
# c: kernels changed after set b — upsample-folded U-Net level kernels and batch edges, vectorised BatchNorm passes / stems /
# weight-gradient side streams of the MuRaL-snv training step, continuous features in train mode, segmented row reductions,
# grid.z-split convs and the graph-replayed step of MuRaL-indel training
SEL_c='tests/test_gpu_indel.py::test_indel_forward_matches_reference[at_ins]
tests/test_gpu_indel.py::test_indel_level_kernels_batch_edges
tests/test_gpu_snv_train.py::test_train_forward_backward_vs_autograd[ex_ckpt6]
tests/test_gpu_snv_train.py::test_graph_step_equals_eager_step
tests/test_gpu_snv_cont.py
tests/test_gpu_indel_train.py::test_indel_train_forward_backward_vs_autograd[hs_ins-500-6]
tests/test_gpu_indel_train.py::test_indel_graph_step_equals_eager_step'
SET=${SET:-a}
if [ "$SET" = b ]; then SEL=$SEL_b; elif [ "$SET" = c ]; then SEL=$SEL_c; else SEL=$SEL_a; fi
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
