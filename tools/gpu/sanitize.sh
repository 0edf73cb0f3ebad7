#!/bin/bash
# compute-sanitizer (memcheck, racecheck) over a small-shape subset of the GPU suite; logs -> gpurun_out/r02b_sanitizer_<tool>.log
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
SUBSET="tests/test_histogram_gpu.py::test_golden_vectors_all_strategies tests/test_histogram_gpu.py::test_out_of_bounds_raises_index_error_like_numpy tests/test_histogram_gpu.py::test_unaligned_event_pointer tests/test_vit_model_gpu.py::test_tiny_pt_vit_matches_golden_and_oracle tests/test_vit_model_gpu.py::test_tiny_ft_vit_matches_golden_and_oracle tests/test_vit_model_gpu.py::test_tiny_ft_vit_cls_token_head tests/test_dvae_gpu.py::test_tiny_tokens_and_logits_vs_golden tests/test_dvae_train_gpu.py::test_training_step_matches_reference tests/test_dvae_train_gpu.py::test_decode_matches_reference tests/test_decode_gpu.py::test_reference_golden tests/test_event_pipeline_gpu.py::test_reference_golden_single_streams tests/test_engine_gpu.py tests/test_engine_ft_gpu.py::test_criteria_kernel_matches_timm_formulas tests/test_randaug_gpu.py::test_module_under_fixed_seeds_vs_reference tests/test_randaug_gpu.py::test_whole_chain_with_rand_aug_vs_reference tests/test_vit_kernels_gpu.py::test_layernorm_bwd_branch_equals_the_two_launches tests/test_vit_kernels_gpu.py::test_vbias_chain_is_the_column_sum_of_dv tests/test_histogram_gpu.py::test_sort_strategy_edges_of_its_layout"
for tool in memcheck racecheck; do
  timeout 1500 $SAN --tool $tool --launch-timeout 0 --error-exitcode 0 --print-limit 20 python -m pytest $SUBSET -q -x -p no:cacheprovider > gpurun_out/r02b_sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" gpurun_out/r02b_sanitizer_$tool.log | sort | uniq -c | head -12
done
# the large-sensor histogram strategies (shared-memory atomics + L2 REDs) under racecheck / memcheck at a reduced size
cat > /tmp/hyb_small.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from mem_b200.process_data import histogram
from oracle.histogram_ref import event_hist_ref
from oracle.make_golden import synth_events
rng = np.random.default_rng(3)
for kind in ("uniform", "edge"):
    ev = synth_events(rng, 1_100_000, 480, 640, kind)
    want = event_hist_ref(ev, 480, 640)
    for s in (0, 5, 6, 7):
        assert np.array_equal(histogram(torch.from_numpy(ev).cuda(), 480, 640, strategy=s).cpu().numpy(), want), (kind, s)
print("hybrid / replicated strategies ok")
PY
for tool in memcheck racecheck; do
  timeout 900 $SAN --tool $tool --launch-timeout 0 --print-limit 20 python /tmp/hyb_small.py > gpurun_out/r02b_sanitizer_hist_large_$tool.log 2>&1
  echo "== hist large $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|ok|Error|hazard" gpurun_out/r02b_sanitizer_hist_large_$tool.log | sort | uniq -c | head
done
