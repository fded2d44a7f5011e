# Round 2, session 21: compute-sanitizer synccheck (named team barriers of K6, CTA barriers of K4) and initcheck on the kernels of the round.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 99 --print-limit 10 python -m pytest -x -q \
  "tests/test_gnofix_gpu.py::test_gnofix_matches_reference_golden" "tests/test_gbt_gpu.py::test_gbt_thresholds_hit_exactly" \
  "tests/test_gbt_gpu.py::test_gbt_ragged_trees_and_nan_default" "tests/test_crf_gpu.py::test_crf_matches_oracle" \
  "tests/test_gnofix_crf_gpu.py::test_gnofix_crf_device_tensors_and_numpy_oracle" \
  > gpurun_out/r2_sanitize_synccheck.log 2>&1; echo synccheck rc=$?
tail -4 gpurun_out/r2_sanitize_synccheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 99 --print-limit 10 python -m pytest -x -q \
  "tests/test_gnofix_gpu.py::test_gnofix_matches_reference_golden" "tests/test_gbt_gpu.py::test_gbt_thresholds_hit_exactly" \
  "tests/test_gnofix_crf_gpu.py::test_gnofix_crf_device_tensors_and_numpy_oracle" "tests/test_lr_gpu.py::test_lr_wide_model_with_four_limbs_stays_selectable" \
  > gpurun_out/r2_sanitize_initcheck.log 2>&1; echo initcheck rc=$?
tail -4 gpurun_out/r2_sanitize_initcheck.log
