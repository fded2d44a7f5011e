# Round 2, session 18: compute-sanitizer over the small parity tests of every kernel touched this round
# (memcheck everywhere, racecheck on the shared-memory-heavy kernels).
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest -x -q \
  "tests/test_gnofix_gpu.py::test_gnofix_matches_reference_golden" "tests/test_gnofix_crf_gpu.py::test_gnofix_crf_device_tensors_and_numpy_oracle" \
  "tests/test_svc_gpu.py::test_kernel_matches_reference_golden" "tests/test_svc_gpu.py::test_svc_proba_matches_reference_golden_and_oracle" \
  "tests/test_crf_gpu.py" "tests/test_gbt_gpu.py::test_gbt_rows_matches_oracle_and_slide_window" "tests/test_gbt_gpu.py::test_gbt_ragged_trees_and_nan_default" \
  "tests/test_gbt_gpu.py::test_gbt_thresholds_hit_exactly" "tests/test_gbt_gpu.py::test_gbt_many_windows_are_segmented" \
  "tests/test_lr_gpu.py::test_lr_numpy_in_numpy_out" "tests/test_lr_gpu.py::test_lr_empty_and_bad_shape" \
  "tests/test_pack_gpu.py" "tests/test_calibrator_gpu.py" "tests/test_gbt_gpu.py::test_gbt_smooth_matches_oracle" \
  > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo memcheck rc=$?
tail -5 gpurun_out/r2_sanitize_memcheck.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/r2_sanitize_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 10 python -m pytest -x -q \
  "tests/test_gnofix_gpu.py::test_gnofix_matches_reference_golden" "tests/test_gbt_gpu.py::test_gbt_rows_matches_oracle_and_slide_window" \
  "tests/test_gbt_gpu.py::test_gbt_thresholds_hit_exactly" "tests/test_gbt_gpu.py::test_gbt_ragged_trees_and_nan_default" \
  "tests/test_crf_gpu.py::test_crf_matches_oracle" "tests/test_calibrator_gpu.py::test_calibrate_matches_reference_golden" \
  "tests/test_gnofix_crf_gpu.py::test_gnofix_crf_device_tensors_and_numpy_oracle" \
  > gpurun_out/r2_sanitize_racecheck.log 2>&1; echo racecheck rc=$?
tail -4 gpurun_out/r2_sanitize_racecheck.log
