#!/bin/bash
# simplification stage on the GPU: parity tests, then the fixture test once more under compute-sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "simplify or cpp_dropin" > gpurun_out/r2f1_tests.log 2>&1
rc=$?
tail -30 gpurun_out/r2f1_tests.log
if [ $rc -eq 0 ]; then
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "simplify_matches_reference_fixtures" > gpurun_out/r2f1_sanitizer.log 2>&1
  echo "sanitizer rc=$?"; tail -8 gpurun_out/r2f1_sanitizer.log
fi
exit $rc
