#!/bin/bash
# simplification stage on the GPU: parity tests, then the bench line (config 3) with its "simplify" entry
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "simplify or cpp_dropin" > gpurun_out/r2f1_tests.log 2>&1
rc=$?
tail -12 gpurun_out/r2f1_tests.log
if [ $rc -eq 0 ]; then
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f1_bench.json 2> gpurun_out/r2f1_bench.err
  echo "bench rc=$?"; python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2f1_bench.json").read().strip().splitlines()[-1])
print({k: l[k] for k in ("value", "ms_per_step", "parity", "simplify")})
PY
fi
exit $rc
