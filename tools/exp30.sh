#!/bin/bash
# final check of the round: new-row tests with the fused kernel restricted to short-row matrices, then a short bench run
# (1 timed step, no CPU leg) that exercises the new time_to_solution object
mkdir -p gpurun_out
( timeout 60 python -m pytest tests/test_zz_next_rows_gpu.py -q -m gpu 2>&1 | tail -30 ) > gpurun_out/r02_pytest_next_rows_final.log
tail -3 gpurun_out/r02_pytest_next_rows_final.log
timeout 80 python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-cpu --secondary-steps 10 > gpurun_out/r02_bench_short_tts.json 2> gpurun_out/r02_bench_short_tts.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_short_tts.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "check", d["check"]["ok"])
    print("tts", json.dumps(d.get("time_to_solution")))
    print("tts3d", json.dumps(d.get("secondary", {}).get("time_to_solution")))
except Exception as e:
    print("parse failed", e)
    print(open("gpurun_out/r02_bench_short_tts.err").read()[-1500:])
PY
