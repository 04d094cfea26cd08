#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_round2.py -m gpu -q -p no:cacheprovider -x > gpurun_out/r2_run4_tests.log 2>&1
echo "tests rc=$?"; tail -15 gpurun_out/r2_run4_tests.log
python - <<'PY' > gpurun_out/r2_run4_kernels.txt 2>&1
import os, sys, json, subprocess
for env in ({}, {"VBX_FORMANT_RESERVE_SMS": "4"}, {"VBX_FORMANT_RESERVE_SMS": "6"}, {"VBX_FORMANT_RESERVE_SMS": "12"}, {"VBX_FORMANT_CHUNKS": "6"}, {"VBX_FORMANT_CHUNKS": "6", "VBX_FORMANT_RESERVE_SMS": "6"}, {"VBX_FORMANT_CHUNKS": "5"}, {"VBX_FORMANT_CHUNKS": "10"}):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "bench.py", "--config", "c3", "--steps", "10", "--warmup", "3", "--no-cpu"], env=e, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(env, "ms/step", round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, "e2e", f'{d["e2e"]["value"]:.3e}', "frac", d["roofline"]["frac"])
    except Exception as ex:
        print(env, "FAILED", ex, out.stderr[-800:])
PY
cat gpurun_out/r2_run4_kernels.txt
