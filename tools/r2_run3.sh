#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lpc_fuseda -c 1 -o gpurun_out/prof_lpca_v1 env VBX_FORMANT_CHUNKS=1 python bench.py --config c3 --utts 1125 --steps 1 --warmup 0 --device-only > gpurun_out/ncu_lpca_v1.log 2>&1
ls -la gpurun_out/*.ncu-rep
python - <<'PY' > gpurun_out/r2_run3_kernels.txt 2>&1
import os, sys, json, subprocess
for env in ({"VBX_LPCA": "0", "VBX_FORMANT_CHUNKS": "1"}, {"VBX_LPCA": "0"}, {"VBX_FORMANT_CHUNKS": "1"}, {}, {"VBX_FORMANT_CHUNKS": "2"}, {"VBX_FORMANT_CHUNKS": "4"}):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "bench.py", "--config", "c3", "--steps", "10", "--warmup", "3", "--no-cpu"], env=e, capture_output=True, text=True)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        print(env, "ms/step", round(d["ms_per_step"], 3), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()}, "e2e", f'{d["e2e"]["value"]:.3e}')
    except Exception as ex:
        print(env, "FAILED", ex, out.stderr[-500:])
PY
cat gpurun_out/r2_run3_kernels.txt
