#!/bin/bash
# N-GPU run of both bench arms under torchrun, as the driver launches them: tools/r2_runN.sh N
N=${1:-8}
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2f_bench_${N}gpu.json 2> gpurun_out/r2f_bench_${N}gpu.err
echo "bench rc=$?"; tail -4 gpurun_out/r2f_bench_${N}gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2f_bench_ref_${N}gpu.json 2> gpurun_out/r2f_bench_ref_${N}gpu.err
echo "ref rc=$?"; tail -4 gpurun_out/r2f_bench_ref_${N}gpu.err
python tools/show_bench.py gpurun_out/r2f_bench_${N}gpu.json 2>/dev/null | head -3 | cut -c1-260
