#!/bin/bash
# 8-GPU (or N-GPU) run: topology, H2D probe, the bench under torchrun (both arms)
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_${N}gpu.txt 2>&1
nproc >> gpurun_out/r2_topo_${N}gpu.txt; free -g >> gpurun_out/r2_topo_${N}gpu.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/r2_topo_${N}gpu.txt
python tools/h2d_probe.py 512 6 > gpurun_out/r2_h2d_probe_${N}gpu.txt 2>&1
VBX_MULTI_NO_BIND=1 python tools/h2d_probe.py 512 6 > gpurun_out/r2_h2d_probe_${N}gpu_unbound.txt 2>&1
tail -5 gpurun_out/r2_h2d_probe_${N}gpu.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench rc=$?"; tail -4 gpurun_out/r2_bench_${N}gpu.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2_bench_ref_${N}gpu.json 2> gpurun_out/r2_bench_ref_${N}gpu.err
echo "ref rc=$?"
