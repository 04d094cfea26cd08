set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
python __graft_entry__.py smoke > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
for c in c2 c3 c4 c5; do python bench.py --config $c > gpurun_out/final_bench_$c.json 2> gpurun_out/final_bench_$c.err; done
python bench.py --impl reference > gpurun_out/final_bench_ref_c2.json 2>/dev/null
python tools/show_bench.py gpurun_out/final_bench_c?.json | grep -v clocks
bash tools/profile_all.sh > gpurun_out/profile_all.log 2>&1
python tools/parity_scale.py 32 > gpurun_out/final_parity.txt 2>&1; tail -3 gpurun_out/final_parity.txt
# summaries on the box; the reports themselves are too large to bring back (64 MiB limit) except the two changed kernels
for k in lpc16 lpc roots tracker lag refine mfcc; do
  bash tools/ncu_summary.sh gpurun_out/prof_${k}_final.ncu-rep gpurun_out/r1_${k}_final_full.txt
done
mkdir -p /tmp/keep && mv gpurun_out/prof_mfcc_final.ncu-rep gpurun_out/prof_lpc16_final.ncu-rep /tmp/keep/ && rm -f gpurun_out/*.ncu-rep && mv /tmp/keep/*.ncu-rep gpurun_out/
du -sh gpurun_out
