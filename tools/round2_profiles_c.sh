# Round-2 profiling set, third pass (tensor-core recurrent kernels of the reduced-precision modes, skinny head GEMMs): full
# capture of rec_fwd3 / rec_bwd3 at the cfg 4 layer shape, launch lists of the cfg 4 (bf16 mode) and default bench commands.
# Run under gpurun from the repository root; every profiled command first runs plainly.
set -x
cd $GRAFT_REPO_ROOT
MRG_PRECISION=bf16 B=256 timeout 100 python tools/prof_rec.py > gpurun_out/r2c_prof_plain_rec3.log 2>&1; echo "rec3 rc=$?"
MRG_PRECISION=bf16 B=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rec_ -s 2 -c 2 -f -o gpurun_out/prof_rec3_r2 python tools/prof_rec.py > gpurun_out/r2c_ncu_rec3.log 2>&1; echo "ncu1 rc=$?"
timeout 300 python bench.py --config 4 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2c_prof_plain_bench4.log 2>&1; echo "bench4 rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2c_cfg4.csv python bench.py --config 4 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2c_ncu_bench4.log 2>&1; echo "ncu2 rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-graph > gpurun_out/r2c_prof_plain_bench.log 2>&1; echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-graph > gpurun_out/r2c_ncu_bench.log 2>&1; echo "ncu3 rc=$?"
echo alldone
