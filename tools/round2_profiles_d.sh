# Round-2 profiling set, fourth pass (tensor-core recurrent kernels after the operand swap): full capture of rec_fwd3 / rec_bwd3
# at the cfg 4 layer shape (B=256) and at the headline batch (B=64), launch list of cfg 4 in bf16 mode.
# Run under gpurun from the repository root; every profiled command first runs plainly.
set -x
cd $GRAFT_REPO_ROOT
MRG_PRECISION=bf16 B=256 timeout 100 python tools/prof_rec.py > gpurun_out/r2d_prof_plain_rec3.log 2>&1; echo "rec3 rc=$?"
MRG_PRECISION=bf16 B=256 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rec_ -s 2 -c 2 -f -o gpurun_out/prof_rec3_r2d python tools/prof_rec.py > gpurun_out/r2d_ncu_rec3.log 2>&1; echo "ncu1 rc=$?"
MRG_PRECISION=tf32 B=64 timeout 100 python tools/prof_rec.py > gpurun_out/r2d_prof_plain_rec3_b64.log 2>&1; echo "rec3b rc=$?"
MRG_PRECISION=tf32 B=64 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rec_ -s 2 -c 2 -f -o gpurun_out/prof_rec3_r2d_b64 python tools/prof_rec.py > gpurun_out/r2d_ncu_rec3_b64.log 2>&1; echo "ncu1b rc=$?"
timeout 300 python bench.py --config 4 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2d_prof_plain_bench4.log 2>&1; echo "bench4 rc=$?"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2d_cfg4.csv python bench.py --config 4 --precision bf16 --steps 1 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/r2d_ncu_bench4.log 2>&1; echo "ncu2 rc=$?"
echo alldone
