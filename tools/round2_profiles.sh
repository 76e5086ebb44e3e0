# Round-2 profiling set (run under gpurun from the repository root; every profiled command first runs plainly).
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-graph > gpurun_out/r2_prof_plain_bench.log 2>&1; echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-graph > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu1 rc=$?"
timeout 100 python tools/prof_rec.py > gpurun_out/r2_prof_plain_rec.log 2>&1; echo "rec rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"rec_|gemm_tc" -s 12 -c 6 -f -o gpurun_out/prof_rec_r2 python tools/prof_rec.py > gpurun_out/r2_ncu_rec.log 2>&1; echo "ncu2 rc=$?"
T=900 timeout 100 python tools/rollout_ncu.py > gpurun_out/r2_prof_plain_rollout.log 2>&1; echo "rollout rc=$?"
T=900 timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout -c 2 -f -o gpurun_out/prof_rollout_r2 python tools/rollout_ncu.py > gpurun_out/r2_ncu_rollout.log 2>&1; echo "ncu3 rc=$?"
timeout 300 python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_prof_plain_bench3.log 2>&1; echo "bench3 rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2_cfg3.csv python bench.py --config 3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_bench3.log 2>&1; echo "ncu4 rc=$?"
echo alldone
