# Round-2 profiling set, second pass (after the tensor-core attention): launch list of the bench command and a full capture
# of the attention kernels.  Run under gpurun from the repository root.
set -x
cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-graph > gpurun_out/r2b_prof_plain_bench.log 2>&1; echo "bench rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library-baseline --no-graph > gpurun_out/r2b_ncu_bench.log 2>&1; echo "ncu1 rc=$?"
timeout 100 python tools/att_bench.py > gpurun_out/r2b_prof_plain_att.log 2>&1; echo "att rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_mma -s 6 -c 3 -f -o gpurun_out/prof_attn_r2 python tools/att_bench.py > gpurun_out/r2b_ncu_att.log 2>&1; echo "ncu2 rc=$?"
echo alldone
