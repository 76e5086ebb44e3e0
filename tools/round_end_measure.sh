set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1d_smoke.log
timeout 300 python bench.py > gpurun_out/r1d_bench_default.log 2>&1; echo "bench rc=$?" >> gpurun_out/r1d_bench_default.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1d_bench_ref.log 2>&1; echo "ref rc=$?" >> gpurun_out/r1d_bench_ref.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/r1d_ncu_bench.log 2>&1; echo "ncu1 rc=$?" >> gpurun_out/r1d_ncu_bench.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:rec_ -s 2 -c 2 -o gpurun_out/prof_rec_r1d python tools/prof_rec.py > gpurun_out/r1d_ncu_rec.log 2>&1; echo "ncu2 rc=$?" >> gpurun_out/r1d_ncu_rec.log
timeout 120 python tools/rec_bench.py > gpurun_out/r1d_rec_bench.log 2>&1
timeout 120 python tools/gru_bench.py > gpurun_out/r1d_gru_bench.log 2>&1
B=256 STEPS=3 timeout 200 python tools/metaformer_bench.py > gpurun_out/r1d_meta_bench.log 2>&1
MODES=wavefront timeout 200 python tools/rollout_bench.py > gpurun_out/r1d_rollout_bench.log 2>&1
timeout 120 python tools/stream_bench.py > gpurun_out/r1d_stream_bench.log 2>&1
echo alldone
