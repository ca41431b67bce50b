set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r1.log; tail -3 gpurun_out/pytest_gpu_r1.log
python bench.py > gpurun_out/bench_r1.json 2> gpurun_out/bench_r1.err; echo "bench rc=$?"; cat gpurun_out/bench_r1.json | head -c 400
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/bench_r1_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
CCB_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_r1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench_r1.log 2>&1; echo "ncu bench rc=$?"
CCB_PROFILE_LAST=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_gptj_r1.csv python tools/quick_gptj.py 16 4 > gpurun_out/ncu_gptj_r1.log 2>&1; echo "ncu gptj rc=$?"
