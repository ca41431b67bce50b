# Final-tree evidence for the GPT-J fused out_proj + fc_out GEMM (run under gpurun, one GPU):
#   default bench line, config-5 launch list, one `--set full` capture of six consecutive weight-streaming GEMM launches deep in
#   the decode loop (q/k/v, fc_in, fused out GEMM x 2 layers), memcheck of the tiny GPT-J fixture tests.
# NOTE (measured): the config-5 launch list alone takes ~14 min of box time -- ncu serialises 4 819 launches of a 12 GB model at ~170 ms each --
# so give this script a 40-minute limit; in round 2 the `--set full` step below was cut off by the GPU budget and never produced a report.
set -x
python bench.py > gpurun_out/r2_bench_final5.json 2> gpurun_out/r2_bench_final5.err; echo rc=$?
CCB_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_config5_fused.csv python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_config5_fused.log 2>&1; echo rc=$?
CCB_BENCH_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 300 -c 6 -o gpurun_out/r2_gptj_gemms -f python bench.py --config 5 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gptj_gemms.log 2>&1; echo rc=$?
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "gptj and not two_launch and not cta_per_unit" -p no:cacheprovider > gpurun_out/memcheck_gptj.log 2>&1; echo memcheck rc=$?; tail -5 gpurun_out/memcheck_gptj.log
