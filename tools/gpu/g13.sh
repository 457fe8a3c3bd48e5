set -x
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc_bwd.py -q -k "writes_only" > gpurun_out/r02_t13.log 2>&1; tail -n 4 gpurun_out/r02_t13.log
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02_bench13.json
ncu --set full --clock-control none --profile-from-start off -f -o gpurun_out/prof_r02_fwd python tools/profile_fwd.py 64 > gpurun_out/ncu_fwd.log 2>&1; tail -n 2 gpurun_out/ncu_fwd.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -n 1 gpurun_out/ncu_bench.log | cut -c1-300
