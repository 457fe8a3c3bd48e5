set -x
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r02_t46.log 2>&1; tail -n 4 gpurun_out/r02_t46.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -n 1 gpurun_out/r02_smoke.log
python tools/profile_fwd.py 64 > gpurun_out/profile_fwd_plain.log 2>&1; echo "profile_fwd rc=$?"
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_r02_fwd python tools/profile_fwd.py 64 > gpurun_out/ncu_fwd.log 2>&1
ncu -i /tmp/prof_r02_fwd.ncu-rep --page raw --csv > gpurun_out/prof_r02_fwd_raw.csv
python tools/capture_traffic.py gpurun_out/prof_r02_fwd_raw.csv 64 profiles/r02_traffic.json > gpurun_out/capture_traffic.log 2>&1; cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/profile_bwd.py 64 train > gpurun_out/profile_bwd_plain.log 2>&1; echo "profile_bwd rc=$?"
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none --profile-from-start off -f -o /tmp/prof_r02_train python tools/profile_bwd.py 64 train > gpurun_out/ncu_train.log 2>&1
ncu -i /tmp/prof_r02_train.ncu-rep --page raw --csv > gpurun_out/prof_r02_train_raw.csv
python tools/bench_sweep.py > gpurun_out/r02_sweep_n1.jsonl 2> gpurun_out/r02_sweep_n1.err
python tools/bench_fgsm_fast.py --batch 64 > gpurun_out/r02_fgsm_fast_b64.json 2> gpurun_out/r02_fgsm_fast_b64.err
python tools/bench_fgsm_fast.py --train --batch 64 > gpurun_out/r02_train_fast_b64.json 2> gpurun_out/r02_train_fast_b64.err
ls -la gpurun_out | head -40
