set -x
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r02_t56.log 2>&1; tail -n 3 gpurun_out/r02_t56.log
python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; tail -n 1 gpurun_out/r02_smoke.log
python tools/profile_fwd.py 64 > gpurun_out/profile_fwd_plain.log 2>&1; echo "profile_fwd rc=$?"
ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_r02_fwd python tools/profile_fwd.py 64 > gpurun_out/ncu_fwd.log 2>&1
ncu -i /tmp/prof_r02_fwd.ncu-rep --page raw --csv > gpurun_out/prof_r02_fwd_raw.csv
python tools/capture_traffic.py gpurun_out/prof_r02_fwd_raw.csv 64 profiles/r02_traffic.json > gpurun_out/capture_traffic.log 2>&1; cp profiles/r02_traffic.json gpurun_out/r02_traffic.json
python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 3 --no-aux --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"], "clk", d["clocks"], "fgsm", d["aux"]["fgsm"]["slices_per_s"], "train", d["aux"]["train"]["slices_per_s"])
PY
