set -x
timeout 1500 python -m pytest tests/ -q -m gpu > gpurun_out/r02_t12.log 2>&1; tail -n 6 gpurun_out/r02_t12.log
python tools/sanitize_kernels.py > gpurun_out/r02_sanitize_plain.log 2>&1; tail -n 2 gpurun_out/r02_sanitize_plain.log
for tool in memcheck synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_kernels.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -n 6 gpurun_out/r02_sanitizer_$tool.log
done
