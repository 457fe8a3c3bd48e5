set -x
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:first_conv_tc -f -o /tmp/fc python tools/profile_fwd.py 64 > gpurun_out/ncu_fc.log 2>&1; tail -n 2 gpurun_out/ncu_fc.log
ncu -i /tmp/fc.ncu-rep --page source --csv > gpurun_out/r02_fc_source.csv
ncu -i /tmp/fc.ncu-rep --page raw --csv > gpurun_out/r02_fc_raw.csv
ls -la gpurun_out/r02_fc_*
