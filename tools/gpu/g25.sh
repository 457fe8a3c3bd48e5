SN_CTA2=2 ncu --set full --clock-control none --import-source on -k regex:conv_moments_halo -s 2 -c 1 -f -o /tmp/r02_conv5_cta2 python tools/profile_layer.py conv5 64 > gpurun_out/ncu_a.log 2>&1; tail -n 2 gpurun_out/ncu_a.log
ncu -i /tmp/r02_conv5_cta2.ncu-rep --page source --csv > gpurun_out/r02_conv5_cta2_source.csv
ncu -i /tmp/r02_conv5_cta2.ncu-rep --page raw --csv > gpurun_out/r02_conv5_cta2_raw.csv
ls -la gpurun_out/
