export SN_BUILD_KNOBS=1
for d in 0 7; do
  echo -n "single conv5 dbg=$d: "; SN_CTA2=0 SN_HL_DBG=$d python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
  echo -n "single conv5 fake cluster dbg=$d: "; SN_FAKE_CLUSTER=1 SN_CTA2=0 SN_HL_DBG=$d python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
done
