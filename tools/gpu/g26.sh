export SN_BUILD_KNOBS=1
for d in 0 2 4 1 7; do
  echo -n "CTA2 conv5 dbg=$d: "; SN_CTA2=2 SN_HL_DBG=$d python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
done
for sb in 3 5 9; do
  echo -n "CTA2 conv5 sb=$sb: "; SN_CTA2=2 SN_CTA2_SB=$sb python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
done
echo -n "CTA2 conv5 sa=2 (sb 12): "; SN_CTA2=2 SN_CTA2_SA=2 python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
for d in 0 2 7; do
  echo -n "single conv5 dbg=$d: "; SN_CTA2=0 SN_HL_DBG=$d python tools/profile_layer.py conv5 64 2>&1 | tail -n 1
done
