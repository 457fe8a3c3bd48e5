export SN_BUILD_KNOBS=1
for d in 0 8 16 32 48 64 128 208 1; do
  echo -n "KWC(tma) conv1 dbg=$d: "; SN_KWC=2 SN_HL_DBG=$d python tools/profile_layer.py conv1 64 2>&1 | tail -n 1
done
for d in 0 32 8; do
  echo -n "KWC(direct) conv1 dbg=$d: "; SN_KWC=2 SN_TMA_STORE=0 SN_HL_DBG=$d python tools/profile_layer.py conv1 64 2>&1 | tail -n 1
done
