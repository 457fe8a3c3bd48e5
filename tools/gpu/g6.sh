for d in 0 1 2 4 8 3 5 6 7 13; do
  echo -n "KWC dbg=$d: "; SN_HL_DBG=$d python tools/profile_layer.py conv1 64 2>&1 | tail -n 1
done
for d in 0 1 2 4 7; do
  echo -n "noKWC dbg=$d: "; SN_KWC=0 SN_HL_DBG=$d python tools/profile_layer.py conv1 64 2>&1 | tail -n 1
done
for d in 0 1 2 4 7; do
  echo -n "conv3 dbg=$d: "; SN_HL_DBG=$d python tools/profile_layer.py conv3 64 2>&1 | tail -n 1
done
