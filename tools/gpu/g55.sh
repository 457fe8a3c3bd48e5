set -x
timeout 1500 python -m pytest tests/ -q -m gpu -x > gpurun_out/r02_t55.log 2>&1; tail -n 3 gpurun_out/r02_t55.log
cat > /tmp/fc_time.py <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
import supernet_b200 as S
F = S.fastops
B = 64
x = torch.rand(B, 204, 204, 4, device="cuda")
w = torch.randn(3, 3, 4, 32, device="cuda") * 0.1
ws = torch.full((32,), -5.0, device="cuda")
out = F.packed_empty(B, 202, 202, 32, "cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for gen1 in (False, True, False, True):
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); F.first_conv_packed(x, w, ws, F.PackedView(out), relu=True, gen1=gen1); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("gen1" if gen1 else "ws  ", "%.4f ms" % sorted(ts)[3])
PY
python /tmp/fc_time.py
SN_FIRST_BULK=0 python /tmp/fc_time.py
