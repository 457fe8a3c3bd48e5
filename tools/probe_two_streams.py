"""Probe: does replaying two resident engines (two CUDA graphs) on two streams overlap kernel tails?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from supernet_b200.engine import InferenceEngine
from oracle import supernet_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
w = O.make_weights("brats", 32, 4, 4)
model = S.Density_prop_with_pad_UNET(32, 4, variant="brats", mode="fast").load_weight_dict(w, device="cuda")
x = O.make_input("brats", B, alpha=O.BRATS_ALPHA).cuda()
for n_eng in (1, 2, 3):
    engs = [InferenceEngine(model, B, 204, 204, 4, "cuda", graph=True, keep_presoftmax=False) for _ in range(n_eng)]
    streams = [torch.cuda.Stream() for _ in range(n_eng)]
    for e, st in zip(engs, streams):
        e.x_in.copy_(x)
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            for _ in range(3):
                e.forward_resident()
    torch.cuda.synchronize()
    steps = 30
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for st in streams:
        st.wait_stream(torch.cuda.current_stream())
    for i in range(steps):
        with torch.cuda.stream(streams[i % n_eng]):
            engs[i % n_eng].forward_resident()
    for st in streams:
        torch.cuda.current_stream().wait_stream(st)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    print(json.dumps({"engines": n_eng, "batch": B, "ms_per_step": round(ms, 4), "slices_per_s": round(B / ms * 1e3, 1)}))
    del engs
    torch.cuda.empty_cache()
