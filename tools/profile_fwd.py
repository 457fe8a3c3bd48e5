"""One warm-up forward + one profiled forward of the FAST-mode BraTS engine (no CUDA graph), for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from supernet_b200.engine import InferenceEngine
from oracle import supernet_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w = O.make_weights("brats", 32, 4, 4)
model = S.Density_prop_with_pad_UNET(32, 4, variant="brats", mode="fast").load_weight_dict(w, device="cuda")
eng = InferenceEngine(model, B, 204, 204, 4, "cuda", graph=False)
eng.x_in.copy_(O.make_input("brats", B, alpha=O.BRATS_ALPHA))
for _ in range(2):
    eng.forward_resident()
    torch.cuda.synchronize()
print("ok", eng.n_launches, "launches per forward:", ",".join(eng.step_names))
