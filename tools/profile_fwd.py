"""One warm-up forward + one profiled forward (cudaProfilerStart/Stop) of the FAST-mode BraTS engine, eager launches,
for `ncu --profile-from-start off`."""
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
eng.forward_resident()
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: only the second forward is profiled
eng.forward_resident()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", eng.n_launches, "launches per forward:", ",".join(eng.step_names))
