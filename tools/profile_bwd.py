"""One warm-up + one profiled FAST-mode training chain (forward + NLL + data gradients + weight gradients) of the
BraTS GradientEngine, eager launches, bracketed by cudaProfilerStart/Stop for `ncu --profile-from-start off`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from supernet_b200.engine import GradientEngine
from oracle import supernet_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
train = (sys.argv[2] if len(sys.argv) > 2 else "train") == "train"
w = O.make_weights("brats", 32, 4, 4)
model = S.Density_prop_with_pad_UNET(32, 4, variant="brats", mode="fast").load_weight_dict(w, device="cuda")
eng = GradientEngine(model, B, 204, 204, 4, "cuda", graph=False, train=train)
if train:
    eng._loss_scale, eng._clip = 1.0, (1e-12, 1e3)
eng.x_in.copy_(O.make_input("brats", B, alpha=O.BRATS_ALPHA))
eng.y_in.copy_(O.make_labels(B, 186 * 186, 4))
eng.loss_and_input_gradient_resident()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.loss_and_input_gradient_resident()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
# launch order: forward steps, the two NLL kernels, then the backward steps (a tensor-core weight gradient is a memset +
# 3 kernels: GEMM, dsigma, finalize; the thin-layer ones a memset + 2 kernels)
names = list(eng.step_names) + ["nll_fwd", "nll_finalize"]
for n in eng.bwd_step_names:
    if n.endswith("_wgrad"):
        names += [n, n + ":dsigma", n + ":finalize"] if n not in ("conv_input_wgrad", "conv_final_wgrad") else [n, n + ":finalize"]
    else:
        names.append(n)
print("names=" + ",".join(names))
