"""Batch sweep 1-512 (BASELINE.json configs[0] and configs[4]): one CUDA-graph forward of
FAST mode at small batches for both networks, next to the CPU oracle (conv form = best-case CPU, as-written form =
the reference's op sequence) on the host cores."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from supernet_b200.engine import InferenceEngine
from oracle import supernet_oracle as O

out = []
for variant, C, in_ch, hw, batches in (("hippocampus", 3, 1, 64, (1, 8, 64, 512)),
                                       ("brats", 4, 4, 204, (1, 2, 4, 8, 16, 32, 64, 128, 256, 512))):
    w = O.make_weights(variant, 32, C, in_ch)
    model = S.Density_prop_with_pad_UNET(32, C, variant=variant, mode="fast").load_weight_dict(w, device="cuda")
    for B in batches:
        eng = InferenceEngine(model, B, hw, hw, in_ch, "cuda", graph=True, keep_presoftmax=False)
        alpha = O.BRATS_ALPHA if variant == "brats" else 1.0
        eng.x_in.copy_(O.make_input(variant, B, alpha=alpha))
        for _ in range(5):
            eng.forward_resident()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            eng.forward_resident()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        row = {"variant": variant, "batch": B, "gpu_ms": round(ms, 4), "gpu_slices_per_s": round(B / ms * 1e3, 1),
               "launches": eng.n_launches}
        if (variant, B) in (("hippocampus", 8), ("brats", 1)):
            torch.set_num_threads(os.cpu_count())
            for form in ("conv", "as_written"):
                orc = O.UNetOracle(variant, 32, C, in_ch, torch.float32, form=form)
                x = O.make_input(variant, B, alpha=alpha)
                with torch.no_grad():
                    orc(x)
                    t0 = time.perf_counter()
                    n = 5
                    for _ in range(n):
                        orc(x)
                    row[f"cpu_{form}_ms"] = round((time.perf_counter() - t0) / n * 1e3, 2)
            row["cpu_cores"] = os.cpu_count()
        out.append(row)
        print(json.dumps(row), flush=True)
        del eng
        torch.cuda.empty_cache()
