"""Where does the FAST-mode input-gradient error come from?  (BraTS, batch 1, create_adversarial_pattern's loss.)

Three gradients of the same loss w.r.t. the same input:
  G_oracle  fp64 oracle autograd (the reference's tf.GradientTape chain, Brats.py:582-596)
  G_A       fp64 autograd run ON THE FAST FORWARD'S TRAJECTORY: every conv output is replaced by the value the FAST
            engine stored (straight-through), and every ReLU gate / arg-max decision is the FAST engine's; the
            backward arithmetic itself is exact.  G_A - G_oracle = what the forward's ~1e-5 activation error costs
            through flipped gates and re-routed pooling windows.
  G_fast    the tensor-core data-gradient chain.  G_fast - G_A = rounding of the gradient planes / W^2 operands.
usage: grad_error_split.py [seed ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import supernet_b200 as S
from oracle import supernet_oracle as O

C = 4
seeds = [int(a) for a in sys.argv[1:]] or [2025, 2026]
W32 = O.make_weights("brats", 32, C, 4)
W64 = {k: (a.double(), b.double()) for k, (a, b) in W32.items()}
model = S.Density_prop_with_pad_UNET(32, C, variant="brats", mode="fast").load_weight_dict(W32, device="cuda")


def fast_activations(eng):
    acts = {}
    for r in eng.records:
        if r["kind"] == "first":
            name, v = "conv_input", r["dst"]
            oh, ow, c = v.buf.shape[1], v.buf.shape[2], v.buf.shape[4]
        elif r["kind"] == "conv":
            name, v, c = r["name"], r["dst"], r["cout"]
            oh, ow = (2 * r["h"], 2 * r["w"]) if r["upconv"] else (r["h"] - r["k"] + 1, r["w"] - r["k"] + 1)
        else:
            continue
        t = v.buf[:, v.y0:v.y0 + oh, v.x0:v.x0 + ow, :, v.c0:v.c0 + c].float().cpu().double()
        acts[name] = (t[..., 0, :] + t[..., 1, :], t[..., 2, :])
    return acts


def forward(x, acts=None):
    """Brats.py:377-457 with the oracle's layer functions; acts != None: the FAST trajectory (see the docstring)."""
    fill = 0.1

    def post(name, m, s, has_relu):
        if acts is None:
            return O.relu(m, s) if has_relu else (m, s)
        mf, sf = acts[name]
        if has_relu:
            gate = (mf > 0).double()
            m, s = m * gate, s * gate
        return m + (mf - m).detach(), s + (sf - s).detach()

    m, s = post("conv_input", *O.conv_input_conv_form(x, *W64["conv_input"]), True)
    m, s = post("conv1", *O.conv_intermediate_conv_form(m, s, *W64["conv1"]), True)
    skips = [(m, s)]
    ci = 2
    for lvl in range(1, 5):
        m, s = O.maxpooling(m, s)
        if lvl == 4:
            m, s = O.padding(m, s, (1, 0), fill)
        for _ in range(2):
            m, s = post(f"conv{ci}", *O.conv_intermediate_conv_form(m, s, *W64[f"conv{ci}"]), True)
            ci += 1
        if lvl < 4:
            skips.append((m, s))
    for d in range(1, 5):
        me, se = skips[4 - d]
        m, s = O.upsampling(m, s)
        m, s = post(f"up{d}_conv2x2", *O.conv_intermediate_conv_form(m, s, *W64[f"up{d}_conv2x2"]), False)
        m, s = O.padding(m, s, (3, 3), fill)
        m, s = O.conc(m, s, me, se)
        m, s = post(f"up{d}_conv1", *O.conv_intermediate_conv_form(m, s, *W64[f"up{d}_conv1"]), True)
        m, s = O.padding(m, s, (2, 2), fill)
        m, s = post(f"up{d}_conv2", *O.conv_intermediate_conv_form(m, s, *W64[f"up{d}_conv2"]), True)
    mf, sf = O.conv_intermediate_conv_form(m, s, *W64["conv_final"])
    return O.softmax_closed_form(mf, sf)


def grad(x, y, acts=None):
    xr = x.double().clone().requires_grad_(True)
    p, v = forward(xr, acts)
    loss = 0.5 * O.nll_gaussian(y, p, torch.clamp(v, -1e4, 1e3))
    (g,) = torch.autograd.grad(loss, xr)
    return g, float(loss)


def cmp(a, b):
    big = b.abs() > 1e-3 * b.abs().max()
    return O.rel_l2(a, b), float((torch.sign(a)[big] == torch.sign(b)[big]).double().mean())


print("| input seed | G_fast vs G_oracle | G_A vs G_oracle (gates / arg-max of the FAST forward) | G_fast vs G_A (gradient-plane rounding) |")
print("|---|---|---|---|")
for seed in seeds:
    x = O.make_input("brats", 1, seed=seed, alpha=O.BRATS_ALPHA)
    y = O.make_labels(1, 186 * 186, C, seed=7).double()
    eng = model.grad_engine_for(x.cuda())
    _, g_fast = eng.input_gradient(x.cuda(), y.float().cuda())
    g_fast = g_fast.cpu().double()
    acts = fast_activations(eng)
    g_or, l_or = grad(x, y)
    g_a, l_a = grad(x, y, acts)
    t, a, b = cmp(g_fast, g_or), cmp(g_a, g_or), cmp(g_fast, g_a)
    print(f"| {seed} | {t[0]:.2e} (sign {t[1]:.4f}) | {a[0]:.2e} (sign {a[1]:.4f}) | {b[0]:.2e} (sign {b[1]:.4f}) |", flush=True)
