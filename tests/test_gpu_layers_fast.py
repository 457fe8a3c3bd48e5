"""FAST mode behind the reference's per-layer call surface (VERDICT r1 "missing #1"): layer-by-layer
code in the style of Brats.py:379-455 -- including a U-Net that is NOT one of the two built-in graphs -- runs the
tcgen05 kernels and meets the north-star bars against the fp64 oracle."""
import pytest
import torch

from oracle import supernet_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import supernet_b200 as S_
    assert S_._lib.load().sn_device_check() == 0
    return S_


def rel(a, b):
    return O.rel_l2(a.detach().cpu(), b.detach().cpu())


class ThreeLevelUNet(torch.nn.Module):
    """A 3-level SUPER U-Net (the reference ships a 4-level BraTS and a 2-level Hippocampus one) wired from the layer
    classes exactly the way Density_prop_with_pad_UNET.call wires them (Brats.py:377-457)."""

    def __init__(self, S, n=32, n_labels=4, fill=0.05):
        super().__init__()
        hi = dict(sigma_min=-4.6, sigma_max=-2.2)
        self.conv_input = S.myConv_input(kernel_num=n, kernel_size=3)
        self.conv1 = S.myConv_intermediate(kernel_num=n, kernel_size=3)
        self.conv2 = S.myConv_intermediate(kernel_num=2 * n, kernel_size=3)
        self.conv3 = S.myConv_intermediate(kernel_num=2 * n, kernel_size=3)
        self.conv4 = S.myConv_intermediate(kernel_num=4 * n, kernel_size=3)
        self.conv5 = S.myConv_intermediate(kernel_num=4 * n, kernel_size=3)
        self.conv6 = S.myConv_intermediate(kernel_num=8 * n, kernel_size=3)
        self.conv7 = S.myConv_intermediate(kernel_num=8 * n, kernel_size=3)
        self.up1_conv2x2 = S.myConv_intermediate(kernel_num=4 * n, kernel_size=2, **hi)
        self.up1_conv1 = S.myConv_intermediate(kernel_num=4 * n, kernel_size=3)
        self.up1_conv2 = S.myConv_intermediate(kernel_num=4 * n, kernel_size=3)
        self.up2_conv2x2 = S.myConv_intermediate(kernel_num=2 * n, kernel_size=2, **hi)
        self.up2_conv1 = S.myConv_intermediate(kernel_num=2 * n, kernel_size=3)
        self.up2_conv2 = S.myConv_intermediate(kernel_num=2 * n, kernel_size=3)
        self.up3_conv2x2 = S.myConv_intermediate(kernel_num=n, kernel_size=2)
        self.up3_conv1 = S.myConv_intermediate(kernel_num=n, kernel_size=3)
        self.up3_conv2 = S.myConv_intermediate(kernel_num=n, kernel_size=3)
        self.conv_final = S.myConv_intermediate(kernel_num=n_labels, kernel_size=1, **hi)
        self.maxp, self.myups, self.myrelu = S.mymaxpooling(), S.myupsampling(), S.myReLU()
        self.myconc, self.mysoft = S.myConc(), S.mysoftmax()
        self.mypad = S.mypadding(pad_size=[2, 2], sigma_fill=fill)
        self.mypad_up6 = S.mypadding(pad_size=[3, 3], sigma_fill=fill)
        self.fill = fill

    def forward(self, x):
        # written one call per line like the reference: the SAME code runs fp32 tensors or FAST handles
        m, s = self.conv_input(x)
        m, s = self.myrelu(m, s)
        m, s = self.conv1(m, s)
        m1, s1 = self.myrelu(m, s)
        m, s = self.maxp(m1, s1)
        m, s = self.conv2(m, s)
        m, s = self.myrelu(m, s)
        m, s = self.conv3(m, s)
        m2, s2 = self.myrelu(m, s)
        m, s = self.maxp(m2, s2)
        m, s = self.conv4(m, s)
        m, s = self.myrelu(m, s)
        m, s = self.conv5(m, s)
        m3, s3 = self.myrelu(m, s)
        m, s = self.maxp(m3, s3)
        m, s = self.conv6(m, s)
        m, s = self.myrelu(m, s)
        m, s = self.conv7(m, s)
        m, s = self.myrelu(m, s)
        for up, c1, c2, (me, se) in ((self.up1_conv2x2, self.up1_conv1, self.up1_conv2, (m3, s3)),
                                     (self.up2_conv2x2, self.up2_conv1, self.up2_conv2, (m2, s2)),
                                     (self.up3_conv2x2, self.up3_conv1, self.up3_conv2, (m1, s1))):
            m, s = self.myups(m, s)
            m, s = up(m, s)
            m, s = self.mypad_up6(m, s)
            m, s = self.myconc(m, s, me, se)
            m, s = c1(m, s)
            m, s = self.myrelu(m, s)
            m, s = self.mypad(m, s)
            m, s = c2(m, s)
            m, s = self.myrelu(m, s)
        mf, sf = self.conv_final(m, s)
        p, v = self.mysoft(mf, sf)
        return p, v, mf, sf

    def oracle(self, x):
        """The same graph through the oracle's layer functions, fp64, same weights."""
        W = {n: tuple(t.detach().cpu().double() for t in l.weights()) for n, l in self.named_children()
             if hasattr(l, "weights")}
        c = O.conv_intermediate_conv_form
        m, s = O.relu(*O.conv_input_conv_form(x.double(), *W["conv_input"]))
        m1, s1 = O.relu(*c(m, s, *W["conv1"]))
        m, s = O.maxpooling(m1, s1)
        m, s = O.relu(*c(m, s, *W["conv2"]))
        m2, s2 = O.relu(*c(m, s, *W["conv3"]))
        m, s = O.maxpooling(m2, s2)
        m, s = O.relu(*c(m, s, *W["conv4"]))
        m3, s3 = O.relu(*c(m, s, *W["conv5"]))
        m, s = O.maxpooling(m3, s3)
        m, s = O.relu(*c(m, s, *W["conv6"]))
        m, s = O.relu(*c(m, s, *W["conv7"]))
        for d, (me, se) in ((1, (m3, s3)), (2, (m2, s2)), (3, (m1, s1))):
            m, s = O.upsampling(m, s)
            m, s = c(m, s, *W[f"up{d}_conv2x2"])
            m, s = O.padding(m, s, (3, 3), self.fill)
            m, s = O.conc(m, s, me, se)
            m, s = O.relu(*c(m, s, *W[f"up{d}_conv1"]))
            m, s = O.padding(m, s, (2, 2), self.fill)
            m, s = O.relu(*c(m, s, *W[f"up{d}_conv2"]))
        mf, sf = c(m, s, *W["conv_final"])
        p, v = O.softmax_as_written(mf, sf)
        return p, v, mf, sf


def test_three_level_unet_from_layer_classes_runs_on_tensor_cores(S):
    torch.manual_seed(11)
    net = ThreeLevelUNet(S)
    x = torch.rand(2, 100, 100, 4) / 256.0
    with S.fast_mode():
        m, s = net.conv_input(x.cuda())                 # weights are created at the first call, like Keras' build()
        assert isinstance(m, S.PackedMoments) and m is s and m.shape == (2, 98, 98, 32)
        with torch.no_grad():
            p, v, mf, sf = net(x.cuda())
    assert isinstance(mf, S.fastlayers.PendingHead) and p.shape == (2, 82 * 82, 4) and p.dtype == torch.float32
    pre_m, pre_v = mf.unpack()
    p_ref, v_ref, mf_ref, sf_ref = net.oracle(x)
    errs = dict(pre_mu=rel(pre_m, mf_ref), pre_var=rel(pre_v, sf_ref), p=rel(p, p_ref), v=rel(v, v_ref),
                argmax=O.argmax_agreement(p.cpu(), p_ref))
    print(errs)
    assert errs["pre_mu"] < 1e-3 and errs["p"] < 1e-3, errs
    assert errs["pre_var"] < 1e-2 and errs["v"] < 1e-2, errs
    assert errs["argmax"] >= 0.999, errs
    assert float(v.min()) >= 0.0 and bool(torch.isfinite(v).all())
    # the same module, same weights, through the FP32-mode kernels: the per-layer surface is one and the same
    with torch.no_grad():
        p32, v32, mf32, sf32 = net(x.cuda())
    assert torch.is_tensor(mf32) and rel(p32, p_ref) < 1e-4 and rel(v32, v_ref) < 1e-4


@pytest.mark.parametrize("variant,C,in_ch,B,alpha", [("hippocampus", 3, 1, 3, 1.0), ("brats", 4, 4, 1, O.BRATS_ALPHA)])
def test_layerwise_fast_forward_equals_the_engine(S, variant, C, in_ch, B, alpha):
    """The built-in graphs, called layer by layer with FAST handles, launch the same kernels with the same flags as
    engine.InferenceEngine: bit-identical outputs."""
    w32 = O.make_weights(variant, 32, C, in_ch)
    model = S.Density_prop_with_pad_UNET(32, C, variant=variant, mode="fast").load_weight_dict(w32, device="cuda")
    x = O.make_input(variant, B, alpha=alpha).cuda()
    with torch.no_grad():
        p, v, mf, sf = model(x, return_presoftmax=True)
    calls = []
    orig = S.fastops.conv_moments_tc_head
    S.fastops.conv_moments_tc_head = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    try:
        p2, v2, mf2, sf2 = model.forward_layerwise_fast(x, return_presoftmax=True)
    finally:
        S.fastops.conv_moments_tc_head = orig
    assert torch.equal(p, p2) and torch.equal(v, v2) and torch.equal(mf, mf2) and torch.equal(sf, sf2)
    # ... including the last launch: the pending 3x3 conv + ReLU + conv_final + softmax as ONE kernel
    assert len(calls) >= 1


def test_fast_handles_general_paths(S):
    """Compositions outside the fused patterns still compute the reference's semantics: ReLU after a pool
    (sn_relu_packed), padding a tensor that is already in memory (window copy), up-sampling consumed by something
    other than a 2x2 conv (dense zero-stuffing), a handle kept aside while its successor is gated."""
    F = S.fastops
    FL = S.fastlayers
    g = torch.Generator().manual_seed(2)
    mu = torch.randn(2, 9, 11, 32, generator=g)
    var = torch.rand(2, 9, 11, 32, generator=g)
    h = S.PackedMoments((2, 9, 11, 32), "cuda", view=F.PackedView(F.pack_moments(mu.cuda(), var.cuda())))
    m0, v0 = h.unpack()                                      # bf16-rounded values: the reference point for exactness
    relu, pad, pool, ups = S.myReLU(), S.mypadding(pad_size=[1, 2], sigma_fill=0.3), S.mymaxpooling(), S.myupsampling()
    # ReLU on a materialised tensor
    r, r_ = relu(h, h)
    mr, vr = r.unpack()
    m_ref, v_ref = O.relu(m0.cpu(), v0.cpu())
    assert torch.equal(mr.cpu(), m_ref) and torch.equal(vr.cpu(), v_ref)
    assert torch.equal(h.unpack()[0], m0)                    # the original handle is untouched
    # pool -> ReLU -> pad (pool output pending, ReLU not fusable into a pool, pad of the gated tensor)
    q, q_ = pad(*relu(*pool(h, h)))
    mq, vq = q.unpack()
    pm, pv = O.padding(*O.relu(*O.maxpooling(m0.cpu(), v0.cpu())), (1, 2), 0.3)
    assert mq.shape == pm.shape and torch.equal(mq.cpu(), pm)
    assert torch.equal(vq.cpu(), pv.bfloat16().float())
    # up-sampling consumed by a pad: dense zero-stuffed tensor
    u, u_ = pad(*ups(h, h))
    mu_u, var_u = u.unpack()
    um, uv = O.padding(*O.upsampling(m0.cpu(), v0.cpu()), (1, 2), 0.3)
    assert torch.equal(mu_u.cpu(), um) and torch.equal(var_u.cpu(), uv.bfloat16().float())
    # a pending conv kept aside (skip) and its gated successor are two different tensors
    conv = S.myConv_intermediate(kernel_num=32, kernel_size=3)
    c, c_ = conv(h, h)
    cr, cr_ = relu(c, c)
    mc, vc = c.unpack()
    mcr, vcr = cr.unpack()
    assert bool((mc < 0).any()) and float(mcr.min()) >= 0.0
    assert torch.equal(mcr, torch.relu(mc)) and torch.equal(vcr, vc * (mc > 0))
    w, ws = (t.detach().cpu().double() for t in conv.weights())
    mo, vo = O.conv_intermediate_conv_form(m0.cpu().double(), v0.cpu().double(), w, ws)
    assert rel(mc, mo) < 1e-4 and rel(vc, vo) < 5e-3
    # misuse is an error, not a silent fallback
    with pytest.raises(RuntimeError):
        conv(h, m0)                                          # the pair was taken apart
    with pytest.raises(RuntimeError):
        S.myConv_intermediate(kernel_num=48, kernel_size=3)(h, h)
    with pytest.raises(RuntimeError):
        FL.conv_input(S.myConv_input(kernel_num=32, kernel_size=3, in_channels=4), torch.rand(1, 8, 8, 4))   # CPU tensor
