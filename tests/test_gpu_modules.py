"""Stand-alone `forward()` of every class / function of SURVEY.md 8b's "Python surface to keep" on reference-layout
(NCHW / NCDHW fp32) tensors, against the CPU oracle with the same seeded weights.  The fused pipeline never calls these
wrappers (it chains `_forward_cl`), so they get their own parity tests.  `pytest -m gpu`."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = 2e-4      # relative to the abs-max of the reference tensor (split-bf16 products: observed <= 5e-5)


def rel(got, ref):
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


@pytest.fixture(scope="module")
def gb():
    import __graft_entry__ as entry
    entry.build()
    torch.set_num_threads(os.cpu_count() or 1)
    G, sd = entry.load_seeded_gbase("cuda")
    return G, sd


def run(mod, *xs):
    with torch.no_grad():
        out = mod(*[x.cuda() for x in xs])
    return out.cpu()


def test_conv2d_ws_forward(gb):
    """Conv2d_WS.forward (model.py:61-69)."""
    import gbase_oracle as O
    G, sd = gb
    x = rnd(2, 64, 32, 48, seed=1)
    got = run(G.appearanceEncoder.resblock_128.conv_ws, x)
    assert rel(got, O.conv2d_ws(x, sd, "appearanceEncoder.resblock_128.conv_ws")) <= TOL


def test_conv3d_ws_forward():
    """Conv3D_WS.forward (model.py:71-86): no instance inside Gbase, so a fresh module with seeded weights."""
    from megaportrait_hack_b200 import model
    m = model.Conv3D_WS(16, 32, 3, padding=1)
    with torch.no_grad():
        m.weight.copy_(rnd(32, 16, 3, 3, 3, seed=2, scale=0.1) + 0.02)
        m.bias.copy_(rnd(32, seed=3, scale=0.1))
    x = rnd(2, 16, 4, 16, 16, seed=4)
    w = m.weight.detach()
    w = w - w.mean(dim=(1, 2, 3, 4), keepdim=True)
    w = w / (w.reshape(32, -1).std(dim=1).view(-1, 1, 1, 1, 1) + 1e-5)
    ref = F.conv3d(x, w, m.bias.detach(), padding=1)
    got = run(m.cuda(), x)
    assert got.shape == ref.shape and rel(got, ref) <= TOL


def test_adaptive_group_norm_forward(gb):
    """AdaptiveGroupNorm.forward (model.py:314-316)."""
    import gbase_oracle as O
    G, sd = gb
    x = rnd(2, 96, 3, 8, 8, seed=5, scale=2.0) + 0.5
    got = run(G.appearanceEncoder.resblock3D_96.norm1, x)
    assert rel(got, O.adaptive_group_norm(x, sd, "appearanceEncoder.resblock3D_96.norm1")) <= 1e-5


def test_resblock_custom_forward(gb):
    """ResBlock_Custom.forward, dimension 2 (model.py:110-130), incl. its output-shape asserts."""
    import gbase_oracle as O
    G, sd = gb
    x = rnd(1, 64, 64, 64, seed=6)
    got = run(G.appearanceEncoder.resblock_128, x)
    ref = O.resblock_custom2d(x, sd, "appearanceEncoder.resblock_128")
    assert got.shape == (1, 128, 64, 64) and rel(got, ref) <= TOL


def test_resblock3d_adaptive_forward(gb):
    """ResBlock3D_Adaptive.forward (model.py:385-408): identity residual (Eapp) and 1x1x1 residual conv (FlowField)."""
    import gbase_oracle as O
    G, sd = gb
    x = rnd(1, 96, 4, 16, 16, seed=7)
    got = run(G.appearanceEncoder.resblock3D_96, x)
    assert rel(got, O.resblock3d_adaptive(x, sd, "appearanceEncoder.resblock3D_96")) <= TOL
    x = rnd(2, 128, 16, 4, 4, seed=8)
    got = run(G.warp_generator_c2d.flowfield.resblock3, x)
    ref = O.resblock3d_adaptive(x, sd, "warp_generator_c2d.flowfield.resblock3")
    assert got.shape == (2, 64, 16, 4, 4) and rel(got, ref) <= TOL


def test_resblock3d_forward(gb):
    """ResBlock3D.forward (model.py:512-528): identity shortcut and 1x1x1 shortcut."""
    import gbase_oracle as O
    G, sd = gb
    x = rnd(1, 96, 4, 16, 16, seed=9)
    assert rel(run(G.G3d.downsampling[0], x), O.resblock3d(x, sd, "G3d.downsampling.0")) <= TOL
    got = run(G.G3d.downsampling[2], x)
    ref = O.resblock3d(x, sd, "G3d.downsampling.2")
    assert got.shape == (1, 192, 4, 16, 16) and rel(got, ref) <= TOL


def test_resblock3d_upsample_branch(gb):
    """`upsample=True` of ResBlock3D / ResBlock3D_Adaptive (model.py:405-406, 525-526; no call site in the reference, VERDICT
    round 1 "missing" #8): the block followed by F.interpolate(scale_factors, trilinear, align_corners=False)."""
    import gbase_oracle as O
    from megaportrait_hack_b200 import model
    G, sd = gb
    x = rnd(1, 96, 4, 16, 16, seed=21)
    for cls, src, prefix, fn in ((model.ResBlock3D, G.G3d.downsampling[0], "G3d.downsampling.0", O.resblock3d),
                                 (model.ResBlock3D_Adaptive, G.appearanceEncoder.resblock3D_96, "appearanceEncoder.resblock3D_96",
                                  O.resblock3d_adaptive)):
        blk = cls(96, 96, upsample=True, scale_factors=(1, 2, 2)).cuda().eval()
        blk.load_state_dict(src.state_dict())
        got = run(blk, x)
        ref = F.interpolate(fn(x, sd, prefix), scale_factor=(1, 2, 2), mode="trilinear", align_corners=False)
        assert got.shape == (1, 96, 4, 32, 32) and rel(got, ref) <= TOL


def test_resblock2d_forward(gb):
    """ResBlock2D.forward (model.py:621-640): an identity block of G2d's chain (called on its own it takes split-bf16
    planes, not the chain's F16_Q8 planes) and an up-block with the Conv1x1 + BatchNorm shortcut."""
    import gbase_oracle as O
    G, sd = gb
    x = rnd(2, 512, 16, 16, seed=10)
    assert rel(run(G.G2d.res_blocks[3], x), O.resblock2d(x, sd, "G2d.res_blocks.3")) <= TOL
    got = run(G.G2d.upsample1[1], x)
    ref = O.resblock2d(x, sd, "G2d.upsample1.1")
    assert got.shape == (2, 256, 16, 16) and rel(got, ref) <= TOL


def test_flowfield_forward(gb):
    """FlowField.forward(zs, adaptive_gamma, adaptive_beta) (model.py:439-471); the two extra arguments are ignored,
    as in the reference."""
    import gbase_oracle as O
    G, sd = gb
    zs = rnd(3, 512, 1, 1, seed=11)
    ff = G.warp_generator_s2c.flowfield
    with torch.no_grad():
        got = ff(zs.cuda(), None, None).cpu()
    ref = O.flowfield(zs, sd, "warp_generator_s2c.flowfield")
    assert got.shape == (3, 3, 16, 16, 16) and (got - ref).abs().max().item() <= TOL
    assert got.min().item() >= 0.0 and got.max().item() < 1.0


def test_compute_rotation_and_rt_warp():
    """compute_rotation_matrix (model.py:811-856) and compute_rt_warp (model.py:777-809), invert False / True."""
    import gbase_oracle as O
    from megaportrait_hack_b200 import model
    g = torch.Generator().manual_seed(12)
    R = (torch.rand(3, 3, generator=g) - 0.5) * 150.0
    t = (torch.rand(3, 3, generator=g) - 0.5) * 0.6
    got = model.compute_rotation_matrix(R.cuda()).cpu()
    assert (got - O.rotation_matrix(R)).abs().max().item() <= 2e-6
    for invert in (False, True):
        w = model.compute_rt_warp(R.cuda(), t.cuda(), invert=invert, grid_size=64).cpu()
        ref = O.rt_warp(R, t, invert, 64)
        assert w.shape == (3, 3, 64, 64, 64) and (w - ref).abs().max().item() <= 5e-6
    w16 = model.compute_rt_warp(R.cuda(), t.cuda(), invert=False, grid_size=16).cpu()
    assert (w16 - O.rt_warp(R, t, False, 16)).abs().max().item() <= 5e-6


def test_image_pyramide_and_eapp_forward(gb):
    """ImagePyramide.forward (model.py:1081-1085) and Eapp.forward -> (vs, es) (model.py:245-299) as modules."""
    import gbase_oracle as O
    G, sd = gb
    x = torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(13))
    with torch.no_grad():
        pyr = G.image_pyramid(x.cuda())
    ref = O.image_pyramid(x)
    assert set(pyr) == set(ref) == {"prediction_0.5", "prediction_0.25"}
    for k in ref:
        assert (pyr[k].cpu() - ref[k]).abs().max().item() <= 2e-6


def test_g3d_trains_through_libmpb200():
    """Row f-2: `G3d(x).backward()` with autograd recording on runs the residual blocks and the final convolution through
    ops.ConvFunction / ops.GroupNormFunction (CUDA forward and backward); gradients of all 70 G3d tensors and of the input against
    autograd of the CPU oracle in float64 (<= 1e-3 relative).

    ReLU masks: the three-pass forward differs from float64 by ~1e-5 of abs-max, so a few pre-activations per million change sign;
    each flipped element carries an O(1) gradient error (sparse outliers: max-norm 1e-2 on a weight gradient, 0.17 on one input
    element, mean error 1e-4 of the mean gradient) that says nothing about the backward kernels.  The oracle is therefore
    evaluated with THIS run's masks (relu(z) := z * mask): what is compared is the gradient of the function the GPU computed.
    The unmasked comparison and float32 ATen autograd (which flips 100x fewer masks: forward error 1e-7) are printed beside it."""
    import unittest.mock as mock
    import gbase_oracle as O
    from megaportrait_hack_b200 import lib, model, seeded
    lib.build()
    sd = {k: v for k, v in seeded.seeded_state_dict(seed=0).items() if k.startswith("G3d.")}
    G = model.G3d(96)
    G.load_state_dict({k[len("G3d."):]: v for k, v in sd.items()})
    G = G.cuda().train()
    x = torch.randn(1, 96, 8, 32, 32, generator=torch.Generator().manual_seed(7))
    go = torch.randn(1, 96, 8, 32, 32, generator=torch.Generator().manual_seed(8))
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- GPU run, recording the ReLU masks in call order
    masks = []
    real_relu = torch.relu

    def recording_relu(z):
        y = real_relu(z)
        masks.append((y > 0).cpu())
        return y

    xc = x.cuda().requires_grad_(True)
    with mock.patch.object(torch, "relu", recording_relu):
        out = G(xc)
    assert out.requires_grad and len(masks) == 14
    out.backward(go.cuda())

    def oracle(dtype, use_masks):
        sdd = {k: v.to(dtype).requires_grad_(True) for k, v in sd.items()}
        xd = x.to(dtype).clone().requires_grad_(True)
        it = iter(masks)
        flips = [0, 0]

        def masked_relu(z, inplace=False):
            m = next(it)
            own = z > 0
            flips[0] += int((own != m).sum())
            flips[1] += m.numel()
            return z * (m if use_masks else own).to(z.dtype)

        with mock.patch.object(O.F, "relu", masked_relu):
            ref = O.g3d(xd, sdd)
        ref.backward(go.to(dtype))
        return ref.detach(), xd.grad, {k: v.grad for k, v in sdd.items()}, flips

    def worst_of(grads, ref_grads):
        worst = ("", 0.0)
        for name, p in grads.items():
            want = ref_grads[name].float()
            e = ((p.float().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-20)).item()
            if e > worst[1]:
                worst = (name, e)
        return worst

    ours = {"G3d." + n: p.grad for n, p in G.named_parameters()}
    assert len(ours) == 70 and all(g is not None for g in ours.values())
    ref_m, gx_m, gp_m, flips = oracle(torch.float64, True)
    ref_u, gx_u, gp_u, _ = oracle(torch.float64, False)
    _, gx_32, gp_32, _ = oracle(torch.float32, False)
    fwd = ((out.detach().cpu() - ref_u.float()).abs().max() / ref_u.abs().max()).item()
    wm, wu, w32 = worst_of(ours, gp_m), worst_of(ours, gp_u), worst_of(gp_32, gp_u)
    exm = ((xc.grad.cpu() - gx_m.float()).abs().max() / gx_m.abs().max()).item()
    exu = ((xc.grad.cpu() - gx_u.float()).abs().max() / gx_u.abs().max()).item()
    ex32 = ((gx_32.double() - gx_u).abs().max() / gx_u.abs().max()).item()
    print(f"G3d forward vs float64 oracle: {fwd:.3e}; ReLU masks that differ from float64: {flips[0]} of {flips[1]}")
    print(f"gradients vs float64 oracle with this run's masks: worst of 70 tensors {wm[1]:.3e} ({wm[0]}), input {exm:.3e}")
    print(f"gradients vs float64 oracle with its own masks:    worst {wu[1]:.3e} ({wu[0]}), input {exu:.3e}")
    print(f"float32 ATen autograd vs float64 (own masks):      worst {w32[1]:.3e} ({w32[0]}), input {ex32:.3e}")
    assert fwd < 1e-4
    assert wm[1] < 1e-3 and exm < 1e-3                       # the bar: gradient of the computed function
    assert flips[0] <= 1e-4 * flips[1]                       # a handful of masks per million may flip
    assert wu[1] < 5e-2                                      # ... and that is all the unmasked comparison sees
    # and the inference path is untouched: no_grad takes the fused channels-last kernels
    with torch.no_grad():
        y = G.eval()(x.cuda())
    assert not y.requires_grad and ((y.cpu() - ref_u.float()).abs().max() / ref_u.abs().max()).item() < 1e-4


def _masked_grad_parity(mod, oracle_fn, sd, prefix, x, go):
    """Gradients of `mod(x)` (libmpb200 Functions) vs float64 autograd of `oracle_fn(x, sd)` evaluated with the GPU run's ReLU masks
    (see test_g3d_trains_through_libmpb200).  Returns (forward error, worst parameter-gradient error, input-gradient error)."""
    import unittest.mock as mock
    import gbase_oracle as O
    masks = []
    real_relu = torch.relu

    def recording_relu(z):
        y = real_relu(z)
        masks.append((y > 0).cpu())
        return y

    xc = x.cuda().requires_grad_(True)
    with mock.patch.object(torch, "relu", recording_relu):
        out = mod(xc)
    assert out.requires_grad
    out.backward(go.cuda())
    sdd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xd = x.double().clone().requires_grad_(True)
    it = iter(masks)
    with mock.patch.object(O.F, "relu", lambda z, inplace=False: z * next(it).to(z.dtype)):
        ref = oracle_fn(xd, sdd)
    ref.backward(go.double())
    fwd = ((out.detach().cpu() - ref.detach().float()).abs().max() / ref.abs().max()).item()
    worst = 0.0
    n = 0
    for name, p in mod.named_parameters():
        want = sdd[prefix + name].grad
        assert p.grad is not None and want is not None, name
        worst = max(worst, ((p.grad.cpu().double() - want).abs().max() / want.abs().max().clamp_min(1e-20)).item())
        n += 1
    ex = ((xc.grad.cpu().double() - xd.grad).abs().max() / xd.grad.abs().max()).item()
    return fwd, worst, ex, n


@pytest.mark.parametrize("which", ["conv2d_ws", "resblock_custom", "adaptive_group_norm", "resblock3d_adaptive", "resblock3d"])
def test_path_modules_are_differentiable_through_libmpb200(which):
    """Row f-2: the building blocks of Eapp / the warp generators / G3d (`Conv2d_WS`, `ResBlock_Custom`, `AdaptiveGroupNorm`,
    `ResBlock3D_Adaptive`, `ResBlock3D`: model.py:54-130, 304-408, 500-528) called with autograd recording on: forward AND backward
    on libmpb200, every parameter gradient and the input gradient against float64 autograd of the CPU oracle."""
    import gbase_oracle as O
    from megaportrait_hack_b200 import lib, model, seeded
    lib.build()
    g = torch.Generator().manual_seed(31)
    if which == "conv2d_ws":
        mod, x = model.Conv2d_WS(64, 128, 3, padding=1), torch.randn(2, 64, 24, 40, generator=g)
        fn = lambda xx, sdd: O.conv2d_ws(xx, sdd, "m")
    elif which == "resblock_custom":
        mod, x = model.ResBlock_Custom(2, 64, 128), torch.randn(1, 64, 32, 48, generator=g)
        fn = lambda xx, sdd: O.resblock_custom2d(xx, sdd, "m")
    elif which == "adaptive_group_norm":
        mod, x = model.AdaptiveGroupNorm(96), torch.randn(2, 96, 4, 8, 12, generator=g) * 1.4 + 0.2
        fn = lambda xx, sdd: O.adaptive_group_norm(xx, sdd, "m")
    elif which == "resblock3d_adaptive":
        mod, x = model.ResBlock3D_Adaptive(128, 64), torch.randn(1, 128, 8, 8, 8, generator=g)
        fn = lambda xx, sdd: O.resblock3d_adaptive(xx, sdd, "m")
    else:
        mod, x = model.ResBlock3D(96, 192), torch.randn(1, 96, 4, 16, 16, generator=g)
        fn = lambda xx, sdd: O.resblock3d(xx, sdd, "m")
    with torch.no_grad():
        for name, p in mod.named_parameters():       # away from the (1, 0) defaults of the affine parameters
            if p.dim() >= 4 and p.shape[0] != 1:
                p.copy_(seeded.seeded_tensor("t." + name, p.shape, seeded.CONV_W, 5))
            elif "bias" in name:
                p.copy_(seeded.seeded_tensor("t." + name, p.shape, seeded.NORM_B, 5))
            else:
                p.copy_(seeded.seeded_tensor("t." + name, p.shape, seeded.NORM_W, 5))
    sd = {"m." + k: v.detach().clone() for k, v in mod.state_dict().items()}
    mod = mod.cuda().train()
    with torch.no_grad():
        shape = mod(x.cuda()).shape
    go = torch.randn(shape, generator=g)
    fwd, worst, ex, n = _masked_grad_parity(mod, fn, sd, "m.", x, go)
    print(f"{which}: forward {fwd:.2e}; {n} parameter gradients, worst {worst:.2e}; input gradient {ex:.2e}")
    assert fwd < 5e-5 and worst < 1e-4 and ex < 1e-4
