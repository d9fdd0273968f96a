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
