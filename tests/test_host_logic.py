"""Host-side logic of the product modules (packing, folding, permutations, wiring, split-bf16 operand format)
checked on CPU against the oracle, with the kernels emulated by tests/fake_ops.py.  No GPU, no product fallback."""
import pytest
import torch

from conftest import synthetic_pair


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()


@pytest.fixture(scope="module")
def setup(seeded_sd):
    from megaportrait_hack_b200 import model, seeded
    G = model.Gbase().eval()
    G.load_state_dict({k: v for k, v in seeded_sd.items() if not k.startswith(seeded.ROTNET_PREFIX)}, strict=False)
    G.motionEncoder.rotation_net.model.load_state_dict(
        {k[len(seeded.ROTNET_PREFIX):]: v for k, v in seeded_sd.items() if k.startswith(seeded.ROTNET_PREFIX)})
    return G


def test_pipeline_wiring_against_oracle(setup, seeded_sd):
    import fake_ops
    import gbase_oracle as O
    G = setup
    xs, xd = synthetic_pair(2)
    with torch.no_grad():
        rgb_o, pyr_o, st = O.gbase_forward_shared_source(xs, xd, seeded_sd, stages=True)
        with fake_ops.installed():
            src = G.encode_source(xs, keep_stages=True)
            rgb, pyr, drv = G.drive(src, xd, keep_stages=True)
            w = G.warp_generator_c2d(st["Rd"], st["td"], st["zd"], st["es"].expand(2, -1))
    v = lambda a: fake_ops._to_ncdhw(fake_ops._val(a))
    assert rel(v(src["vs"]), st["vs"]) < 1e-4
    assert rel(v(src["vc"]), st["vc"]) < 1e-4
    assert rel(v(src["vc2d"]), st["vc2d"]) < 1e-4
    assert rel(v(drv["projected"]).squeeze(2), st["projected"]) < 1e-4
    assert rel(w, st["w_c2d"]) < 1e-4
    # motion encoder on the two-pass fp16 plans (im2col stems, fp16 activation planes): pooled outputs stay fp32-grade
    for k in ("Rs", "ts", "zs"):
        assert rel(src[k], st[k]) < 1e-4, k
    for k in ("Rd", "td", "zd"):
        assert rel(drv[k], st[k]) < 1e-4, k
    # split-bf16 (3-pass) operand format keeps RGB ~40x inside the 1e-3 budget
    assert (rgb - rgb_o).abs().max().item() < 1e-4
    assert (pyr["prediction_0.25"] - pyr_o["prediction_0.25"]).abs().max().item() < 1e-4


def test_folding_helpers():
    from megaportrait_hack_b200 import ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 4, 3, 3, generator=g)
    b = torch.randn(8, generator=g)
    bn = {"weight": torch.rand(8, generator=g) + 0.5, "bias": torch.randn(8, generator=g),
          "running_mean": torch.randn(8, generator=g), "running_var": torch.rand(8, generator=g) + 0.5}
    x = torch.randn(2, 4, 9, 9, generator=g)
    ref = F.batch_norm(F.conv2d(x, w, b, padding=1), bn["running_mean"], bn["running_var"], bn["weight"], bn["bias"],
                       False, 0.0, 1e-5)
    wf, bf = ops.fold_bn(w, b, bn)
    assert torch.allclose(F.conv2d(x, wf.float(), bf.float(), padding=1), ref, atol=1e-5)
    ws = ops.standardize_weight(w).float()
    m = w - w.mean(dim=(1, 2, 3), keepdim=True)
    assert torch.allclose(ws, m / (m.view(8, -1).std(dim=1).view(-1, 1, 1, 1) + 1e-5), atol=1e-6)
    pw = ops.pack_conv(w, b)
    assert pw.w_hi.shape == (16, 36) and pw.Cout == 8 and pw.Cout_pad == 16 and pw.k == (1, 3, 3)
    rec = (pw.w_hi.float() + pw.w_lo.float())[:8].view(8, 3, 3, 4).permute(0, 3, 1, 2)
    assert (rec - w).abs().max() <= w.abs().max() * 2.0 ** -16
    assert (pw.w_hi[8:] == 0).all()


def test_pipeline_wiring_with_split_bf16_switches():
    """The A/B switches (three-pass split-bf16 for the motion encoder and for G2d's res-blocks) are read at import time:
    re-run the wiring test in a fresh interpreter with both set, so the alternative plans stay healthy."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, MPB200_EMTN_PREC="split", MPB200_G2D_PREC="split")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_host_logic.py"), "-q", "-x",
                        "-k", "test_pipeline_wiring_against_oracle"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def _train_grad_check(G, seeded_sd, xs, xd, tag):
    """Run the differentiable train-mode forward (Gbase._forward_autograd) and the oracle with the SAME rotations, back-propagate
    the same scalar, and return {name: (relative L2 error, relative max error)} over every registered parameter."""
    import gbase_oracle as O
    from megaportrait_hack_b200 import seeded
    import unittest.mock as mock
    G.train()
    for p in G.parameters():
        p.grad = None
    # ReLU masks of THIS run, in call order (tests/test_gpu_modules.py::test_g3d_trains_through_libmpb200 explains why: a
    # pre-activation within rounding distance of zero flips its mask and carries an O(1) gradient error that says nothing about
    # the backward operators; the oracle is evaluated as the function this run computed)
    masks = []
    real_relu = torch.relu

    def recording_relu(z):
        y = real_relu(z)
        masks.append((y > 0).cpu())
        return y

    with mock.patch.object(torch, "relu", recording_relu):
        rgb, pyr = G(xs, xd) if xs.is_cuda else G._forward_autograd(xs, xd)     # (GPU: the public entry point)
    loss = rgb.mean() + pyr["prediction_0.5"].mean() + pyr["prediction_0.25"].mean()
    loss.backward()
    with torch.no_grad():
        from megaportrait_hack_b200 import emtn_cuda
        rots = (emtn_cuda.rotation_forward(G.motionEncoder, xs), emtn_cuda.rotation_forward(G.motionEncoder, xd))
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and not k.startswith(seeded.ROTNET_PREFIX) else v)
          for k, v in seeded_sd.items()}
    it = iter(masks)
    flips = [0, 0]

    def masked_relu(z, inplace=False):
        m = next(it)
        assert m.shape == z.shape, (m.shape, z.shape)
        flips[0] += int(((z > 0) != m).sum())
        flips[1] += m.numel()
        return z * m.to(z.dtype)

    O.BN_TRAINING = True
    try:
        with mock.patch.object(O.F, "relu", masked_relu):
            rgb_o, pyr_o, _ = O.gbase_forward_train(xs.cpu(), xd.cpu(), sd, rotations=tuple(r.cpu() for r in rots))
    finally:
        O.BN_TRAINING = False
    assert next(it, None) is None, "the oracle used fewer ReLUs than the product path"
    print(f"[{tag}] ReLU masks that differ from the oracle's own: {flips[0]} of {flips[1]}")
    (rgb_o.mean() + pyr_o["prediction_0.5"].mean() + pyr_o["prediction_0.25"].mean()).backward()
    assert (rgb.detach().cpu() - rgb_o.detach()).abs().max().item() < 1e-3, tag
    errs = {}
    for name, p in G.named_parameters():
        want = sd[name].grad
        if name.endswith("adaptive_matrix_beta"):
            assert p.grad is None or float(p.grad.abs().max()) == 0.0       # unused by the reference's forward (model.py:934-935)
            continue
        assert p.grad is not None and want is not None, name
        d = (p.grad.cpu() - want).double()
        errs[name] = (float(d.norm() / want.double().norm().clamp_min(1e-30)),
                      float(d.abs().max() / want.abs().max().clamp_min(1e-30)), float(want.double().norm()))
    return errs


def significant(errs, floor=1e-6):
    """Biases in front of a normalisation layer have a mathematically zero gradient: both sides hold rounding noise there
    (norm ~1e-10 against ~1e-3 for the weights).  Parity is asserted on the tensors whose oracle gradient is not noise."""
    return {k: v for k, v in errs.items() if v[2] > floor}


@pytest.mark.skipif(not __import__("os").environ.get("MPB200_SLOW_CPU"), reason="several minutes of CPU convolutions; "
                    "set MPB200_SLOW_CPU=1 (the same check runs on the GPU in tests/test_gpu_train.py)")
def test_train_mode_wiring_against_oracle(setup, seeded_sd):
    """Row f-2 host logic: `Gbase.train()` forward + backward on the emulated kernels vs autograd of the train-mode oracle:
    every registered parameter receives the oracle's gradient (wiring, channel padding, folded 1x1 pair, BatchNorm-as-GroupNorm)."""
    import fake_ops
    import copy
    G = copy.deepcopy(setup)
    xs, xd = synthetic_pair(1)
    with fake_ops.installed():
        errs = _train_grad_check(G, seeded_sd, xs, xd, "cpu")
    sig = significant(errs)
    worst = sorted(sig.items(), key=lambda kv: -kv[1][0])[:8]
    print(f"{len(errs)} parameters, {len(sig)} with a non-zero gradient; worst relative-L2 gradient errors:", worst)
    print("smallest significant norms:", sorted(sig.items(), key=lambda kv: kv[1][2])[:5])
    print("noise tensors:", sorted(((k, v[2]) for k, v in errs.items() if k not in sig), key=lambda kv: -kv[1])[:5])
    groups = {}
    for k, v in sig.items():
        g = ".".join(k.split(".")[:3])
        groups[g] = max(groups.get(g, 0.0), v[0])
    print("worst relative-L2 error per module:", sorted(groups.items(), key=lambda kv: -kv[1]))
    assert len(errs) >= 640 and len(sig) >= 400
    assert worst[0][1][0] < 1e-2, worst


def test_train_operators_on_emulated_kernels():
    """Row f-2 host logic of the differentiable operators (ops.conv_train: channel padding, stride 2 through the zero-spread
    gradient, sub-sampled 1x1 shortcuts; ops.batch_norm_train: BatchNorm as a one-sample GroupNorm, running statistics) against
    ATen autograd, with the kernels emulated on CPU."""
    import fake_ops
    import torch.nn as nn
    import torch.nn.functional as F
    from megaportrait_hack_b200 import ops
    g = torch.Generator().manual_seed(3)
    cases = [((2, 3, 16, 16), (8, 3, 7, 7), 1), ((2, 3, 16, 16), (8, 3, 7, 7), 2), ((1, 32, 12, 12), (3, 32, 3, 3), 1),
             ((2, 16, 8, 8), (32, 16, 1, 1), 2), ((2, 16, 8, 8), (32, 16, 3, 3), 2), ((1, 16, 4, 6, 6), (16, 16, 3, 3, 3), 1),
             # 64 output channels on an RGB frame: the im2col route of the stem weight gradient (ops.RgbStemConvFunction)
             ((2, 3, 16, 16), (64, 3, 3, 3), 1), ((2, 3, 16, 16), (64, 3, 7, 7), 2)]
    with fake_ops.installed():
        for xs, ws, stride in cases:
            x = torch.randn(xs, generator=g, requires_grad=True)
            w = (torch.randn(ws, generator=g) * 0.1).requires_grad_(True)
            b = torch.randn(ws[0], generator=g, requires_grad=True)
            y = ops.conv_train(x, w, b, stride=stride)
            fn = F.conv3d if len(xs) == 5 else F.conv2d
            ref = fn(x, w, b, stride=stride, padding=tuple(k // 2 for k in ws[2:]))
            assert y.shape == ref.shape, (xs, ws, stride)
            go = torch.randn(ref.shape, generator=g)
            got = torch.autograd.grad(y, (x, w, b), go)
            want = torch.autograd.grad(ref, (x, w, b), go)
            assert rel(y, ref) < 1e-4, (xs, ws, stride)
            for a, r in zip(got, want):
                assert rel(a, r) < 1e-4, (xs, ws, stride)
        bn, bn_ref = nn.BatchNorm2d(16), nn.BatchNorm2d(16)
        with torch.no_grad():
            bn.weight.uniform_(0.5, 1.5, generator=g)
            bn.bias.normal_(generator=g)
            bn_ref.load_state_dict(bn.state_dict())
        for mode in (True, True, False):
            bn.train(mode), bn_ref.train(mode)
            x = (torch.randn(3, 16, 5, 7, generator=g) * 2 + 1).requires_grad_(True)
            y, ref = ops.batch_norm_train(x, bn), bn_ref(x)
            go = torch.randn(ref.shape, generator=g)
            got = torch.autograd.grad(y, (x, bn.weight, bn.bias), go)
            want = torch.autograd.grad(ref, (x, bn_ref.weight, bn_ref.bias), go)
            assert rel(y, ref) < 1e-5
            for a, r in zip(got, want):
                assert rel(a, r) < 1e-4
            assert rel(bn.running_mean, bn_ref.running_mean) < 1e-5 and rel(bn.running_var, bn_ref.running_var) < 1e-5
            assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)


def test_train_mode_submodules_against_oracle(setup, seeded_sd):
    """Row f-2 host logic in the default CPU suite (the whole-Gbase version above takes minutes): the train-mode / differentiable
    forward of the sub-modules at small sizes on the emulated kernels -- G2d (folded 1x1 pair, BatchNorm batch statistics, linear
    upsample Function, 3-channel head), G3d (pool / trilinear Functions), a warp generator (flow-field tower, nearest-upsample
    Function, rigid grid), the ResNet-50 descriptor trunk (7x7 stride-2 stem, Bottleneck blocks, sub-sampled 1x1 shortcuts) --
    outputs and parameter gradients against autograd of the train-mode oracle."""
    import copy
    import fake_ops
    import gbase_oracle as O
    G = copy.deepcopy(setup).train()
    g = torch.Generator().manual_seed(11)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in seeded_sd.items()}
    z, e = torch.randn(2, 512, generator=g) * 0.3, torch.randn(2, 512, generator=g) * 0.3
    R, t = torch.randn(2, 3, generator=g) * 10, torch.randn(2, 3, generator=g) * 0.1
    cases = [
        ("G2d", lambda: G.G2d._forward_autograd(torch.randn(2, 96, 8, 8, generator=torch.Generator().manual_seed(1))),
         lambda: O.g2d(torch.randn(2, 96, 8, 8, generator=torch.Generator().manual_seed(1)), sd)),
        ("G3d", lambda: G.G3d._forward_autograd(torch.randn(1, 96, 8, 16, 16, generator=torch.Generator().manual_seed(2))),
         lambda: O.g3d(torch.randn(1, 96, 8, 16, 16, generator=torch.Generator().manual_seed(2)), sd)),
        ("warp_generator_c2d", lambda: G.warp_generator_c2d._forward_autograd(R, t, z, e),
         lambda: O.warp_generator(R, t, z, e, sd, "warp_generator_c2d", invert=False)[0]),
        ("appearanceEncoder.custom_resnet50",
         lambda: G.appearanceEncoder.custom_resnet50._forward_autograd(torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3))),
         lambda: O.custom_resnet50(torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(3)), sd,
                                   "appearanceEncoder.custom_resnet50")),
    ]
    O.BN_TRAINING = True
    try:
        with fake_ops.installed():
            for prefix, ours, ref in cases:
                for p in G.parameters():
                    p.grad = None
                y, want = ours(), ref()
                # (the 43 train-mode BatchNorms of the ResNet-50 trunk see only 32 .. 2 048 values per channel at this input size and
                #  amplify the emulated operand rounding: 3e-4 here, 4e-6 at the real 512 x 512 input)
                tol = 1e-3 if "resnet50" in prefix else 1e-4
                assert y.shape == want.shape and rel(y.detach(), want.detach()) < tol, (prefix, rel(y.detach(), want.detach()))
                go = torch.randn(want.shape, generator=g)
                y.backward(go)
                want.backward(go)
                n = 0
                mine = [(name, p) for name, p in G.named_parameters()
                        if name.startswith(prefix + ".") and not name.endswith("adaptive_matrix_beta")]
                floor = 1e-5 * max(float(sd[name].grad.norm()) for name, _ in mine if sd[name].grad is not None)
                for name, p in mine:
                    wg = sd[name].grad
                    assert p.grad is not None and wg is not None, name
                    if float(wg.norm()) > floor:           # (biases in front of a normalisation: zero gradient, rounding noise)
                        err = float((p.grad - wg).norm() / wg.norm())
                        assert err < (1e-1 if "resnet50" in prefix else 2e-2), (name, err)     # (ReLU / max-pool ties flip under the split-bf16 emulation)
                        n += 1
                    sd[name].grad = None
                assert n >= 8, (prefix, n)
    finally:
        O.BN_TRAINING = False
