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
