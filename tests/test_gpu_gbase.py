"""Stage and end-to-end parity of the B200 Gbase path against the CPU oracle (same seeded weights, same inputs).
Runs on the B200 box: `pytest -m gpu`.  Tolerances: RGB max-abs <= 1e-3 (BASELINE.json north_star); every
intermediate stage <= 2e-4 relative to its abs-max (RGB is insensitive to the warp path)."""
import os

import pytest
import torch

from conftest import load_frames, synthetic_pair

pytestmark = pytest.mark.gpu

RGB_TOL = 1e-3
STAGE_TOL = 2e-4
MOTION_TOL = 2e-4     # Emtn / ResNet-50 descriptor also run on the split-bf16 tcgen05 kernels (fp32-grade)


def _ncdhw(a):
    v = a.f32 if a.f32 is not None else a.hi.float() + a.lo.float()
    return v.permute(0, 4, 1, 2, 3).contiguous().cpu()


def rel(got, ref):
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def gbase(seeded_sd):
    from megaportrait_hack_b200 import lib, model, seeded
    lib.build()
    G = model.Gbase().eval()
    main = {k: v for k, v in seeded_sd.items() if not k.startswith(seeded.ROTNET_PREFIX)}
    missing = G.load_state_dict(main, strict=False)
    assert not missing.unexpected_keys and all(k.startswith("image_pyramid.") for k in missing.missing_keys)
    G.motionEncoder.rotation_net.model.load_state_dict(
        {k[len(seeded.ROTNET_PREFIX):]: v for k, v in seeded_sd.items() if k.startswith(seeded.ROTNET_PREFIX)})
    return G.to("cuda")


@pytest.fixture(scope="module")
def oracle_synth(seeded_sd):
    import gbase_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    xs, xd = synthetic_pair(3)
    with torch.no_grad():
        rgb, pyr, st = O.gbase_forward_shared_source(xs, xd, seeded_sd, stages=True)
    return xs, xd, rgb, pyr, st


def test_stage_parity(gbase, oracle_synth):
    """Every hot-path stage against the oracle (default settings: split-bf16 generator, fp16x2 motion encoder, fp16 + e4m3
    cross-term G2d res-blocks)."""
    xs, xd, rgb_o, pyr_o, st = oracle_synth
    src = gbase.encode_source(xs.cuda(), keep_stages=True)
    rgb, pyr, drv = gbase.drive(src, xd.cuda(), keep_stages=True)
    checks = {
        "vs": (_ncdhw(src["vs"]), st["vs"], STAGE_TOL), "es": (src["es"].cpu(), st["es"], STAGE_TOL),
        "zs": (src["zs"].cpu(), st["zs"], STAGE_TOL), "ts": (src["ts"].cpu(), st["ts"], STAGE_TOL),
        "Rs": (src["Rs"].cpu(), st["Rs"], STAGE_TOL),
        "vc": (_ncdhw(src["vc"]), st["vc"], STAGE_TOL), "vc2d": (_ncdhw(src["vc2d"]), st["vc2d"], STAGE_TOL),
        "zd": (drv["zd"].cpu(), st["zd"], STAGE_TOL), "td": (drv["td"].cpu(), st["td"], STAGE_TOL),
        "projected": (_ncdhw(drv["projected"]).squeeze(2), st["projected"], STAGE_TOL),
    }
    errs = {k: rel(g, r) for k, (g, r, _) in checks.items()}
    print("stage rel errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, (g, r, tol) in checks.items():
        assert errs[k] <= tol, f"stage {k}: {errs[k]:.3e} > {tol}"
    assert (rgb.cpu() - rgb_o).abs().max().item() <= RGB_TOL
    for k in pyr_o:
        assert (pyr[k].cpu() - pyr_o[k]).abs().max().item() <= RGB_TOL


def test_rgb_parity_shared_source_split_bf16_g2d(gbase, oracle_synth):
    """BASELINE config 2 semantics (1 source x N drivers) with G2d's res-blocks forced onto the three-pass split-bf16
    plans (what the pipeline uses when the F16_Q8 range check rejects a tensor): same budget."""
    from megaportrait_hack_b200 import model
    xs, xd, rgb_o, pyr_o, st = oracle_synth
    saved = model._Q8_ENABLED
    model._Q8_ENABLED = False
    model.invalidate_plans(gbase.G2d)
    try:
        src = gbase.encode_source(xs.cuda(), keep_stages=True)
        rgb, pyr, drv = gbase.drive(src, xd.cuda(), keep_stages=True)
    finally:
        model._Q8_ENABLED = saved
        model.invalidate_plans(gbase.G2d)
    err = (rgb.cpu() - rgb_o).abs().max().item()
    print("rgb max-abs (split-bf16 G2d):", err, "zd rel:", rel(drv["zd"].cpu(), st["zd"]))
    assert err <= RGB_TOL
    assert rel(drv["zd"].cpu(), st["zd"]) <= MOTION_TOL
    assert rel(_ncdhw(src["vc2d"]), st["vc2d"]) <= STAGE_TOL


def test_forward_signature_and_real_frames(gbase, seeded_sd):
    """BASELINE config 1: the two real junk/ frames through `Gbase.forward(xs, xd) -> (img, pyramids)`."""
    import gbase_oracle as O
    xs, xd = load_frames()
    with torch.no_grad():
        rgb_o, pyr_o = O.gbase_forward(xs, xd, seeded_sd)
        out = gbase(xs.cuda(), xd.cuda())
    assert isinstance(out, tuple) and len(out) == 2 and set(out[1]) == {"prediction_0.5", "prediction_0.25"}
    assert out[0].shape == (1, 3, 512, 512) and out[1]["prediction_0.25"].shape == (1, 3, 128, 128)
    assert (out[0].cpu() - rgb_o).abs().max().item() <= RGB_TOL
    assert (out[1]["prediction_0.5"].cpu() - pyr_o["prediction_0.5"]).abs().max().item() <= RGB_TOL


def test_batch_invariance_and_determinism(gbase):
    xs, xd = synthetic_pair(4)
    src = gbase.encode_source(xs.cuda())
    a, _ = gbase.drive(src, xd.cuda())
    b, _ = gbase.drive(src, xd.cuda())
    assert torch.equal(a, b), "run-to-run determinism"
    c, _ = gbase.drive(src, xd[1:3].cuda())
    assert (a[1:3] - c).abs().max().item() <= 1e-4, "batch invariance (cuDNN picks per-batch algorithms in Emtn)"
    # reference semantics Bs == Bd (model.py:993): per-sample sources
    with torch.no_grad():
        full, _ = gbase(xs.expand(2, -1, -1, -1).contiguous().cuda(), xd[:2].cuda())
    assert (full - a[:2]).abs().max().item() <= 1e-4


def test_module_level_dropins(gbase, oracle_synth):
    """The PairwiseTransferLoss access pattern (model.py:2192-2214): sub-modules called one by one on NCDHW tensors."""
    from megaportrait_hack_b200 import model
    xs, xd, rgb_o, pyr_o, st = oracle_synth
    w = gbase.warp_generator_s2c(st["Rs"].cuda(), st["ts"].cuda(), st["zs"].cuda(), st["es"].cuda())
    assert w.shape == (1, 3, 64, 64, 64) and rel(w.cpu(), st["w_s2c"]) <= STAGE_TOL
    vc = model.apply_warping_field(st["vs"].cuda(), w)
    assert rel(vc.cpu(), st["vc"]) <= STAGE_TOL
    vc2d = gbase.G3d(st["vc"].cuda())
    assert rel(vc2d.cpu(), st["vc2d"]) <= STAGE_TOL
    img = gbase.G2d(st["projected"].cuda())
    assert (img.cpu() - st["rgb"]).abs().max().item() <= RGB_TOL
    with pytest.raises(AssertionError):
        gbase.warp_generator_c2d(st["Rs"].cuda(), st["ts"].cuda(), st["zs"].cuda(), st["es"].cuda().repeat(2, 1))


def test_motion_encoder_and_descriptor_on_libmpb200(gbase, oracle_synth, seeded_sd):
    """Row f-1 / a7: Emtn (2x ResNet-18 + RepVGG-B1g2) and CustomResNet50 on the tcgen05 kernels vs the oracle, and the
    stock cuDNN backend (fp32) as a second opinion."""
    import gbase_oracle as O
    from megaportrait_hack_b200 import ops
    xs, xd, rgb_o, pyr_o, st = oracle_synth
    l0 = ops.LAUNCHES
    with torch.no_grad():
        R, t, z = gbase.motionEncoder(xd.cuda())
    assert ops.LAUNCHES - l0 > 50, "Emtn must run on libmpb200 kernels"
    assert rel(z.cpu(), st["zd"]) <= STAGE_TOL and rel(t.cpu(), st["td"]) <= STAGE_TOL
    assert rel(R.cpu(), st["Rd"]) <= STAGE_TOL
    with torch.no_grad():
        vs, es = gbase.appearanceEncoder(xs.cuda())
    assert rel(es.cpu(), st["es"]) <= STAGE_TOL and rel(vs.cpu(), st["vs"]) <= STAGE_TOL
    gbase.motionEncoder.backend = "cudnn"
    try:
        with torch.no_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            R2, t2, z2 = gbase.motionEncoder(xd.cuda())
    finally:
        gbase.motionEncoder.backend = "mpb200"
    assert rel(z2.cpu(), st["zd"]) <= STAGE_TOL


def test_no_cpu_fallback(gbase):
    from megaportrait_hack_b200 import model
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.G2d(96).eval()(torch.zeros(1, 96, 64, 64))


def test_cuda_graph_engine_matches_eager(gbase):
    """engine.GraphedGbase replays encode_source / drive from CUDA graphs: same numbers as the eager calls."""
    from megaportrait_hack_b200.engine import GraphedGbase
    xs, xd = synthetic_pair(2)
    eager, pyr_e = gbase.drive(gbase.encode_source(xs.cuda()), xd.cuda())
    eng = GraphedGbase(gbase, 2, "cuda")
    for _ in range(2):      # second replay: static buffers are reused correctly
        out, pyr = eng.step(xs.cuda(), xd.cuda())
        torch.cuda.synchronize()
        assert (out - eager).abs().max().item() <= 1e-6
        assert (pyr["prediction_0.5"] - pyr_e["prediction_0.5"]).abs().max().item() <= 1e-6
    xd2 = torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(9))
    out2, _ = eng.step(xs.cuda(), xd2.cuda())
    ref2, _ = gbase.drive(gbase.encode_source(xs.cuda()), xd2.cuda())
    assert (out2 - ref2).abs().max().item() <= 1e-6


@pytest.mark.parametrize("variant", ["seed3_gain", "outlier_channels", "small_weights"])
def test_rgb_parity_other_weight_statistics(variant):
    """VERDICT round 1 ("all evidence is seeded U(-1/sqrt(fan_in)) weights"): end-to-end RGB parity against the CPU oracle under
    other weight statistics -- another seed with 1.12x larger weights everywhere (activations of the BatchNorm-eval decoder grow
    ~10x through its 22 convolutions), heavy-tailed G2d weights (4 output channels of every decoder convolution 6x larger: outlier
    activation channels meet the FP8 cross-term planes and their per-tensor scales), and a decoder with 2^-6 x smaller weights
    (tiny activations: e4m3 underflow without the byte-plane scales).  Budget: 1e-3 max-abs (north star)."""
    import gbase_oracle as O
    from megaportrait_hack_b200 import lib, model, seeded
    lib.build()
    old_gain = seeded.WEIGHT_GAIN
    try:
        if variant == "seed3_gain":
            seeded.WEIGHT_GAIN = 1.12
            sd = seeded.seeded_state_dict(seed=3)
        else:
            sd = seeded.seeded_state_dict(seed=0)
    finally:
        seeded.WEIGHT_GAIN = old_gain
    if variant == "outlier_channels":
        g = torch.Generator().manual_seed(5)
        for k, v in sd.items():
            if k.startswith("G2d.") and k.endswith(".weight") and v.dim() == 4 and v.shape[0] >= 64:
                idx = torch.randperm(v.shape[0], generator=g)[:4]
                v[idx] *= 6.0
    if variant == "small_weights":
        for k, v in sd.items():
            if k.startswith("G2d.res_blocks.") and k.endswith("conv1.weight"):
                v *= 2.0 ** -6          # (BatchNorm eval does not re-normalise: the inner activations shrink 64x)
    G = model.Gbase().eval()
    G.load_state_dict({k: v for k, v in sd.items() if not k.startswith(seeded.ROTNET_PREFIX)}, strict=False)
    G.motionEncoder.rotation_net.model.load_state_dict(
        {k[len(seeded.ROTNET_PREFIX):]: v for k, v in sd.items() if k.startswith(seeded.ROTNET_PREFIX)})
    G = G.to("cuda")
    xs, xd = synthetic_pair(1)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        rgb_o, _ = O.gbase_forward(xs, xd, sd)
        rgb, _ = G(xs.cuda(), xd.cuda())
    err = (rgb.cpu() - rgb_o).abs().max().item()
    spread = (rgb_o.max() - rgb_o.min()).item()
    P = G.G2d._plan()
    print(f"{variant}: RGB max-abs vs oracle {err:.3e}; oracle RGB range {spread:.3f}; q8_ok={P.get('q8_ok')} "
          f"amax max {max(P.get('q8_amax', [0])):.3g}")
    assert spread > 0.05            # the image is not saturated to a constant: the comparison means something
    assert err <= RGB_TOL
