"""Pins the CPU oracle (oracle/gbase_oracle.py) to golden vectors minted from the REAL reference
(oracle/make_golden.py -> tests/golden/gbase_*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_frames, synthetic_pair

import gbase_oracle as O

# relative-to-absmax tolerance per stage: the reference and the restatement call the same ATen kernels, so
# only summation-order noise (<= 3e-6 relative observed) is expected.
TOL = 2e-5


def _check(st, gold):
    keys = sorted({k.split(".")[0] if not k.startswith("pyr") else k.rsplit(".", 1)[0] for k in gold.files})
    assert len(keys) >= 19
    for k in keys:
        sample, mom = gold[k + ".sample"], gold[k + ".moments"]
        name = {"em_s2c": "w_em16_s2c", "em_c2d": "w_em16_c2d"}.get(k, k)
        if k.startswith("pyr_"):
            t = st["pyramids"]["prediction_" + k[4:]]
        else:
            t = st[name]
        flat = t.reshape(-1)
        assert flat.numel() == int(mom[4]), k
        got = flat[:: int(mom[3])].numpy()
        scale = max(mom[2], 1e-6)
        err = np.abs(got - sample).max() / scale
        assert err <= TOL, f"{k}: rel-to-absmax err {err:.3e}"
        assert abs(flat.double().mean().item() - mom[0]) <= TOL * scale, k
        assert abs(flat.double().pow(2).mean().sqrt().item() - mom[1]) <= TOL * scale, k


def _stages(xs, xd, sd):
    rgb, pyr, st = O.gbase_forward(xs, xd, sd, stages=True)
    # the golden file stores the 16^3 FlowField outputs; recompute them from the oracle's own pieces
    for side, z, e, p in (("s2c", st["zs"], st["es"], "warp_generator_s2c"), ("c2d", st["zd"], st["es"], "warp_generator_c2d")):
        s = torch.matmul(z + e, sd[p + ".adaptive_matrix_gamma"])
        st["w_em16_" + side] = O.flowfield(s[:, :, None, None], sd, p + ".flowfield")
    return st


@pytest.mark.parametrize("case", ["synthetic", "real"])
def test_oracle_matches_reference_golden(case, seeded_sd):
    gold = np.load(os.path.join(GOLDEN, f"gbase_{case}.npz"))
    xs, xd = synthetic_pair(1) if case == "synthetic" else load_frames()
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        st = _stages(xs, xd, seeded_sd)
    _check(st, gold)


def test_oracle_live_against_reference_when_present(seeded_sd):
    """Same comparison on full tensors when /root/reference is importable (authoring container only)."""
    import ref_shim
    if not ref_shim.reference_available():
        pytest.skip("reference tree not present on this box")
    from megaportrait_hack_b200 import seeded
    ref = ref_shim.load_reference_model()
    eapp = ref.Eapp().eval()
    sub = {k[len("appearanceEncoder."):]: v for k, v in seeded_sd.items() if k.startswith("appearanceEncoder.")}
    eapp.load_state_dict(sub, strict=True)
    g = torch.Generator().manual_seed(7)
    x = torch.rand(1, 3, 512, 512, generator=g)
    with torch.no_grad():
        vs_ref, es_ref = eapp(x)
        vs, es = O.eapp(x, seeded_sd)
    assert (vs - vs_ref).abs().max() <= 2e-5 * vs_ref.abs().max()
    assert (es - es_ref).abs().max() <= 2e-5 * es_ref.abs().max()
    # op-level: apply_warping_field on a non-degenerate field
    v = torch.randn(1, 8, 16, 64, 64, generator=g)
    w = (torch.rand(1, 3, 64, 64, 64, generator=g) - 0.5) * 60
    assert torch.equal(ref.apply_warping_field(v, w), O.apply_warping_field(v, w))


def test_oracle_train_mode_matches_reference_golden(seeded_sd):
    """Row f-2: the checker of the CUDA training path -- `O.gbase_forward_train` under `O.BN_TRAINING` (batch-statistics BatchNorm)
    -- against gradients of the REAL reference in `.train()` mode (`oracle/make_golden_train.py`, `tests/golden/gbase_train.npz`):
    loss, a sample of the train-mode image and, for every registered parameter, the gradient's L2 norm and a strided sample.
    Loss and image: fp32 rounding.  Gradients: relative L2 of the samples and relative norm deviation <= 3e-2 (measured 3e-3 .. 1e-2,
    the noise floor of a function with ReLU kinks and grid_sample cell boundaries -- see the comment at the assert); biases in front
    of a normalisation hold pure rounding noise and are compared against the model's scale only."""
    gold = np.load(os.path.join(GOLDEN, "gbase_train.npz"))
    xs, xd = synthetic_pair(1)
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in seeded_sd.items()}
    O.BN_TRAINING = True
    try:
        rgb, pyr, _ = O.gbase_forward_train(xs, xd, sd)
    finally:
        O.BN_TRAINING = False
    loss = rgb.mean() + pyr["prediction_0.5"].mean() + pyr["prediction_0.25"].mean()
    loss.backward()
    assert abs(float(loss) - float(gold["loss"][0])) <= 1e-5 * abs(float(gold["loss"][0]))

    def sample(t, n):
        flat = t.detach().reshape(-1).float()
        stride = max(1, (flat.numel() + 255) // 256)
        s = flat[::stride].numpy()
        assert s.shape == n.shape
        return s

    want = gold["rgb.sample"]
    assert np.abs(sample(rgb, want) - want).max() <= 2e-5
    names = [str(n) for n in gold["names"]]
    assert len(names) >= 640
    top = max(float(gold["grad." + n + ".norm"][0]) for n in names)
    worst, per_module = ("", 0.0), {}
    for n in names:
        g = sd[n].grad
        w_s, w_n = gold["grad." + n + ".sample"], float(gold["grad." + n + ".norm"][0])
        if g is None:
            assert w_n == 0.0, n
            continue
        if w_n < 1e-7 * top:                         # zero-gradient tensors (rounding noise on both sides)
            assert float(g.double().norm()) < 1e-5 * top, n
            continue
        e_norm = abs(float(g.double().norm()) - w_n) / w_n
        e_smp = float(np.linalg.norm(sample(g, w_s).astype(np.float64) - w_s) / max(np.linalg.norm(w_s.astype(np.float64)), 1e-30))
        per_module[n.split(".")[0]] = max(per_module.get(n.split(".")[0], 0.0), e_norm, e_smp)
        if max(e_norm, e_smp) > worst[1]:
            worst = (n, max(e_norm, e_smp))
    print("gradient deviation oracle vs reference (train mode), worst per sub-module:",
          {k: f"{v:.1e}" for k, v in per_module.items()}, "worst tensor:", worst)
    # Loss and image agree to fp32 rounding (asserted above).  The gradients agree only to the noise floor of the function
    # itself: the reference and this restatement differ by ~1e-6 in the forward pass (operation order), which flips the ReLU mask
    # of pre-activations that close to zero and moves sampling coordinates across cell boundaries of grid_sample (whose
    # coordinate gradient is piecewise constant and jumps at the border clamp); each such element changes a gradient by O(1).
    # Measured: 3e-3 (G2d, G3d) .. 1e-2 (flow-field tower of the C2D warp) in relative L2 -- the yardstick for the GPU
    # gradient test (tests/test_gpu_train.py), which removes the ReLU part by evaluating this oracle with the run's masks.
    assert all(v <= 3e-2 for v in per_module.values()), per_module
