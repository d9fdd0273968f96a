"""TEST INFRASTRUCTURE ONLY -- mint golden vectors from the REAL reference (`/root/reference/model.py`).

Run in the authoring container (the reference is not on the GPU box):

    python oracle/make_golden.py

It imports the reference through `oracle/ref_shim.py`, fills it with the per-key seeded weights
(`megaportrait_hack_b200/seeded.py`, seed 0), runs `Gbase(xs, xd)` in eval mode on CPU fp32 for
  * case "synthetic": torch.Generator().manual_seed(1); xs, xd = rand(1,3,512,512) each,
  * case "real": frame 0 of junk/-2KGPYEFnsU_8.mp4 (source) and of junk/-2KGPYEFnsU_11.mp4 (driver), RGB/255
    (BASELINE config 1); the two frames are also saved as PNG fixtures,
captures every stage tensor with forward hooks, and stores for each one a strided sample (<= 16384 values)
plus float64 moments in `tests/golden/gbase_<case>.npz`.  The reference ships no tests or golden vectors of its
own, so these files are what pins `oracle/gbase_oracle.py`.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from megaportrait_hack_b200 import seeded  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MAX_SAMPLES = 16384


def summarise(t: torch.Tensor):
    flat = t.detach().reshape(-1).to(torch.float32)
    stride = max(1, (flat.numel() + MAX_SAMPLES - 1) // MAX_SAMPLES)
    d = flat.double()
    moments = np.array([d.mean().item(), d.pow(2).mean().sqrt().item(), d.abs().max().item(), float(stride),
                        float(flat.numel())])
    return flat[::stride].numpy().copy(), moments


def real_frames():
    import cv2
    out = []
    for name, tag in (("-2KGPYEFnsU_8.mp4", "src"), ("-2KGPYEFnsU_11.mp4", "drv")):
        cap = cv2.VideoCapture(os.path.join(ref_shim.REFERENCE_DIR, "junk", name))
        ok, frame = cap.read()
        assert ok and frame.shape == (512, 512, 3), frame.shape
        cv2.imwrite(os.path.join(GOLD, f"frame_{tag}.png"), frame)  # BGR on disk, as cv2 expects
        rgb = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
        out.append(torch.from_numpy(rgb).permute(2, 0, 1)[None].float() / 255.0)
    return out


def run_case(ref, G, xs, xd):
    st = {}
    handles = []

    def hook(mod, name, fn):
        handles.append(mod.register_forward_hook(lambda m, i, o: fn(i, o)))

    emtn_calls = []
    hook(G.appearanceEncoder, "eapp", lambda i, o: st.update(vs=o[0], es=o[1]))
    hook(G.motionEncoder, "emtn", lambda i, o: emtn_calls.append(o))
    hook(G.warp_generator_s2c.flowfield, "ff_s2c", lambda i, o: st.update(em_s2c=o))
    hook(G.warp_generator_c2d.flowfield, "ff_c2d", lambda i, o: st.update(em_c2d=o))
    hook(G.warp_generator_s2c, "s2c", lambda i, o: st.update(w_s2c=o))
    hook(G.warp_generator_c2d, "c2d", lambda i, o: st.update(w_c2d=o))
    hook(G.G3d, "g3d", lambda i, o: st.update(vc=i[0], vc2d=o))
    hook(G.G2d, "g2d", lambda i, o: st.update(projected=i[0]))
    warps = []
    orig = ref.apply_warping_field

    def awf(v, w):
        out = orig(v, w)
        warps.append(out)
        return out

    ref.apply_warping_field = awf
    try:
        with torch.no_grad():
            rgb, pyr = G(xs, xd)
    finally:
        ref.apply_warping_field = orig
        for h in handles:
            h.remove()
    (st["Rs"], st["ts"], st["zs"]), (st["Rd"], st["td"], st["zd"]) = emtn_calls
    st["warped"] = warps[1]
    st["rgb"] = rgb
    st["pyr_0.5"] = pyr["prediction_0.5"]
    st["pyr_0.25"] = pyr["prediction_0.25"]
    return st


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load_reference_model()
    G = ref.Gbase().eval()
    seeded.apply_seeded(G, 0, rotnet=G.motionEncoder.rotation_net.model)
    g = torch.Generator().manual_seed(1)
    xs = torch.rand(1, 3, 512, 512, generator=g)
    xd = torch.rand(1, 3, 512, 512, generator=g)
    cases = {"synthetic": (xs, xd), "real": tuple(real_frames())}
    for name, (a, b) in cases.items():
        st = run_case(ref, G, a, b)
        blob = {}
        for k, v in st.items():
            blob[k + ".sample"], blob[k + ".moments"] = summarise(v)
        np.savez_compressed(os.path.join(GOLD, f"gbase_{name}.npz"), **blob)
        print(name, {k: tuple(v.shape) for k, v in st.items()})
    # key/shape inventory of the reference state_dict (boundary contract)
    import json
    inv = {k: list(v.shape) for k, v in G.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_keys.json"), "w") as f:
        json.dump(inv, f, indent=0)
    # signatures of the boundary classes / functions
    import inspect
    names = ["Gbase", "Eapp", "Emtn", "WarpGeneratorS2C", "WarpGeneratorC2D", "FlowField", "G3d", "G2d",
             "ResBlock3D", "ResBlock3D_Adaptive", "ResBlock2D", "ResBlock_Custom", "AdaptiveGroupNorm",
             "Conv2d_WS", "Conv3D_WS", "ImagePyramide", "AntiAliasInterpolation2d", "CustomResNet50"]
    sigs = {}
    for n in names:
        c = getattr(ref, n)
        sigs[n] = {"init": str(inspect.signature(c.__init__)), "forward": str(inspect.signature(c.forward))}
    for n in ("apply_warping_field", "compute_rt_warp", "compute_rotation_matrix"):
        sigs[n] = {"call": str(inspect.signature(getattr(ref, n)))}
    with open(os.path.join(GOLD, "signatures.json"), "w") as f:
        json.dump(sigs, f, indent=1)


if __name__ == "__main__":
    main()
