"""Drop-in `model` module: `import model; model.Gbase()` resolves to the B200-native implementation
(megaportrait-hack_b200/model.py) under the reference's own names (reference model.py:54-1180).

Names of the reference's `model.py` that are outside the hot-path scope (losses, discriminator, data utilities --
SURVEY.md section 2, rows 16-19) are not re-implemented; asking for one raises an ImportError that says so.
"""
from megaportrait_hack_b200.model import *  # noqa: F401,F403
from megaportrait_hack_b200.model import (  # noqa: F401
    COMPRESS_DIM, FEATURE_SIZE, FEATURE_SIZE_AVG_POOL, AdaptiveGroupNorm, AntiAliasInterpolation2d, Conv2d_WS,
    Conv3D_WS, CustomResNet50, Eapp, Emtn, FlowField, G2d, G3d, Gbase, ImagePyramide, ResBlock2D, ResBlock3D,
    ResBlock3D_Adaptive, ResBlock_Custom, SixDRepNet_Detector, WarpGeneratorC2D, WarpGeneratorS2C,
    apply_warping_field, compute_rotation_matrix, compute_rt_warp, device)

_OUT_OF_SCOPE = ("PerceptualLoss", "IdentitySimilarityLoss", "PairwiseTransferLoss", "Discriminator", "Genh", "GHR",
                 "Student", "crop_and_warp_face", "get_foreground_mask", "remove_background_and_convert_to_rgb",
                 "GazeBlinkLoss", "MPGazeLoss", "PatchGanEncoder")


def __getattr__(name):
    if name in _OUT_OF_SCOPE:
        raise ImportError(f"model.{name} is outside the B200 hot-path scope (SURVEY.md section 2/8f); "
                          "import it from the reference's own model.py")
    raise AttributeError(name)
