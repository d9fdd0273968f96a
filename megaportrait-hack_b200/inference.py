"""Drop-in for the reference's `inference.py` plumbing: same function names and argument order,
pre- / post-processing on the device (`mp_frames_u8_to_f32`, `mp_frames_f32_to_u8`).

Two deliberate differences from `inference.py:10-44`, both documented in SURVEY.md 3.2: (1) `Gbase.forward` returns
`(image, pyramids)` and the reference calls `.squeeze` on that tuple (a TypeError as written) -- here the image is taken;
(2) only the uint8 frame -- for JPEG sources only the compressed bitstream, decoded by nvJPEG on the device -- crosses
PCIe.  PNG and video decoding stay with PIL / OpenCV on the host."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _is_jpeg(path) -> bool:
    try:
        with open(path, "rb") as f:
            return f.read(3) == b"\xff\xd8\xff"
    except OSError:
        return False


def load_image_device(image_path, device="cuda") -> torch.Tensor:
    """inference.py:10-13 with the decode on the GPU: JPEG files go through nvJPEG (`mp_decode_jpeg_frames`) straight into
    a device uint8 frame [1, H, W, 3]; other formats (the reference's own PNG frames) are decoded by PIL on the host."""
    if _is_jpeg(image_path):
        with open(image_path, "rb") as f:
            return ops.decode_jpeg_frames([f.read()], device)
    return load_image(image_path).to(device)


def load_image(image_path, transform=None) -> torch.Tensor:
    """inference.py:10-13 -- RGB uint8 frame [1, H, W, 3] (the normalisation `transform` of the reference runs on the
    device inside `inference_base`; a callable passed here is applied to the PIL image first, as in the reference)."""
    from PIL import Image
    image = Image.open(image_path).convert("RGB")
    if transform is not None:
        image = transform(image)
        if torch.is_tensor(image):       # the reference's own ToTensor/Normalize pipeline: already fp32 CHW
            return image.unsqueeze(0)
    return torch.from_numpy(np.asarray(image).copy()).unsqueeze(0)


@torch.no_grad()
def inference_frames(source_u8: torch.Tensor, driving_u8: torch.Tensor, Gbase, device="cuda") -> torch.Tensor:
    """uint8 HWC frames [1 or N, 512, 512, 3] / [N, 512, 512, 3] -> uint8 HWC output frames [N, 512, 512, 3] on the host."""
    Gbase.eval()
    src = ops.frames_u8_to_f32(source_u8.to(device, non_blocking=True).contiguous())
    drv = ops.frames_u8_to_f32(driving_u8.to(device, non_blocking=True).contiguous())
    state = Gbase.encode_source(src[:1]) if src.shape[0] == 1 else None
    out = Gbase.drive(state, drv)[0] if state is not None else Gbase(src, drv)[0]
    return ops.frames_f32_to_u8(out, shift=1.0, scale=0.5, reverse_channels=True).cpu()


def inference_base(source_image_path, driving_image_path, Gbase, device="cuda"):
    """inference.py:15-45 -- returns the uint8 HWC frame that the reference hands to `cv2.imwrite`."""
    src, drv = load_image_device(source_image_path, device), load_image_device(driving_image_path, device)
    return inference_frames(src, drv, Gbase, device)[0].numpy()
