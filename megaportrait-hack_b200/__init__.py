"""megaportrait-hack_b200 -- B200-native Gbase volumetric forward path.

Layout: `csrc/` CUDA kernels + the C-ABI (`include/mpb200.h`), `lib.py` ctypes binding, `ops.py` tensor-level
wrappers, `model.py` the reference-named nn.Modules (drop-in for the reference's `model.py` hot path),
`engine.py` the batched / sharded driver-frame runner, `seeded.py` deterministic synthetic weights.
"""
__version__ = "0.1.0"
