"""microimagelib_b200 -- B200-native (sm_100a) backend for microImageLib's volumetric hot path:
the Richardson-Lucy deconvolution loop and the affine-warp + ZNCC registration cost, behind the
reference's libapi.h C API.  See DESIGN.md.

    from microimagelib_b200 import libapi      # reference-shaped API on numpy arrays
    from microimagelib_b200 import device      # device-resident handles (numpy or torch CUDA tensors)
"""
__version__ = "0.1.0"
