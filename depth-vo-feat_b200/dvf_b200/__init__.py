"""dvf_b200 -- B200 (sm_100a) implementation of Depth-VO-Feat's inverse warp + reconstruction losses.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (include/dvf_b200.h)
  dvf_b200/        ctypes loader, torch.autograd bindings, synthetic KITTI-shaped inputs
  inverse_warp.py, loss_functions.py, loss_functions_sfm.py, loss_function_sfm_old.py
                   drop-in modules with the reference's names and signatures: put this directory
                   in front of pytorch_version/ on sys.path and the training scripts pick them up.
"""
from ._lib import DvfError, load  # noqa: F401

__version__ = "0.1.0"
