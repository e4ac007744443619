"""Alias: unsupervise_sfm.py:29 imports `loss_function_sfm`, a module the reference never shipped;
its call site (:98-100) uses the signature of loss_function_sfm_old.py."""
from loss_function_sfm_old import *  # noqa: F401,F403
from loss_function_sfm_old import photometric_reconstruction_loss, explainability_loss, smooth_loss  # noqa: F401
