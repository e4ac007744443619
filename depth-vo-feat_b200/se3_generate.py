"""Drop-in for pytorch_version/se3_generate.py of Depth-VO-Feat: the se(3) -> SE(3) exponential map as a
device kernel (csrc/dvf_se3.cu) instead of a numpy round trip through the host.  Same names and shapes:
generate_se3(input[B,6,1,1]) -> [B,1,4,4] float64, input = (rotation vector w, translation generator u),
output [[R, R u],[0,1]] (se3_generate.py:13-46).  CUDA tensors only."""
from dvf_b200 import ops as _ops

SE3_Generator_KITTI = _ops.SE3Exp
generate_se3 = _ops.SE3Exp.apply
