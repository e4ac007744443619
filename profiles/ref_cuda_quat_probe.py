import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "depth-vo-feat_b200"))
import numpy as np, torch
from dvf_b200 import ops, synthetic as syn
from oracle import torch_port as tp
torch.backends.cuda.matmul.allow_tf32 = False
ops.ARITHMETIC = "ref_cuda"
for kind in ("kitti", "tiny", "large"):
    pose = syn.pose(4096, kind, 5).cuda()
    K, Kinv = [x.cuda() for x in syn.intrinsics(4096, 128, 416)]
    P_ref = K @ tp.pose_matrix(pose, "quat")
    _, P, _ = ops.pose_proj_fwd(pose, K, None, 1, "quat", [1.0])
    same = float((P[0].view(torch.int32) == P_ref.view(torch.int32)).float().mean())
    R_ref = tp.pose_matrix(pose, "quat")
    pm, _, _ = ops.pose_proj_fwd(pose, None, None, 1, "quat", [], want_posemat=True)
    same_pm = float((pm.view(torch.int32) == R_ref.view(torch.int32)).float().mean())
    print(kind, "P identical %.4f %%" % (same * 100), "posemat identical %.4f %%" % (same_pm * 100))
