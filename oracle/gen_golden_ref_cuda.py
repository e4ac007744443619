"""Generate tests/golden/ref_cuda_warp.npz: the reference's torch operator sequence (oracle/torch_port.py, verified bit for
bit against the unmodified reference on torch-CPU by tests/test_torch_port.py) executed by torch-CUDA EAGER -- the
secondary oracle of SURVEY 8(c) / App. B.3.  Needs a GPU, so it runs on the GPU box:

    gpurun -- 'python oracle/gen_golden_ref_cuda.py gpurun_out/ref_cuda_warp.npz'      # then copy into tests/golden/

The fixture holds the inputs (regenerable from the seed; their SHA-256 guards against generator drift), torch-CUDA's
projection matrices P = K @ pose_vec2mat(pose) and torch-CUDA's warped images for both padding modes.  The CPU oracle's
DVFO_REF_CUDA restatement and the kernels' DVF_FLAG_REF_CUDA path are pinned to it.
"""
import hashlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "depth-vo-feat_b200"))
from dvf_b200 import synthetic as syn  # noqa: E402
from oracle import torch_port as tp     # noqa: E402

out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(REPO, "tests", "golden", "ref_cuda_warp.npz")
torch.backends.cuda.matmul.allow_tf32 = False
B, H, W, SEED = 2, 32, 104, 17
d = syn.stereo_temporal_batch(B, H, W, seed=SEED)
t = {k: v.cuda() for k, v in d.items()}
out = {"B": B, "H": H, "W": W, "seed": SEED,
       "inputs_sha": np.array([hashlib.sha256(np.ascontiguousarray(d[k].numpy()).tobytes()).hexdigest() for k in sorted(d)]),
       "torch": torch.__version__, "device": torch.cuda.get_device_name(0)}
for tag, pose in (("temporal", "T_2to1"), ("stereo", "T_R2L")):
    out["P_" + tag] = (t["intrinsics"] @ tp.pose_matrix(t[pose])).cpu().numpy()
    for pad in ("zeros", "border"):
        w = tp.warp(t["img_R1"], t["depth"], t[pose], t["intrinsics"], t["intrinsics_inv"], padding_mode=pad)
        out[f"warped_{tag}_{pad}"] = w.cpu().numpy()
np.savez_compressed(out_path, **out)
print("wrote", out_path, {k: getattr(v, "shape", v) for k, v in out.items()})
