"""CPU: host-side logic of the drop-in modules -- reference-compatible names, signatures, assertion
behaviour, loud failure without CUDA tensors, and product/oracle separation."""
import inspect
import os
import re

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "depth-vo-feat_b200")


def test_dropin_signatures_match_reference():
    import inverse_warp as iw
    import loss_function_sfm_old as old
    import loss_functions as lf
    import loss_functions_sfm as sfm

    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values()]

    E = inspect.Parameter.empty
    # inverse_warp.py:160
    assert params(iw.inverse_warp)[:7] == [("img", E), ("depth", E), ("pose", E), ("intrinsics", E), ("intrinsics_inv", E),
                                           ("rotation_mode", "euler"), ("padding_mode", "zeros")]
    assert params(iw.pose_vec2mat) == [("vec", E), ("rotation_mode", "euler")]                  # :141
    assert [p[0] for p in params(iw.pixel2cam)] == ["depth", "intrinsics_inv"]                   # :26
    assert [p[0] for p in params(iw.cam2pixel)] == ["cam_coords", "proj_c2p_rot", "proj_c2p_tr", "padding_mode"]  # :43
    # loss_functions.py:7
    assert params(lf.photometric_reconstruction_loss) == [
        ("img_R2", E), ("img_R1", E), ("img_L2", E), ("depth", E), ("T_2to1", E), ("T_R2L", E), ("intrinsics", E),
        ("intrinsics_inv", E), ("rotation_mode", "euler"), ("padding_mode", "zeros")]
    assert params(lf.smooth_loss) == [("pred_map", E), ("scale_factor", 1)]
    # loss_functions_sfm.py:9
    # (the reference's nine parameters, then keyword-only-in-practice extensions with defaults that keep its behaviour)
    assert params(sfm.photometric_reconstruction_loss)[:9] == [
        ("tgt_img", E), ("ref_imgs", E), ("intrinsics", E), ("intrinsics_inv", E), ("depth", E),
        ("explainability_mask", E), ("pose", E), ("rotation_mode", "euler"), ("padding_mode", "zeros")]
    assert params(sfm.photometric_reconstruction_loss)[9:] == [("disparity_eps", None)]
    for name in ("explainability_loss", "smooth_loss", "compute_errors", "inverse_warp"):
        assert hasattr(sfm, name)
    # loss_function_sfm_old.py:7
    assert params(old.photometric_reconstruction_loss) == [
        ("img_R2", E), ("img_R1", E), ("img_L2", E), ("depth", E), ("T_2to1", E), ("T_R2L", E), ("mask", E),
        ("intrinsics", E), ("intrinsics_inv", E), ("rotation_mode", "euler"), ("padding_mode", "zeros")]
    import loss_function_sfm as alias   # the module unsupervise_sfm.py:29 imports
    assert alias.photometric_reconstruction_loss is old.photometric_reconstruction_loss


def test_check_sizes_assertion_text():
    import inverse_warp as iw
    with pytest.raises(AssertionError, match="wrong size for img, expected Bx3xHxW"):
        iw.check_sizes(torch.zeros(2, 4, 5, 6), "img", "B3HW")
    iw.check_sizes(torch.zeros(2, 3, 5, 6), "img", "B3HW")
    with pytest.raises(AssertionError):
        iw.check_sizes(torch.zeros(2, 5), "pose", "B6")


def test_cpu_tensors_fail_loudly():
    from dvf_b200 import DvfError
    import inverse_warp as iw
    import loss_functions as lf
    B, H, W = 1, 8, 12
    img, depth, pose = torch.rand(B, 3, H, W), torch.rand(B, H, W) + 1, torch.zeros(B, 6)
    K = torch.eye(3).unsqueeze(0)
    with pytest.raises(DvfError, match="no CPU fallback"):
        iw.inverse_warp(img, depth, pose, K, K)
    with pytest.raises(DvfError, match="no CPU fallback"):
        lf.photometric_reconstruction_loss(img, img, img, depth, pose, pose, K, K)
    with pytest.raises(DvfError):
        iw.pose_vec2mat(pose)


def test_regularisers_refuse_cpu_tensors():
    from dvf_b200 import DvfError
    import loss_functions as lf
    import loss_functions_sfm as sfm
    m = torch.rand(2, 1, 8, 8) + 0.5
    with pytest.raises(DvfError, match="no CPU fallback"):
        lf.smooth_loss([m], 2.0)
    with pytest.raises(DvfError, match="no CPU fallback"):
        sfm.explainability_loss([m * 0.5])


def test_product_never_imports_the_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle|cpu_oracle|libdvf_oracle|dvfo_", re.M)
    for root, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not pat.search(src), f"{f} references the oracle"
    for f in os.listdir(os.path.join(REPO, "include")):
        assert not pat.search(open(os.path.join(REPO, "include", f)).read())


def test_synthetic_inputs_are_deterministic():
    from dvf_b200 import synthetic as syn
    a = syn.stereo_temporal_batch(2, 16, 52, seed=3)
    b = syn.stereo_temporal_batch(2, 16, 52, seed=3)
    for k in a:
        assert torch.equal(a[k], b[k])
    assert a["depth"].min() > 3.0 and a["depth"].max() <= 50.0 + 1e-3
    assert float(a["T_R2L"][0, 0]) == pytest.approx(0.53233)


def test_bench_reference_arm_line_and_config_objects():
    """bench.py: both arms describe the workload with the SAME config object; the reference arm runs on the CPU alone and
    prints one JSON line with the keys the driver reads (tiny batch here: C1 at batch 2, one step)."""
    import json
    import os
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, repo)
    import bench
    for name, wl in bench.WORKLOADS.items():
        a = type("A", (), dict(batch=0, iid_depth=False))()
        c1, c8 = bench.config_dict(name, wl, 1, a), bench.config_dict(name, wl, 8, a)
        assert c1["workload"] == wl["title"] and c1["name"] == name
        assert c8["global_batch"] == (wl["batch"] if wl["batch_is"] == "global" else wl["batch"] * 8)
        assert c8["scaling"] == ("strong" if wl["batch_is"] == "global" else "weak")
        assert bench.warped_px(wl, 2) > 0
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "1",
                          "--warmup", "1", "--cpu-batch", "2"], capture_output=True, text=True, timeout=300, cwd=repo)
    assert out.returncode == 0, out.stderr[-400:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["config"] == bench.config_dict("C1", bench.WORKLOADS["C1"], 1, type("A", (), dict(batch=0, iid_depth=False))())
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0


def test_arithmetic_profile_selector():
    """ops.ARITHMETIC picks the rounding profile of the pose chain (include/dvf_b200.h: DVF_ROT_REF_CUDA); anything but the two
    documented names is an error, and the header, the loader and the binding agree on the flag values."""
    from dvf_b200 import _lib, ops
    hdr = open(os.path.join(REPO, "include", "dvf_b200.h")).read()
    assert int(re.search(r"#define\s+DVF_ROT_REF_CUDA\s+(0x[0-9a-fA-F]+)", hdr).group(1), 16) == ops.ROT_REF_CUDA
    assert int(re.search(r"DVF_FLAG_REF_CUDA\s*=\s*(\d+)", hdr).group(1)) == _lib.FLAG_REF_CUDA
    assert ops.ARITHMETIC == "ref_cpu"
    assert ops._rot("euler") == _lib.ROTATION["euler"] and ops._rot("quat") == _lib.ROTATION["quat"]
    try:
        ops.ARITHMETIC = "ref_cuda"
        assert ops._rot("euler") == (_lib.ROTATION["euler"] | ops.ROT_REF_CUDA)
        assert ops._rot("quat") == (_lib.ROTATION["quat"] | ops.ROT_REF_CUDA)
        ops.ARITHMETIC = "fast"
        with pytest.raises(_lib.DvfError):
            ops._rot("euler")
    finally:
        ops.ARITHMETIC = "ref_cpu"
