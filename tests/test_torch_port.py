"""CPU: oracle/torch_port.py (the torch-op restatement timed as the CPU baseline) reproduces the golden
vectors of the real reference exactly -- it issues the same operator sequence."""
import numpy as np
import pytest
import torch

from helpers import golden, golden_names


@pytest.mark.parametrize("name", golden_names("iw_"))
def test_port_warp_bit_exact(name):
    from oracle import torch_port as tp
    g = golden(name)
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    torch.set_num_threads(1)
    img, depth, pose = t("img").requires_grad_(True), t("depth").requires_grad_(True), t("pose").requires_grad_(True)
    w = tp.warp(img, depth, pose, t("K"), t("Kinv"), g["rotation_mode"], g["padding_mode"], align_corners=name.endswith("_align"))
    assert np.array_equal(w.detach().numpy(), g["warped"])
    w.backward(t("gout"))
    assert np.array_equal(depth.grad.numpy(), g["gdepth"])
    np.testing.assert_allclose(pose.grad.numpy(), g["gpose"], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("name", [n for n in golden_names("sfm_") if "old" not in n])
def test_port_multiscale_loss(name):
    from oracle import torch_port as tp
    g = golden(name)
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    n = int(g["n_scales"])
    depths = [t(f"depth{s}") for s in range(n)]
    masks = [t(f"mask{s}") if bool(g["with_mask"]) else None for s in range(n)]
    loss = tp.loss_multi_scale(t("img_R2"), [t("img_R1"), t("img_L2")], t("intrinsics"), t("intrinsics_inv"), depths, masks,
                               t("pose"), g["rotation_mode"], g["padding_mode"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))


def test_port_two_view_loss():
    from oracle import torch_port as tp
    g = golden("lf_images")
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    loss = tp.loss_two_view(t("img_R2"), t("img_R1"), t("img_L2"), t("depth"), t("T_2to1"), t("T_R2L"), t("intrinsics"),
                            t("intrinsics_inv"))
    assert abs(float(loss) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
