"""``-m gpu``: the fused entries of SURVEY.md 8(f) — the reconstruction loss evaluated in the blend epilogue
(``render_l1_loss``; gs.py:107 + /root/reference/core/loss/whole_loss.py:126-130) against the unfused composition
(our rasteriser + torch elementwise ops) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from gpu_utils import TAN, oracle_forward, scene_tensors, to_dev, view_tensors
from sigman_release_b200 import rasterizer, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("with_mask", [False, True])
@pytest.mark.parametrize("hw", [(64, 64), (100, 75)])
def test_fused_l1_loss_matches_unfused_and_oracle(with_mask, hw):
    H, W = hw
    views = [30, 65, 8]
    sc = scenes.body_gaussians(4000, seed=3)
    sc["colors"] = (sc["colors"] * 1.3 - 0.1).astype(np.float32)       # some pixels saturate the clamp on both sides
    rng = np.random.default_rng(5)
    target = rng.uniform(0, 1, (1, len(views), 3, H, W)).astype(np.float32)
    mask = (rng.uniform(0, 1, (1, len(views), 1, H, W)) > 0.3).astype(np.float32) if with_mask else None
    bg = to_dev(np.array([1.0, 0.9, 1.2], np.float32))
    vmt, pmt, vm, pm = view_tensors(views)
    tt, mt = to_dev(target), (to_dev(mask) if with_mask else None)

    a = scene_tensors(sc, requires_grad=True)
    loss, image, radii, depth, alpha = rasterizer.render_l1_loss(a["means3D"], a["cov3D"], a["colors"], a["opacities"],
                                                                 vmt, pmt, bg, H, W, TAN, TAN, tt, mt)
    (loss * 2.5).backward()

    b = scene_tensors(sc, requires_grad=True)
    color, radii_b, depth_b, alpha_b = rasterizer.rasterize_batch(b["means3D"], b["cov3D"], b["colors"], b["opacities"],
                                                                  vmt, pmt, bg, H, W, TAN, TAN)
    img_b = color.clamp(0, 1)
    m = mt if with_mask else 1.0
    loss_b = (img_b * m - tt * m).abs().mean()
    (loss_b * 2.5).backward()

    assert torch.equal(image, img_b) and torch.equal(depth, depth_b) and torch.equal(alpha, alpha_b)
    assert torch.equal(radii, radii_b)
    assert abs(float(loss) - float(loss_b)) <= 2e-6 * abs(float(loss_b))
    for k in a:
        ga, gb = a[k].grad, b[k].grad
        assert float((ga - gb).abs().max()) <= 3e-4 * float(gb.abs().max()) + 1e-12, k

    # and against the oracle (fp64 loss of the fp32 images)
    ref = 0.0
    for v in range(len(views)):
        _, ora = oracle_forward(sc, vm[v], pm[v], H, W, bg=(1.0, 0.9, 1.2))
        mm = mask[0, v] if with_mask else 1.0
        ref += np.abs(np.clip(ora.color, 0, 1).astype(np.float64) * mm - target[0, v] * mm).sum()
    ref /= target.size
    assert abs(float(loss) - ref) <= 2e-6 * ref


def test_fused_loss_chunked_batches_and_scale_pointer():
    """B x V larger than one chunk: the loss is the sum over chunks, gradients scale with the upstream gradient."""
    H = W = 48
    views = [30, 37, 45]
    B = 3
    scs = [scenes.body_gaussians(1500, seed=s) for s in range(B)]
    t = {k: torch.cat([scene_tensors(sc)[k] for sc in scs]).requires_grad_(True) for k in ("means3D", "cov3D", "colors", "opacities")}
    vmt, pmt, _, _ = view_tensors(views)
    vmt, pmt = vmt.repeat(B, 1, 1, 1), pmt.repeat(B, 1, 1, 1)
    target = torch.rand((B, len(views), 3, H, W), device="cuda")
    bg = torch.ones(3, device="cuda")
    out = rasterizer.render_l1_loss(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt, bg, H, W, TAN, TAN,
                                    target, None, renders_per_chunk=2)
    out[0].backward()
    g1 = {k: v.grad.clone() for k, v in t.items()}
    for v in t.values():
        v.grad = None
    out2 = rasterizer.render_l1_loss(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt, bg, H, W, TAN, TAN,
                                     target, None)
    (out2[0] * 3.0).backward()
    assert abs(float(out[0]) - float(out2[0])) <= 1e-6 * float(out2[0])
    assert abs(float(out[0]) - float((out[1] - target).abs().mean())) <= 2e-6 * float(out[0])
    for k, v in t.items():
        assert float((v.grad - 3.0 * g1[k]).abs().max()) <= 3e-4 * 3.0 * float(g1[k].abs().max()) + 1e-12, k
