"""``-m gpu``: the fused entries of SURVEY.md 8(f) — the reconstruction loss evaluated in the blend epilogue
(``render_l1_loss``; gs.py:107 + /root/reference/core/loss/whole_loss.py:126-130) against the unfused composition
(our rasteriser + torch elementwise ops) and against the CPU oracle."""
import numpy as np
import pytest
import torch

from gpu_utils import TAN, oracle_forward, scene_tensors, to_dev, view_tensors
from sigman_release_b200 import rasterizer, scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("reduction", ["reference", "mean"])
@pytest.mark.parametrize("with_mask", [False, True])
@pytest.mark.parametrize("hw", [(64, 64), (100, 75)])
def test_fused_l1_loss_matches_unfused_and_oracle(with_mask, hw, reduction):
    """reduction="reference" is the reference's own: loss_l1 = |pred*m - gt*m| unreduced (whole_loss.py:49-50,130),
    then torch.sum(loss_l1) / loss_l1.shape[0] with shape[0] = B*V (whole_loss.py:139)."""
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
                                                                 vmt, pmt, bg, H, W, TAN, TAN, tt, mt,
                                                                 reduction=reduction, exact_exp=True)
    (loss * 2.5).backward()

    b = scene_tensors(sc, requires_grad=True)
    color, radii_b, depth_b, alpha_b = rasterizer.rasterize_batch(b["means3D"], b["cov3D"], b["colors"], b["opacities"],
                                                                  vmt, pmt, bg, H, W, TAN, TAN, exact_exp=True)
    img_b = color.clamp(0, 1)
    m = mt if with_mask else 1.0
    l1 = (img_b * m - tt * m).abs()                                   # whole_loss.py:130 (l1() does not reduce)
    loss_b = torch.sum(l1) / (l1.shape[0] * l1.shape[1]) if reduction == "reference" else l1.mean()
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
    ref /= len(views) if reduction == "reference" else target.size
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
                                    target, None, renders_per_chunk=2, reduction="mean")
    out[0].backward()
    g1 = {k: v.grad.clone() for k, v in t.items()}
    for v in t.values():
        v.grad = None
    out2 = rasterizer.render_l1_loss(t["means3D"], t["cov3D"], t["colors"], t["opacities"], vmt, pmt, bg, H, W, TAN, TAN,
                                     target, None, reduction="mean")
    (out2[0] * 3.0).backward()
    assert abs(float(out[0]) - float(out2[0])) <= 1e-6 * float(out2[0])
    assert abs(float(out[0]) - float((out[1] - target).abs().mean())) <= 2e-6 * float(out[0])
    for k, v in t.items():
        assert float((v.grad - 3.0 * g1[k]).abs().max()) <= 3e-4 * 3.0 * float(g1[k].abs().max()) + 1e-12, k


def _prep_reference(s, R, d2):
    """gs.py:17-38,69-73 with torch ops (fp64) for autograd comparison."""
    scales_ = torch.sqrt(torch.clamp_min(d2, 0.0000001))[..., None].repeat(1, 3).detach()
    scale = (s + 1) * scales_
    Lm = torch.zeros_like(R)
    Lm[:, 0, 0] = scale[:, 0]; Lm[:, 1, 1] = scale[:, 1]; Lm[:, 2, 2] = scale[:, 2]
    cov = R @ (Lm ** 2) @ R.permute(0, 2, 1)
    return torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], 1)


def test_prep_cov3d_kernel_and_backward():
    import oracle
    from sigman_release_b200 import prep_cov3d
    rng = np.random.default_rng(11)
    n = 5000
    s = rng.uniform(-1, 1, (n, 3)).astype(np.float32); R = rng.normal(size=(n, 3, 3)).astype(np.float32)
    d2 = np.concatenate([rng.uniform(1e-6, 1e-4, n - 2), [0.0, 1e-9]]).astype(np.float32)
    st, Rt, dt = to_dev(s).requires_grad_(True), to_dev(R).requires_grad_(True), to_dev(d2)
    cov = prep_cov3d(st, Rt, dt)
    want = oracle.prep_cov3d(s, R, d2, np.float64)
    # off-diagonal entries cancel: tolerance relative to the Gaussian's largest covariance entry
    scale = np.abs(want).max(axis=1, keepdims=True)
    assert float((np.abs(cov.detach().cpu().numpy() - want) / scale).max()) <= 3e-6
    g = rng.normal(size=(n, 6)).astype(np.float32)
    (cov * to_dev(g)).sum().backward()
    s64 = torch.tensor(s, dtype=torch.float64, requires_grad=True); R64 = torch.tensor(R, dtype=torch.float64, requires_grad=True)
    (_prep_reference(s64, R64, torch.tensor(d2, dtype=torch.float64)) * torch.tensor(g, dtype=torch.float64)).sum().backward()
    for got, ref, name in ((st.grad, s64.grad, "scale"), (Rt.grad, R64.grad, "rotation")):
        err = (got.cpu().double() - ref).abs().max()
        assert float(err) <= 1e-5 * float(ref.abs().max()), (name, float(err))
    # bf16-autocast emulation: every output is a bf16 value and within bf16 rounding of the fp32 result
    cov_b = prep_cov3d(st.detach(), Rt.detach(), dt, bf16_autocast=True)
    assert torch.equal(cov_b, cov_b.bfloat16().float())
    rel = ((cov_b - cov.detach()).abs() / (cov.detach().abs() + 1e-12))
    assert float(rel.median()) < 1e-2


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_colors_kernel_and_backward(deg):
    import oracle
    from sigman_release_b200 import sh_colors
    rng = np.random.default_rng(20 + deg)
    n, K = 3000, 16
    m = rng.normal(size=(n, 3)).astype(np.float32); sh = (rng.normal(size=(n, K, 3)) * 0.4).astype(np.float32)
    cp = np.array([0.3, -0.2, 2.5], np.float32)
    mt, sht = to_dev(m).requires_grad_(True), to_dev(sh).requires_grad_(True)
    col = sh_colors(mt, sht, to_dev(cp), deg)
    want, clamped = oracle.sh_colors(m, sh, cp, deg, np.float64)
    np.testing.assert_allclose(col.detach().cpu().numpy(), want, atol=3e-6)
    g = rng.normal(size=(n, 3)).astype(np.float32)
    (col * to_dev(g)).sum().backward()
    # reference gradients: fp64 autograd of an independent torch restatement
    from test_oracle import _torch_sh
    m64 = torch.tensor(m, dtype=torch.float64, requires_grad=True); sh64 = torch.tensor(sh, dtype=torch.float64, requires_grad=True)
    (_torch_sh(m64, sh64, torch.tensor(cp, dtype=torch.float64), deg) * torch.tensor(g, dtype=torch.float64)).sum().backward()
    near = np.abs(want) < 1e-5                                   # channels sitting on the clamp: either branch is fine
    keep = ~near.any(axis=1)
    for got, ref, name in ((mt.grad, m64.grad, "means3D"), (sht.grad, sh64.grad, "shs")):
        ref = torch.zeros_like(got.cpu().double()) if ref is None else ref     # degree 0 does not depend on the view
        err = (got.cpu().double()[keep] - ref[keep]).abs().max()
        assert float(err) <= 2e-5 * float(ref.abs().max()) + 1e-7, (name, float(err))


def test_rasterizer_module_accepts_shs():
    """GaussianRasterizer.forward(shs=...) == forward(colors_precomp=sh_colors(...)) and gradients reach shs."""
    from sigman_release_b200 import GaussianRasterizationSettings, GaussianRasterizer, sh_colors
    from sigman_release_b200 import cameras
    sc = scenes.body_gaussians(2000, seed=4)
    rng = np.random.default_rng(4)
    N = sc["means3D"].shape[0]
    sh = to_dev((rng.normal(size=(N, 16, 3)) * 0.3).astype(np.float32)).requires_grad_(True)
    vm, pm, cp = cameras.orbit_cameras([45])
    settings = GaussianRasterizationSettings(image_height=64, image_width=64, tanfovx=TAN, tanfovy=TAN,
                                             bg=torch.ones(3, device="cuda"), scale_modifier=1.0, viewmatrix=to_dev(vm[0]),
                                             projmatrix=to_dev(pm[0]), sh_degree=2, campos=to_dev(cp[0]), prefiltered=False,
                                             debug=False)
    r = GaussianRasterizer(settings)
    means = to_dev(sc["means3D"]); op = to_dev(sc["opacities"]).reshape(-1, 1); cov = to_dev(sc["cov3D"])
    img, radii, depth, alpha = r(means3D=means, means2D=torch.zeros_like(means), opacities=op, shs=sh, cov3D_precomp=cov)
    col = sh_colors(means, sh.detach(), to_dev(cp[0]), 2)
    img2 = r(means3D=means, means2D=torch.zeros_like(means), opacities=op, colors_precomp=col, cov3D_precomp=cov)[0]
    assert torch.equal(img, img2)
    img.sum().backward()
    assert sh.grad is not None and float(sh.grad[:, :9].abs().sum()) > 0 and float(sh.grad[:, 9:].abs().sum()) == 0


def test_step_captured_in_a_cuda_graph_replays_correctly():
    """The launch path has no host read / blocking wait, so forward + backward can be captured in a CUDA graph
    (sigman_release_b200.GraphedStep): replays with refreshed static inputs reproduce the eager results."""
    from sigman_release_b200 import GraphedStep, graph_status
    H = W = 96
    views = [30, 65]
    vmt, pmt, _, _ = view_tensors(views)
    bg = torch.ones(3, device="cuda")
    target = torch.rand((1, len(views), 3, H, W), device="cuda")
    scenes_np = [scenes.body_gaussians(5000, seed=s) for s in (1, 2)]
    static = scene_tensors(scenes_np[0], requires_grad=True)

    def step():
        for v in static.values():                              # backward must ASSIGN the (static) .grad tensors, not add to them
            v.grad = None
        loss = rasterizer.render_l1_loss(static["means3D"], static["cov3D"], static["colors"], static["opacities"], vmt, pmt,
                                         bg, H, W, TAN, TAN, target)[0]
        loss.backward()
        return loss

    graphed = GraphedStep(step)
    for sc in scenes_np[::-1] + scenes_np:                   # refresh the static inputs in place, replay, compare to eager
        fresh = scene_tensors(sc, requires_grad=True)
        with torch.no_grad():
            for k in static:
                static[k].copy_(fresh[k])
        loss_g = graphed.replay()
        torch.cuda.synchronize()
        assert graph_status()["overflow"] == 0
        loss_e = rasterizer.render_l1_loss(fresh["means3D"], fresh["cov3D"], fresh["colors"], fresh["opacities"], vmt, pmt,
                                           bg, H, W, TAN, TAN, target)[0]
        loss_e.backward()
        assert abs(float(loss_g) - float(loss_e)) <= 1e-6 * abs(float(loss_e))
        for k in static:
            ge, gg = fresh[k].grad, static[k].grad
            assert float((ge - gg).abs().max()) <= 3e-4 * float(ge.abs().max()) + 1e-12, k
