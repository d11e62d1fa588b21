"""4D Gaussians on the GPU (BASELINE.json configs[4]; the FasterGS4D branch is not in the reference
tree: parity unpinned, oracle/fourd_oracle.py restates Yang et al., ICLR 2024).  The same comparison runs on the CPU
against the emulated kernels (tests/emu_python_layer.py::check_fourd_api); written without access to a GPU."""
import pytest
import torch

from tests import util


def make_scene_4d(n, seed, width):
    from fgs_b200 import synthetic as syn
    sc = syn.make_scene(n, seed, width=width)
    g = torch.Generator().manual_seed(seed + 77)
    means4 = torch.cat([sc.means, torch.rand(n, 1, generator=g)], dim=1).contiguous()
    scales4 = torch.cat([sc.scales, torch.randn(n, 1, generator=g) * 0.3 - 1.2], dim=1).contiguous()
    rot_l = torch.randn(n, 4, generator=g).contiguous()
    rot_r = (torch.randn(n, 4, generator=g) * 0.3 + torch.tensor([1.0, 0.0, 0.0, 0.0])).contiguous()
    return [means4, scales4, rot_l, rot_r, sc.opacities, sc.sh_coefficients_0, sc.sh_coefficients_rest]


@pytest.mark.gpu
def test_fourd_forward_backward_vs_oracle(cuda_device):
    import fgs_b200
    from fgs_b200 import fourd as F4, synthetic as syn
    from oracle import fourd_oracle as FO, torch_oracle as O
    n, w, h, t = 1000, 128, 96, 0.4
    params = make_scene_4d(n, 5, w)
    cam = syn.make_camera(5, w, h)
    st_cpu = util.settings_for(cam, bg=(0.05, 0.1, 0.15))
    st = util.settings_for(cam, bg=(0.05, 0.1, 0.15), device=cuda_device)
    gi = syn.make_grad_image(w, h, 5)
    leaves = [p.to(cuda_device).clone().requires_grad_(True) for p in params]
    img = F4.diff_rasterize_4d(*leaves, t, torch.empty(0), fgs_b200.RasterizerSettings(*st))
    img.backward(gi.to(cuda_device))
    o_leaves = [p.double().clone().requires_grad_(True) for p in params]
    s = O.Settings(*st_cpu)
    s = s._replace(w2c=s.w2c.double(), cam_position=s.cam_position.double(), bg_color=s.bg_color.double())
    out = FO.render_4d(*o_leaves, t, s)
    out.image.backward(gi.double())
    assert (img.detach().cpu().double() - out.image.detach()).abs().mean().item() <= 1e-4
    for name, a, b in zip(('means4', 'scales4', 'rot_l', 'rot_r', 'opacities', 'sh0', 'sh_rest'), leaves, o_leaves):
        e = util.rel_err(a.grad.cpu(), b.grad)
        assert e <= 5e-3, (name, e)


@pytest.mark.gpu
def test_fourd_c5_shape_properties(cuda_device):
    """BASELINE.json configs[4] shape (500 k 4D Gaussians, 1080p): finite images and gradients over several time steps, a
    Gaussian's contribution fades away from its own time, inference == training forward."""
    import fgs_b200
    from fgs_b200 import fourd as F4, synthetic as syn
    n, w, h = 500_000, 1920, 1080
    params = [p.to(cuda_device) for p in make_scene_4d(n, 1, w)]
    cam = syn.make_camera(1, w, h)
    rs = fgs_b200.RasterizerSettings(*util.settings_for(cam, device=cuda_device))
    prev = None
    for t in (0.1, 0.5, 0.9):
        leaves = [p.clone().requires_grad_(True) for p in params]
        img = F4.diff_rasterize_4d(*leaves, t, torch.empty(0), rs)
        img.mean().backward()
        assert torch.isfinite(img).all() and all(torch.isfinite(p.grad).all() for p in leaves)
        inf = F4.rasterize_4d(*params, t, rs, to_chw=True, clamp_output=False)
        assert (inf - img.detach()).abs().mean().item() <= 1e-6
        if prev is not None:
            assert (img.detach() - prev).abs().mean().item() > 1e-4  # the scene changes over time
        prev = img.detach()
    # temporal marginal: far from every Gaussian's time the image is the background
    far = F4.rasterize_4d(*params, 50.0, rs, to_chw=True, clamp_output=False)
    assert float(far.abs().max()) <= 1e-6
