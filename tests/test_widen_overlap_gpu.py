"""GPU test of the opt-in overlapped schedule of the fused backward + Adam (fgs_debug_set_overlap_adam; csrc/c_abi.cu
backward_impl, csrc/preprocess_backward.cu adam_rest_invisible_kernel): two fused training steps at BASELINE.json's C2
shape under both schedules.  Written without access to a GPU; the kernel logic is covered on the CPU by
tests/test_emulated_pipeline.py::test_emulated_overlapped_adam_schedule_matches_default."""
import pytest
import torch

from tests import util

LRS = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]


@pytest.mark.gpu
@pytest.mark.parametrize('n,w,h', [(5000, 160, 120), (1_000_000, 1920, 1080)])
def test_overlapped_adam_schedule_matches_default(cuda_device, n, w, h):
    import fgs_b200
    from fgs_b200 import _capi, synthetic as syn
    lib = _capi.load()
    sc = syn.make_scene(n, 0).to(cuda_device)
    cam = syn.make_camera(0, w, h)
    st = util.settings_for(cam, device=cuda_device)
    gi = (syn.make_grad_image(w, h) / (3 * w * h)).to(cuda_device)

    def two_steps(overlap):
        lib.fgs_debug_set_overlap_adam(int(overlap))
        try:
            params = [p.clone().requires_grad_(True) for p in sc]
            opt = fgs_b200.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, LRS)], lr=0.0, eps=1e-15)
            g = torch.Generator(device='cpu').manual_seed(5)
            for p in params:  # non-zero moments: the zero-gradient update of invisible rows must be observable
                opt.state[p] = {'step': 4, 'exp_avg': (torch.randn(p.shape, generator=g) * 1e-3).to(cuda_device),
                                'exp_avg_sq': (torch.rand(p.shape, generator=g) * 1e-6 + 1e-9).to(cuda_device)}
            vis = None
            for _ in range(2):
                img = fgs_b200.diff_rasterize_fused(*params, torch.empty(0), fgs_b200.RasterizerSettings(*st), opt)
                img.backward(gi)
                opt.step()
            torch.cuda.synchronize()
            return [p.detach() for p in params], [opt.state[p]['exp_avg'] for p in params], [opt.state[p]['exp_avg_sq'] for p in params]
        finally:
            lib.fgs_debug_set_overlap_adam(0)

    pa, ma, va = two_steps(False)
    pb, mb, vb = two_steps(True)
    for t in range(6):
        assert torch.isfinite(pb[t]).all()
        # the float-atomic order of the 2D gradients differs per run by an ulp; beta (x - g) + g amplifies that by
        # 1 / (1 - beta): 10x in m, 1000x in v
        assert util.rel_err(pb[t] - sc[t], pa[t] - sc[t]) <= 5e-3, t
        assert util.rel_err(mb[t], ma[t]) <= 2e-4 and util.rel_err(vb[t], va[t]) <= 2e-3, t
    # rows whose sh_rest moments are bit-identical in both runs are the invisible ones (zero gradient): there must be
    # many of them and their parameters must be bit-identical too
    same = (mb[5].reshape(n, -1) == ma[5].reshape(n, -1)).all(1) & (vb[5].reshape(n, -1) == va[5].reshape(n, -1)).all(1)
    assert int(same.sum()) >= n // 20
    assert torch.equal(pb[5][same], pa[5][same])
