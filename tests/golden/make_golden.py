"""Golden vectors from the UNMODIFIED compiled reference (baseline/_ref), generated on a B200.

    gpurun -- python tests/golden/make_golden.py        # writes gpurun_out/golden/*.npz
    cp gpurun_out/golden/*.npz tests/golden/            # commit

The reference repository ships no tests or fixtures, so these files are the pin for
the CPU oracle (tests/test_oracle_cpu.py) and for the CUDA path when baseline/_ref is not on the box.
Inputs are not stored: they are regenerated from the seeds by fgs_b200.synthetic."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'faster-gaussian-splatting_b200'))
from fgs_b200 import synthetic as syn  # noqa: E402
from tests import util  # noqa: E402

CASES = {
    # name: (n, w, h, seed, active_sh_bases, bg, aa)
    'c1_deg3': (1000, 128, 128, 0, 16, (0.1, 0.2, 0.3), False),
    'c1_deg1': (1000, 128, 128, 0, 4, (0.0, 0.0, 0.0), False),
    'c1_deg0': (1000, 128, 128, 0, 1, (0.0, 0.0, 0.0), False),
    'c1_aa': (1000, 128, 128, 1, 16, (0.1, 0.2, 0.3), True),
    'ragged': (300, 101, 67, 2, 9, (1.0, 1.0, 1.0), False),
}


def main():
    ref = util.load_reference()
    assert ref is not None, 'baseline/_ref is missing'
    out_dir = ROOT / 'gpurun_out' / 'golden'
    out_dir.mkdir(parents=True, exist_ok=True)
    for name, (n, w, h, seed, bases, bg, aa) in CASES.items():
        sc = syn.make_scene(n, seed).to('cuda')
        cam = syn.make_camera(seed, w, h)
        st = util.settings_for(cam, active_sh_bases=bases, bg=bg, aa=aa, device='cuda')
        gi = syn.make_grad_image(w, h, seed).cuda()
        params = [p.clone().requires_grad_(True) for p in sc]
        dens = torch.zeros(2, n, device='cuda')
        img = ref.diff_rasterize(*params, dens, ref.RasterizerSettings(*st))
        img.backward(gi)
        inf = ref.rasterize(*sc, ref.RasterizerSettings(*st), to_chw=False)
        # three Adam steps with the training learning rates on the same gradients
        lrs = [1.6e-4, 5e-3, 1e-3, 2.5e-2, 2.5e-3, 1.25e-4]
        opt = ref.FusedAdam([{'params': [p], 'lr': lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15)
        grads = [p.grad.clone() for p in params]
        for _ in range(3):
            opt.step()
        np.savez_compressed(
            out_dir / f'{name}.npz', image=img.detach().cpu().numpy(), inference_hwc=inf.cpu().numpy(),
            densification_info=dens.cpu().numpy(),
            **{f'grad_{k}': g.cpu().numpy() for k, g in zip(['means', 'scales', 'rotations', 'opacities', 'sh0', 'sh_rest'], grads)},
            **{f'adam3_{k}': p.detach().cpu().numpy() for k, p in zip(['means', 'scales', 'rotations', 'opacities', 'sh0', 'sh_rest'], params)},
            meta=np.array([n, w, h, seed, bases, int(aa)]), bg=np.array(bg, dtype=np.float32))
        print(name, 'image mean', float(img.mean()), 'grad means L1', float(grads[0].abs().sum()))


def aux_inputs(device='cpu'):
    """Seeded inputs of the three small operators (shared by the golden dump and the tests)."""
    g = torch.Generator().manual_seed(11)
    n = 4096
    sc = syn.make_scene(n, 11)
    cam = syn.make_camera(3, 640, 360)
    d = dict(positions=sc.means, w2c=cam.w2c, filter_3d=torch.rand(n, 1, generator=g) * 0.02 + 1e-3,
             old_opacities=torch.rand(n, 1, generator=g) * 0.98 + 0.01, old_scales=torch.rand(n, 3, generator=g) * 0.1,
             n_samples=torch.randint(0, 60, (n,), generator=g), raw_scales=sc.scales, raw_rotations=sc.rotations.clone(),
             raw_opacities=sc.opacities, noise=torch.randn(n, 3, generator=g), means=sc.means.clone())
    d['raw_rotations'][::97] = 0.0
    filt = dict(width=640, height=360, focal_x=cam.focal_x, focal_y=cam.focal_y, center_x=cam.center_x,
                center_y=cam.center_y, near_plane=0.2, clipping_tolerance=0.15, distance2filter=0.0011)
    return {k: v.to(device) for k, v in d.items()}, filt


def dump_aux(ref, out_dir):
    d, filt = aux_inputs('cuda')
    f3 = d['filter_3d'].clone()
    vis = torch.zeros(f3.shape[0], 1, dtype=torch.bool, device='cuda')
    ref.update_3d_filter(d['positions'], d['w2c'], f3, vis, **filt)
    new_o, new_s = ref.relocation_adjustment(d['old_opacities'], d['old_scales'], d['n_samples'])
    means = d['means'].clone()
    from FasterGSCudaBackend import _C as ref_C  # add_noise with a fixed noise tensor (the binding draws its own)
    ref_C.add_noise(d['raw_scales'], d['raw_rotations'], d['raw_opacities'], d['noise'], means, 0.8)
    np.savez_compressed(out_dir / 'aux_ops.npz', filter_3d=f3.cpu().numpy(), visibility=vis.cpu().numpy(),
                        new_opacities=new_o.cpu().numpy(), new_scales=new_s.cpu().numpy(), means=means.cpu().numpy())
    print('aux_ops', float(f3.mean()), int(vis.sum()), float(new_o.mean()), float(means.abs().mean()))


def dump_pruning_scores(ref, out_dir):
    n, w, h, seed = 800, 112, 84, 6
    sc = syn.make_scene(n, seed).to('cuda')
    cam = syn.make_camera(seed, w, h)
    bg = (0.2, 0.1, 0.3)
    scores = torch.zeros(n, device='cuda')
    ref.update_pruning_scores(scores, *sc, ref.RasterizerSettings(*util.settings_for(cam, bg=bg, device='cuda')))
    np.savez_compressed(out_dir / 'pruning_scores.npz', scores=scores.cpu().numpy(), meta=np.array([n, w, h, seed]),
                        bg=np.array(bg, dtype=np.float32))
    print('pruning_scores', float(scores.sum()), int((scores > 0).sum()))


if __name__ == '__main__':
    main()
    dump_aux(util.load_reference(), ROOT / 'gpurun_out' / 'golden')
    dump_pruning_scores(util.load_reference(), ROOT / 'gpurun_out' / 'golden')
