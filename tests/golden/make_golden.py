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


if __name__ == '__main__':
    main()
