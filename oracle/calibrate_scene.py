"""Calibration of the synthetic scene's Gaussian size (I/V in [3,6], 20-40
alpha>=1/255 Gaussians per pixel).  TEST INFRASTRUCTURE ONLY.  Usage:
    python oracle/calibrate_scene.py N W H sigma_at_1M [sigma ...]
Prints V, I, I/V and the estimated number of covering Gaussians per pixel (sum of the
alpha>=1/255 ellipse areas / pixels; the dense count is not feasible at 1M x 2M)."""
import math
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'faster-gaussian-splatting_b200'))
sys.path.insert(0, str(ROOT))
from fgs_b200 import synthetic as syn  # noqa: E402
from oracle import torch_oracle as O  # noqa: E402


def stats(n, w, h, sigma_1m, seed=0):
    sc = syn.make_scene(n, seed, log_scale_mean=math.log(sigma_1m * math.sqrt(1e6 / n)))
    cam = syn.make_camera(seed, w, h)
    s = O.Settings(cam.w2c, cam.position, torch.zeros(3), 16, w, h, cam.focal_x, cam.focal_y,
                   cam.center_x, cam.center_y, cam.near_plane, cam.far_plane, False)
    with torch.no_grad():
        p = O.preprocess(*sc.params(), s)
    vis = p.visible
    V = int(vis.sum())
    I = int(p.n_touched_tiles[vis].sum())
    A, B, C = p.conic[vis, 0], p.conic[vis, 1], p.conic[vis, 2]
    tau = torch.log(p.opacity[vis] * 255.0)
    area = 2.0 * math.pi * tau / torch.sqrt((A * C - B * B).clamp_min(1e-12))
    bw = (p.screen_bounds[vis, 1] - p.screen_bounds[vis, 0]).float()
    bh = (p.screen_bounds[vis, 3] - p.screen_bounds[vis, 2]).float()
    area = torch.minimum(area, bw * bh)
    return V, I, I / max(V, 1), float(area.sum()) / (w * h)


if __name__ == '__main__':
    n, w, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    for sg in sys.argv[4:]:
        V, I, r, cov = stats(n, w, h, float(sg))
        print(f'sigma_at_1M={sg}: V={V} I={I} I/V={r:.2f} covering/pixel~{cov:.1f}')
