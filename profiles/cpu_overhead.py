"""Host-side cost of one training step through the Python operator API: run the bench loop on a tiny
scene (GPU time negligible) and report wall-clock per step == CPU enqueue cost."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'faster-gaussian-splatting_b200')); sys.path.insert(0, str(ROOT))
from fgs_b200 import synthetic as syn
from tests import util
impls = {}
import fgs_b200
impls['mine'] = fgs_b200
ref = util.load_reference()
if ref is not None:
    impls['reference'] = ref
dev = torch.device('cuda')
n, w, h = 2000, 128, 96
sc = syn.make_scene(n, 0); cam = syn.make_camera(0, w, h)
gt = syn.make_target(w, h).to(dev)
for name, be in impls.items():
    for fused in ([False, True] if name == 'mine' else [False]):
        params = [p.to(dev).requires_grad_(True) for p in sc]
        opt = be.FusedAdam([{'params': [p], 'lr': 1e-4} for p in params], lr=0.0, eps=1e-15)
        st = util.settings_for(cam, device=dev)
        empty = torch.empty(0)
        def step():
            rs = be.RasterizerSettings(*st)
            img = be.diff_rasterize_fused(*params, empty, rs, opt) if fused else be.diff_rasterize(*params, empty, rs)
            loss = (img - gt).abs().mean(); loss.backward(); opt.step(); opt.zero_grad()
        for _ in range(20): step()
        torch.cuda.synchronize(); t = time.perf_counter()
        K = 300
        for _ in range(K): step()
        torch.cuda.synchronize()
        print(f'{name} fused={fused}: {(time.perf_counter() - t) / K * 1e3:.3f} ms per step on a tiny scene (host cost)')
