"""Where does the GPU idle inside one fused training step?  Uses the library's per-kernel events."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'faster-gaussian-splatting_b200')); sys.path.insert(0, str(ROOT))
import fgs_b200
from fgs_b200 import synthetic as syn, _capi
from tests import util
dev = torch.device('cuda'); n, w, h = 1_000_000, 1920, 1080
sc = syn.make_scene(n, 0); cams = [syn.make_camera(k, w, h) for k in range(8)]
gts = [syn.make_target(w, h, k).to(dev) for k in range(8)]
params = [p.to(dev).requires_grad_(True) for p in sc]
opt = fgs_b200.FusedAdam([{'params': [p], 'lr': 1e-4} for p in params], lr=0.0, eps=1e-15)
sts = [util.settings_for(c, device=dev) for c in cams]; empty = torch.empty(0)
fused = len(sys.argv) < 2 or sys.argv[1] != 'unfused'
def step(k):
    rs = fgs_b200.RasterizerSettings(*sts[k])
    img = fgs_b200.diff_rasterize_fused(*params, empty, rs, opt) if fused else fgs_b200.diff_rasterize(*params, empty, rs)
    loss = (img - gts[k]).abs().mean(); loss.backward(); opt.step(); opt.zero_grad()
for it in range(30): step(it % 8)
torch.cuda.synchronize(); _capi.profile_enable(True)
for it in range(3): step(it % 8)
_capi.profile_read(); recs = list(_capi._last_starts); _capi.profile_enable(False)
prev_end = None; first = None
for name, t0, ms in recs:
    if name == 'preprocess':
        if first is not None: print(f'--- step span {t0 - first:.3f} ms'); 
        first = t0
    gap = (t0 - prev_end) if prev_end is not None else 0.0
    print(f'{name:22s} start {t0:8.3f}  dur {ms:6.3f}  gap before {gap:6.3f}')
    prev_end = t0 + ms
