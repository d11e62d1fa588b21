import sys
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, os.path.join(ROOT,'faster-gaussian-splatting_b200')]
import torch, numpy as np
from tests import emu_driver as E, util
from fgs_b200 import synthetic as syn
lib=E.load(sanitize_threads=True)
n,w,h=120,32,24
sc=syn.make_scene(n,6); cam=syn.make_camera(6,w,h)
params=list(sc); params[0]=params[0].clone(); params[0][::3]+=40.0
st=util.settings_for(cam, bg=(0.1,0.1,0.1)); gi=syn.make_grad_image(w,h,6)
lib.fgs_debug_set_overlap_adam(1)
img,state,scr,inp=E.forward(lib,params,st)
m=[E._copy_aligned(np.full(a.shape,1e-3,np.float32)) for a in inp['p']]; v=[E._copy_aligned(np.full(a.shape,1e-6,np.float32)) for a in inp['p']]
E.backward_adam(lib,img,gi,state,scr,inp,m,v,[3]*6,[1e-3]*6)
print('overlap done', state.n_visible)
# covariance-input mode
c6=torch.rand(n,6)*0.001; c6[:,0]+=0.01; c6[:,3]+=0.01; c6[:,5]+=0.01
p2=[sc.means,c6.contiguous(),None,sc.opacities,sc.sh_coefficients_0,sc.sh_coefficients_rest]
img2,state2,scr2,inp2=E.forward(lib,p2,st); g=E.backward(lib,img2,gi,state2,scr2,inp2)
print('cov done', state2.n_visible)
