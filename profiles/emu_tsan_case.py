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
# round 2: the deferred SH schedule (geometry-only preprocess, colour kernel, deferred preprocess backward, adam_rest_kernel)
# and the hit list of the blend backward, two steps over two views
params=[t.clone() for t in sc]; mm=vv=None
for step in (1,2):
    cam2=syn.make_camera(step,w,h); st2=util.settings_for(cam2, bg=(0.1,0.1,0.1)); gi2=syn.make_grad_image(w,h,step)
    img3,state3,scr3,inp3=E.forward(lib,params,st2,defer_color=True)
    if mm is None:
        mm=[E._copy_aligned(np.zeros_like(a)) for a in inp3['p']]; vv=[E._copy_aligned(np.zeros_like(a)) for a in inp3['p']]
    E.backward_adam(lib,img3,gi2,state3,scr3,inp3,mm,vv,[step]*6,[1e-3]*6,defer_rest=True)
    params=[torch.from_numpy(np.array(a)) for a in inp3['p']]
print('deferred done', state3.n_visible)
