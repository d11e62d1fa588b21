import sys, time, math, random
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path[:0]=[ROOT, os.path.join(ROOT,'faster-gaussian-splatting_b200')]
import torch, numpy as np
from tests import emu_driver as E, util
from fgs_b200 import synthetic as syn
lib=E.load(sanitize_threads=True)
rng=random.Random(7); t_end=time.time()+float(sys.argv[1]); it=0
while time.time()<t_end:
    it+=1
    n=rng.choice([3,33,100,200]); w,h=rng.choice([(16,12),(33,25),(48,36),(64,24)])
    sigma=rng.choice([None,0.1,2.0]); seed=rng.randrange(1000); hint=rng.choice([0,1,50])
    sc=syn.make_scene(n,seed,log_scale_mean=None if sigma is None else math.log(sigma)); cam=syn.make_camera(seed,w,h)
    st=util.settings_for(cam, active_sh_bases=rng.choice([1,4,16]), bg=(0.1,0.2,0.3), aa=rng.random()<0.3)
    lib.fgs_debug_set_capacity_hint(hint)
    img,state,scr,inp=E.forward(lib,list(sc),st)
    E.backward(lib,img,syn.make_grad_image(w,h,seed),state,scr,inp,densification_info=np.zeros((2,n),np.float32))
    print('case',it,n,w,h,sigma,hint,state.n_visible,state.n_instances,flush=True)
print('cases',it)
