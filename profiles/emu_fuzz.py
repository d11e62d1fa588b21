import sys, time, math, random, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/faster-gaussian-splatting_b200')
import torch, numpy as np
from tests import emu_driver as E, util
from fgs_b200 import synthetic as syn
from oracle import torch_oracle as O
lib=E.load()
rng=random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 0)
t_end=time.time()+float(sys.argv[2]) if len(sys.argv)>2 else time.time()+600
it=0
while time.time()<t_end:
    it+=1
    n=rng.choice([1,2,3,17,33,64,100,257,400])
    w=rng.choice([1,7,16,17,33,64,96,130]); h=rng.choice([1,5,12,13,25,48,72,100])
    seed=rng.randrange(1000)
    sigma=rng.choice([None,0.02,0.1,0.5,2.0])
    bases=rng.choice([1,4,9,16]); aa=rng.random()<0.3
    hint=rng.choice([0,0,1,50])
    sc=syn.make_scene(n,seed,log_scale_mean=None if sigma is None else math.log(sigma)); cam=syn.make_camera(seed,w,h)
    st=util.settings_for(cam, active_sh_bases=bases, bg=(rng.random(),rng.random(),rng.random()), aa=aa)
    gi=syn.make_grad_image(w,h,seed)
    tag=f'n={n} w={w} h={h} seed={seed} sigma={sigma} bases={bases} aa={aa} hint={hint}'
    try:
        lib.fgs_debug_set_capacity_hint(hint)
        img,state,scr,inp=E.forward(lib,list(sc),st)
        dens=np.zeros((2,n),dtype=np.float32)
        grads=E.backward(lib,img,gi,state,scr,inp,densification_info=dens)
        out,ograds=O.render_and_grads(list(sc),O.Settings(*st),gi,dtype=torch.float64)
        l1=float((torch.from_numpy(np.array(img)).double()-out.image.detach()).abs().mean())
        errs=[util.rel_err(g,og) if float(og.abs().sum())>0 else float(g.abs().sum()) for g,og in zip(grads,ograds)]
        bad = (not np.isfinite(np.array(img)).all()) or l1>2e-5 or max(errs)>2e-3 or state.n_visible!=int(out.proj.visible.sum())
        print(('BAD ' if bad else 'ok  ')+tag+f' V={state.n_visible} I={state.n_instances} l1={l1:.2e} maxerr={max(errs):.2e}', flush=True)
    except Exception as e:
        print('EXC '+tag+' '+repr(e), flush=True); traceback.print_exc()
print('iterations', it)
