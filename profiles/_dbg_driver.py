import os, sys, math, types, torch
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
from loco_edit_b200.weights import random_state_dict, tiny_arch
from oracle import ddpm_ref, pullback_ref
from test_gpu_driver import _make
g=torch.load('tests/golden/driver_tiny.pt',weights_only=False)
dev=torch.device('cuda:0')
arch=tiny_arch(resolution=32, ch_mult=(1,2), attn_resolutions=(16,), num_res_blocks=1)
sd=random_state_dict(arch, seed=1234, perturb_norm=0.1)
unet=ddpm_ref.RefUNet(arch, sd); sched=pullback_ref.RefScheduler()
torch.manual_seed(g['seed']); d=g['x0'].numel()
v0a,_=torch.linalg.qr(torch.randn(d,2)); v0b,_=torch.linalg.qr(torch.randn(d,3))
noises=[torch.randn(5,3,32,32) for _ in range(40)]
xT=pullback_ref.ddim_inversion(unet, sched, g['x0'])
xt,t,idx=pullback_ref.ddim_forward(unet, sched, xT, 0, 40)
import tempfile
e=_make(dev, tempfile.mkdtemp(), g)
gxT=e.run_DDIMinversion(7)
def rel(a,b): return float((a.cpu()-b).norm()/b.norm())
print('xT rel err', rel(gxT,xT))
gxt,_,_=e.DDIMforwardsteps(gxT,0,40)
print('xt rel err', rel(gxt,xt))
vref=[v for k,v in g['files'].items() if k.endswith('pc_000-vT.pt')][0]
batch=pullback_ref.edit_batch(xt, vref[0], 0.5, 4, 2)
gb=e.build_edit_batch(gxt, vref[0].to(dev), 2)
print('batch rel err', rel(gb,batch))
# eta=0 path from the ORACLE batch
img0=pullback_ref.ddim_forward(unet, sched, batch, 40, -1)
gimg0=e.DDIMforwardsteps(batch.to(dev),40,-1,save_image=False)
print('eta0 final rel err', rel(gimg0,img0))
nz={79+i: noises[i] for i in range(20)}
img1=pullback_ref.ddim_forward(unet, sched, batch, 40, -1, boost_idx=79, noises=nz)
it=iter(noises); e.noise_fn=lambda i,x: next(it).to(dev)
gimg1=e.DDIMforwardsteps(batch.to(dev),40,-1,save_image=False,performance_boosting=True)
print('eta1 final rel err', rel(gimg1,img1), 'ref check', rel(img1, g['finals'][0]))
# step-by-step divergence for eta=0
x_o=batch.clone(); x_g=batch.to(dev)
sched.set_timesteps(100); e.scheduler.set_timesteps(100, device=dev)
for i in range(40, 99):
    tt=sched.timesteps[i]
    with torch.no_grad():
        eo=unet(x_o, tt)
    eg=e.unet(x_g, e.scheduler._ts_host[i])
    if i in (40,41,50,60,70,79,85,90,95,98):
        print(i, 'eps rel', rel(eg,eo), 'x rel', rel(x_g,x_o), 'eps norm', float(eo.norm()), 'x norm', float(x_o.norm()))
    x_o,_=sched.step(eo, tt, x_o, eta=0)
    x_g=e.scheduler.step(eg, e.scheduler._ts_host[i], x_g, eta=0, t_idx=i).prev_sample
