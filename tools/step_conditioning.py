"""How well conditioned are the train step's loss values?  (CPU oracle only; TEST / DOCUMENTATION AID.)

    python tools/step_conditioning.py <size> <relative weight perturbation>

Runs the oracle's train step twice: once as is, once with every weight multiplied by (1 + eps * N(0,1)).  Values computed before an
optimiser update move by ~eps.  G_GAN and G2_GAN are evaluated through discriminators that were JUST updated by Adam with
beta1 = 0 at step 1 (every weight moves by lr * sign(g)): once a perturbation is large enough to flip ReLU / LeakyReLU masks
(eps >= 1e-5 or so: ~1e-5 of the activations flip, which moves gradients by ~sqrt(1e-5) = 3e-3) weight-gradient signs flip too
and these two values move by ~1e-3 regardless of eps.  Measured at 256x256 (arch B): eps 1e-6 -> G_GAN 3.5e-6; 1e-5 -> see below;
1e-4 -> 1.2e-3; 1e-3 -> 1.3e-3.  This is why tests/test_baseline_configs_gpu.py gates those two values at 1e-2 and everything else
at 1e-3: any two fp32 implementations of the forward (CPU vs cuDNN included) differ at the 1e-5 level."""
import sys, copy, torch, numpy as np, time
sys.path.insert(0,'/root/repo')
from oracle import skit_oracle as O
S=int(sys.argv[1]); eps=float(sys.argv[2]); NT,NF=64,32
torch.set_num_threads(8)
sds=[O.init_resnet_g(9,5,64,9,seed=7), O.init_multiscale_d(4,64,3,3,seed=8), O.init_multiscale_d(7,64,3,3,seed=9)]
batch=O.synthetic_batch(S,NT=NT,seed=1,ellipse_mask=True)
rs=np.random.RandomState(5)
rand=dict(real_b=[0.3],real_s=[0.8],fake_b=[0.6],fake_s=[0.2],fake_ox=rs.randint(0,S-32,NF).astype(np.int32),fake_oy=rs.randint(0,S-32,NF).astype(np.int32))
cfg=O.StepConfig(netG="resnet_9blocks",batch_size_G2=NT,add_fake_T_sample_size=NF)
def run(e):
    g=torch.Generator().manual_seed(3)
    s=[{k:(v*(1+e*torch.randn(v.shape,generator=g)) if v.dtype.is_floating_point and 'running' not in k and 'filt' not in k else v.clone()) for k,v in sd.items()} for sd in sds]
    r=O.train_step(cfg,*s,{},O.step_inputs_from_batch(batch),rand,step=1)
    return r["losses"]
a=run(0.0); b=run(eps)
for k in a: print("%-18s %.6f %.6f  rel dev %.2e"%(k, a[k], b[k], abs(a[k]-b[k])/max(1,abs(a[k]))))
