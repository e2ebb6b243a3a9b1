"""Host-side cost of the autograd wrapper (wall clock per call, GPU kept busy by queueing)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_3d_reconstruction_b200 import SMPL, synthetic
dev = torch.device("cuda:0")
layer = SMPL.synthetic(0).to(dev)
n = 128
b, p, c = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 1))
gb, gp, gc = (t.clone().requires_grad_() for t in (b, p, c))
gj = torch.randn(n, 24, 3, device=dev); gk = torch.randn(n, 24, 2, device=dev); gv = torch.randn(n, 6890, 3, device=dev)

def wall(fn, it=300):
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return (t1 - t0) / it * 1e6, (t2 - t0) / it * 1e6

def fwd_nograd():
    with torch.no_grad(): layer(b, p, c)
def fwd_grad(): layer(gb, gp, gc)
def fwd_bwd_joints():
    v, j, k = layer(gb, gp, gc)
    torch.autograd.backward([j, k], [gj, gk])
def fwd_bwd_all():
    v, j, k = layer(gb, gp, gc)
    torch.autograd.backward([v, j, k], [gv, gj, gk])
def loss_only():
    j = gj.clone().requires_grad_(); k = gk.clone().requires_grad_()
    (k.abs().mean() + j.pow(2).mean()).backward()
for name, fn in (("forward no_grad", fwd_nograd), ("forward with autograd node", fwd_grad),
                 ("fwd+bwd joints/kp2d (direct grads)", fwd_bwd_joints), ("fwd+bwd all (direct grads)", fwd_bwd_all),
                 ("loss ops alone (abs.mean + pow.mean, fwd+bwd)", loss_only)):
    host, total = wall(fn)
    print(f"{name:48s} host {host:7.1f} us/call   incl. GPU drain {total:7.1f} us/call", flush=True)
