"""Runs the small-batch forward a few times (for an ncu launch list): N=64 and N=1, fp32 FMA path."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_3d_reconstruction_b200 import SMPL, synthetic
dev = torch.device("cuda:0")
model = synthetic.make_model(0)
for n in (64, 1):
    b, p, c = synthetic.make_inputs(n, 1)
    tb, tp, tc = (torch.from_numpy(x).to(dev) for x in (b, p, c))
    layer = SMPL(model, precision="fp32", lbs="fma").to(dev)
    for _ in range(4):
        layer(tb, tp, tc)
    torch.cuda.synchronize()
