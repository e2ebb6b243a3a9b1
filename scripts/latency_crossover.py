import os, sys, time, torch
sys.path.insert(0, "/root/repo")
from human_3d_reconstruction_b200 import SMPL, synthetic, GraphedSMPL
dev = torch.device("cuda:0")
model = synthetic.make_model(0)
for n in [int(x) for x in os.environ.get("NS", "8,16,32,48,64,96,128,192,256").split(",")]:
    row = []
    for kw in (dict(precision="fp32", lbs="fma"), dict(precision="bf16x3", lbs="tc"), dict(precision="bf16x3", lbs="fma"), dict(precision="fp32", lbs="tc")):
        layer = SMPL(model, **kw).to(dev)
        g = GraphedSMPL(layer, n, dev)
        for _ in range(5): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): g.replay()
        e1.record(); torch.cuda.synchronize()
        row.append(f"{kw['precision']}/{kw['lbs']} {e0.elapsed_time(e1)/200*1e3:6.1f}")
    print(f"N={n:4d}  " + "   ".join(row), flush=True)
