import sys, torch
sys.path.insert(0, "/root/repo")
from human_3d_reconstruction_b200 import SMPL, synthetic
dev = torch.device("cuda:0")
for reg in ("sparse", "dense"):
    model = synthetic.make_model(0, regressor=reg)
    for n in (128, 1024, 4096):
        b, p, c = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 1))
        row = []
        for joints in ("kinematic", "regressed"):
            layer = SMPL(model, joints=joints).to(dev)
            with torch.no_grad():
                for _ in range(5): layer(b, p, c)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50): layer(b, p, c)
                e1.record(); torch.cuda.synchronize()
            row.append(f"{joints} {e0.elapsed_time(e1)/50*1e3:7.1f} us")
        print(f"regressor={reg:6s} N={n:5d}  " + "   ".join(row), flush=True)
