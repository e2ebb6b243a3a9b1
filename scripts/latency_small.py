"""Small-batch latency of the forward (configs[0]/[1]): per-call GPU time at N = 1, 8, 64, 256."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_3d_reconstruction_b200 import SMPL, synthetic
dev = torch.device("cuda:0")
model = synthetic.make_model(0)
for n in (1, 8, 64, 256, 1024):
    b, p, c = synthetic.make_inputs(n, 1)
    tb, tp, tc = (torch.from_numpy(x).to(dev) for x in (b, p, c))
    for kw in (dict(precision="fp32", lbs="fma"), dict(precision="auto", lbs="auto"), dict(precision="bf16x3", lbs="tc")):
        layer = SMPL(model, **kw).to(dev)
        with torch.no_grad():
            for _ in range(5): layer(tb, tp, tc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter(); e0.record()
            for _ in range(50): layer(tb, tp, tc)
            e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"N={n:5d} {kw['precision']:>6}/{kw['lbs']:<4}: gpu {e0.elapsed_time(e1)/50*1e3:8.1f} us/call   wall {(t1-t0)/50*1e6:8.1f} us/call   {n/(e0.elapsed_time(e1)/50*1e-3)/1e6:7.3f} M bodies/s")

from human_3d_reconstruction_b200 import GraphedSMPL
for n in (1, 64, 256):
    for kw in (dict(precision="fp32", lbs="fma"), dict(precision="bf16x3", lbs="tc")):
        layer = SMPL(model, **kw).to(dev)
        g = GraphedSMPL(layer, n, dev)
        for _ in range(5): g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(200): g.replay()
        e1.record(); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"GRAPH N={n:5d} {kw['precision']:>6}/{kw['lbs']:<4}: gpu {e0.elapsed_time(e1)/200*1e3:8.1f} us/replay   wall {(t1-t0)/200*1e6:8.1f} us/replay")
