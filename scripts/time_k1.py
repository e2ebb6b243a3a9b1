import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_3d_reconstruction_b200 import SMPL, synthetic, capi
from human_3d_reconstruction_b200 import smpl as ops
dev = torch.device("cuda:0"); n = int(os.environ.get("N", 4096))
layer = SMPL(synthetic.make_model(0)).to(dev)
b, p, c = synthetic.make_inputs(n, 1)
tb, tp = torch.from_numpy(b).to(dev), torch.from_numpy(p).to(dev)
coef, A, j = ops.pose_chain(layer, tb, tp)
h = layer.handle(dev); lib = capi.lib(); s = torch.cuda.current_stream(dev).cuda_stream
vposed = torch.empty((n, 3, h.padded_verts), device=dev)
for prec in sys.argv[1:]:
    flags = capi.make_flags(precision=prec)
    wsb = int(lib.smplb200_blendshapes_workspace_bytes(h.ptr, n, flags))
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    fn = lambda: lib.smplb200_blendshapes(h.ptr, coef.data_ptr(), n, vposed.data_ptr(), ws.data_ptr(), wsb, flags, s)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(30): fn()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"dbg={os.environ.get('SMPLB200_DBG','0')} {prec}: gpu {e0.elapsed_time(e1)/30*1e3:.1f} us/call, host issue {(t1-t0)/30*1e6:.1f} us/call")
