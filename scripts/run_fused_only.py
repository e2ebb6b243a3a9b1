"""Runs the fused blendshapes+skinning kernel alone (operand images packed once) -- the ncu target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_3d_reconstruction_b200 import SMPL, capi, synthetic
from human_3d_reconstruction_b200 import smpl as ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
layer = SMPL(synthetic.make_model(0), precision="f16").to(dev)
betas, pose, cam = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 1))
coef, A, joints = ops.pose_chain(layer, betas, pose)
h, lib = layer.handle(dev), capi.lib()
wsf = int(lib.smplb200_blend_skin_workspace_bytes(h.ptr, n))
ws = torch.empty(wsf, dtype=torch.uint8, device=dev)
verts = torch.empty((n, 6890, 3), device=dev)
s = torch.cuda.current_stream(dev).cuda_stream
capi.check(lib.smplb200_blend_skin(h.ptr, coef.data_ptr(), A.data_ptr(), n, verts.data_ptr(), ws.data_ptr(), wsf, s), "pack+fused")
torch.cuda.synchronize()
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for _ in range(iters):
    capi.check(lib.smplb200_blend_skin(h.ptr, None, None, n, verts.data_ptr(), ws.data_ptr(), wsf, s), "fused")
en.record()
torch.cuda.synchronize()
print(f"k_fused_tc n={n}: {st.elapsed_time(en) / iters * 1e3:.1f} us per launch")
