"""Runs the unfused skinning entry point (k_pack_a + k_lbs_tc) alone with preallocated buffers (A/B timing / ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from human_3d_reconstruction_b200 import SMPL, capi, synthetic
from human_3d_reconstruction_b200 import smpl as ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda:0")
layer = SMPL(synthetic.make_model(0), precision="f16x3", lbs="tc").to(dev)
betas, pose, cam = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 1))
coef, A, joints = ops.pose_chain(layer, betas, pose)
vposed = ops.blendshapes(layer, coef)
h, lib = layer.handle(dev), capi.lib()
flags = layer.flags
wsl = int(lib.smplb200_lbs_workspace_bytes(h.ptr, n, flags))
ws = torch.empty(max(wsl, 256), dtype=torch.uint8, device=dev)
verts = torch.empty((n, 6890, 3), device=dev)
s = torch.cuda.current_stream(dev).cuda_stream
def call():
    capi.check(lib.smplb200_lbs(h.ptr, vposed.data_ptr(), A.data_ptr(), n, verts.data_ptr(), None, None, None,
                                ws.data_ptr(), wsl, flags, s), "k3")
for _ in range(3):
    call()
torch.cuda.synchronize()
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for _ in range(iters):
    call()
en.record()
torch.cuda.synchronize()
print(f"smplb200_lbs (k_pack_a + k_lbs_tc) n={n}: {st.elapsed_time(en) / iters * 1e3:.1f} us per call")
