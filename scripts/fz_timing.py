"""Timing build (-DSMPLB200_FZ_TIMING): cycles each role of k_fused_tc spends in each of its waits, per unit."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from human_3d_reconstruction_b200 import SMPL, capi, synthetic
from human_3d_reconstruction_b200 import smpl as ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = torch.device("cuda:0")
layer = SMPL(synthetic.make_model(0), precision="f16").to(dev)
betas, pose, cam = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 1))
coef, A, joints = ops.pose_chain(layer, betas, pose)
h, lib = layer.handle(dev), capi.lib()
wsf = int(lib.smplb200_blend_skin_workspace_bytes(h.ptr, n))
ws = torch.empty(wsf, dtype=torch.uint8, device=dev)
verts = torch.empty((n, 6890, 3), device=dev)
s = torch.cuda.current_stream(dev).cuda_stream
dbg = C.CDLL(capi.LIB_PATH)
out = (C.c_longlong * (148 * 32))()
capi.check(lib.smplb200_blend_skin(h.ptr, coef.data_ptr(), A.data_ptr(), n, verts.data_ptr(), ws.data_ptr(), wsf, s), "pack+fused")
dbg.smplb200_debug_fz_timing(out)     # discard the cold run
capi.check(lib.smplb200_blend_skin(h.ptr, None, None, n, verts.data_ptr(), ws.data_ptr(), wsf, s), "fused")
dbg.smplb200_debug_fz_timing(out)
nblk = (n + 63) // 64
total = 54 * nblk
names = {0: "A'prod wait aempty", 2: "T0 wait tempty", 3: "T0 wait afull", 6: "T1 wait tempty", 7: "T1 wait afull",
         10: "epi(s0) wait dfull", 11: "epi(s0) wait tfull", 12: "epi(s1) wait dfull", 13: "epi(s1) wait tfull",
         16: "D wait cfull", 17: "D wait dempty", 18: "D wait pace", 31: "kernel"}
import statistics
for k, nm in names.items():
    per_unit = []
    for b in range(148):
        units = total * (b + 1) // 148 - total * b // 148
        per_unit.append(out[b * 32 + k] / max(units, 1))
    print(f"{nm:22s} median {statistics.median(per_unit):9.0f} clk/unit   max {max(per_unit):9.0f}")
