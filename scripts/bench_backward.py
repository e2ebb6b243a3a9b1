"""Times smplb200_backward (direct C-ABI calls, CUDA events) next to smplb200_forward.

usage: python scripts/bench_backward.py [--n 32 128 1024 4096] [--iters 20]
Prints one JSON line per batch size: forward us, backward us with the vertex path
(g_vertices + g_joints + g_kp2d) and backward us for a joints/kp2d-only loss.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_3d_reconstruction_b200 import SMPL, capi, synthetic  # noqa: E402


def timed(fn, iters, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs="+", default=[32, 128, 1024, 4096])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--once", action="store_true", help="one backward per batch size, no timing (ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    layer = SMPL.synthetic(0, precision=a.precision).to(dev)
    h = layer.handle(dev)
    lib = capi.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream(dev).cuda_stream
    for n in a.n:
        b, p, c = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 1))
        v = torch.empty(n, 6890, 3, device=dev); j = torch.empty(n, 24, 3, device=dev); k = torch.empty(n, 24, 2, device=dev)
        gv, gj, gk = torch.randn_like(v), torch.randn_like(j), torch.randn_like(k)
        gb, gp, gc = torch.empty_like(b), torch.empty_like(p), torch.empty_like(c)
        wf = torch.empty(h.workspace_bytes(n, layer.flags), dtype=torch.uint8, device=dev)
        wb = torch.empty(lib.smplb200_backward_workspace_bytes(h.ptr, n, layer.flags, 1), dtype=torch.uint8, device=dev)

        def fwd():
            capi.check(lib.smplb200_forward(h.ptr, b.data_ptr(), p.data_ptr(), c.data_ptr(), n, v.data_ptr(),
                                            j.data_ptr(), k.data_ptr(), wf.data_ptr(), wf.numel(), layer.flags, s), "fwd")

        def bwd(vertex, reuse=False):
            capi.check(lib.smplb200_backward(h.ptr, b.data_ptr(), p.data_ptr(), c.data_ptr(), n, j.data_ptr(),
                                             gv.data_ptr() if vertex else None, gj.data_ptr(), gk.data_ptr(),
                                             gb.data_ptr(), gp.data_ptr(), gc.data_ptr(),
                                             wf.data_ptr() if reuse else None, wf.numel() if reuse else 0,
                                             wb.data_ptr(), wb.numel(), layer.flags, s), "bwd")
        if a.once:
            fwd(); bwd(True); bwd(True, True); bwd(False); torch.cuda.synchronize()
            continue
        print(json.dumps({"n": n, "forward_us": round(timed(fwd, a.iters, flush), 1),
                          "backward_vertex_path_us": round(timed(lambda: bwd(True), a.iters, flush), 1),
                          "backward_vertex_path_reusing_forward_ws_us": round(timed(lambda: bwd(True, True), a.iters, flush), 1),
                          "backward_joints_only_us": round(timed(lambda: bwd(False), a.iters, flush), 1),
                          "backward_workspace_MB": round(wb.numel() / 2**20, 1)}), flush=True)


if __name__ == "__main__":
    main()
