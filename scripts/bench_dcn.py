"""Times the fused DCNv2 forward on the 16 DeformConv layers of the reference DLA-34 network
(shapes recorded from reference src/lib/models/model.py at 512x512 input: 7 distinct layers), batch 32
(BASELINE.json configs[4]), next to torchvision.ops.deform_conv2d on the same GPU (library kernel:
materialised columns + cuBLAS) and, on a bounded sample, the CPU oracle port.

usage: python scripts/bench_dcn.py [--batch 32] [--iters 20]   -> one JSON line per layer + a total
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from human_3d_reconstruction_b200 import dcn_v2_conv  # noqa: E402

LAYERS = [  # (count in the network, Ci, Co, H, W)
    (1, 512, 256, 16, 16), (1, 256, 256, 32, 32), (2, 256, 128, 32, 32), (2, 128, 128, 64, 64),
    (4, 128, 64, 64, 64), (5, 64, 64, 128, 128), (1, 256, 64, 32, 32)]


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--once", action="store_true", help="one launch per layer, no timing (for ncu)")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        from torchvision.ops import deform_conv2d
    except Exception:
        deform_conv2d = None
    g = torch.Generator(device="cpu").manual_seed(317)
    tot_ours = tot_tv = 0.0
    for cnt, Ci, Co, H, W in LAYERS:
        B = a.batch
        x = torch.randn(B, Ci, H, W, generator=g).to(dev)
        w = (torch.randn(Co, Ci, 3, 3, generator=g) / (Ci * 9) ** 0.5).to(dev)
        b = torch.randn(Co, generator=g).to(dev)
        off = (torch.randn(B, 18, H, W, generator=g) * 2.0).to(dev)
        m = torch.rand(B, 9, H, W, generator=g).to(dev)
        with torch.no_grad():
            if a.once:
                dcn_v2_conv(x, off, m, w, b)
                torch.cuda.synchronize()
                continue
            us = timed(lambda: dcn_v2_conv(x, off, m, w, b), a.iters)
            xcl = x.contiguous(memory_format=torch.channels_last)
            us_cl = timed(lambda: dcn_v2_conv(xcl, off, m, w, b), a.iters)
            row = {"layer": f"{Ci}->{Co} @ {H}x{W}", "count_in_network": cnt, "batch": B, "ours_us": round(us, 1),
                   "ours_channels_last_input_us": round(us_cl, 1)}
            tot_cl = globals().setdefault("_tot_cl", [0.0]); tot_cl[0] += cnt * us_cl
            flops = 2.0 * B * H * W * Co * Ci * 9
            bytes_ = 4.0 * (B * Ci * H * W + 27 * B * H * W + B * Co * H * W)
            row["tflops_fp32_equivalent"] = round(flops / us * 1e-6, 1)
            row["algorithmic_GBps"] = round(bytes_ / us * 1e-3, 1)
            if deform_conv2d is not None:
                tv = timed(lambda: deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m), max(3, a.iters // 4))
                ref = deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m)
                got = dcn_v2_conv(x, off, m, w, b)
                row["torchvision_cuda_us"] = round(tv, 1)
                row["max_abs_diff_vs_torchvision_fp32"] = float((got - ref).abs().max())
                row["max_abs_output"] = float(ref.abs().max())
                tot_tv += cnt * tv
            tot_ours += cnt * us
            print(json.dumps(row), flush=True)
    if a.once:
        return
    out = {"network_total_16_layers_us": round(tot_ours, 1),
           "network_total_channels_last_input_us": round(globals().get("_tot_cl", [0.0])[0], 1), "torchvision_cuda_total_us": round(tot_tv, 1) if tot_tv else None}
    # CPU port on a bounded sample: one image of the 64->64 @128x128 layer
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle.dcn_ref import dcn_v2_forward
    x = torch.randn(1, 64, 128, 128, generator=g); w = torch.randn(64, 64, 3, 3, generator=g) / 24.0
    off = torch.randn(1, 18, 128, 128, generator=g) * 2; m = torch.rand(1, 9, 128, 128, generator=g)
    t0 = time.perf_counter()
    dcn_v2_forward(x, w, torch.zeros(64), off, m)
    out["cpu_oracle_port_one_image_64to64_128x128_ms"] = round((time.perf_counter() - t0) * 1e3, 1)
    out["cpu_threads"] = torch.get_num_threads()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
