import sys, torch, time
sys.path.insert(0, "/root/repo")
from human_3d_reconstruction_b200 import decode_gather
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
heat = torch.sigmoid(torch.randn(32, 1, 128, 128, generator=g) * 2.0).to(dev)
heads = [torch.randn(32, ch, 128, 128, generator=g).to(dev) for ch in (72, 10, 3)]
for _ in range(5): decode_gather(heat, heads, 32)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100): decode_gather(heat, heads, 32)
e1.record(); torch.cuda.synchronize()
print("decode_gather us/call", e0.elapsed_time(e1) / 100 * 1e3)
# kernel time without the Python wrapper: CUDA-graph replay
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    decode_gather(heat, heads, 32)
torch.cuda.current_stream().wait_stream(side)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    out = decode_gather(heat, heads, 32)
for _ in range(5): gr.replay()
torch.cuda.synchronize()
e0.record()
for _ in range(200): gr.replay()
e1.record(); torch.cuda.synchronize()
print("decode_gather graph replay us", e0.elapsed_time(e1) / 200 * 1e3)
