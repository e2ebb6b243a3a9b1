"""Host-side multi-GPU logic on CPU: shard bounds and the optional all-gather of joints, run
with the gloo backend at world_size 2."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from human_3d_reconstruction_b200 import sharding


@pytest.mark.parametrize("n,world", [(0, 2), (1, 2), (7, 2), (8, 2), (65536, 8), (10, 4), (3, 8)])
def test_shard_bounds_partition(n, world):
    spans = [sharding.shard_bounds(n, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    assert max(b - a for a, b in spans) <= sharding.shard_size(n, world)


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        betas = torch.randn(n, 10, generator=g)
        pose = torch.randn(n, 72, generator=g)
        cam = torch.randn(n, 3, generator=g)

        def fake_forward(b, p, c):  # stands in for the CUDA layer: per-body independent maps
            verts = (b.sum(1, keepdim=True) + p[:, :3]).unsqueeze(1).expand(-1, 5, -1).contiguous()
            joints = p.view(-1, 24, 3) * 2.0
            kp = c[:, None, 0:1] * (joints[:, :, :2] + c[:, None, 1:3])
            return verts, joints, kp

        sh = sharding.ShardedSMPL(fake_forward)
        verts, joints, kp2d, (lo, hi) = sh.forward(betas, pose, cam, gather=True)
        ref_v, ref_j, ref_k = fake_forward(betas, pose, cam)
        assert (lo, hi) == sharding.shard_bounds(n, world, rank)
        assert torch.equal(verts, ref_v[lo:hi])          # vertices stay sharded
        assert torch.equal(joints, ref_j) and torch.equal(kp2d, ref_k)  # small outputs gathered
        v2, j2, k2, _ = sh.forward(betas, pose, cam, gather=False)
        assert torch.equal(j2, ref_j[lo:hi]) and torch.equal(k2, ref_k[lo:hi])
        # the exchange object (collective transport on CPU; peer stores on the GPU box), two steps so both
        # slots are used, with different data per step
        ex = sharding.PeerExchange(n, "cpu")
        assert ex.transport == "collective" and ex.world == world
        sh2 = sharding.ShardedSMPL(fake_forward, exchange=ex)
        for scale in (1.0, -3.0):
            v3, j3, k3, _ = sh2.forward(betas, pose * scale, cam, gather=True)
            rv, rj, rk = fake_forward(betas, pose * scale, cam)
            assert torch.equal(v3, rv[lo:hi]) and torch.equal(j3, rj) and torch.equal(k3, rk)
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 8, 1])
def test_sharded_forward_and_gather_gloo_world2(tmp_path, n):
    port = 29500 + (os.getpid() % 500) + n
    mp.spawn(_worker, args=(2, port, n, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))


def _gpu_worker(rank, world, port, n, tmp, transport="auto"):
    """Two processes, one GPU each, NCCL for rendezvous only: the joints | kp2d rows travel by peer stores."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from human_3d_reconstruction_b200 import SMPL, synthetic
        layer = SMPL.synthetic(0).to(dev)
        ex = sharding.PeerExchange(n, dev, transport=transport)
        sh = sharding.ShardedSMPL(layer, exchange=ex)
        with torch.no_grad():
            for step in range(5):                                   # > 2 steps: both slots are reused
                betas, pose, cam = (torch.from_numpy(x).to(dev) for x in synthetic.make_inputs(n, 50 + step))
                verts, joints, kp2d, (lo, hi) = sh.forward(betas, pose, cam, gather=True)
                ref = layer(betas, pose, cam)                       # the whole batch on this rank
                torch.cuda.synchronize()
                assert torch.equal(verts, ref[0][lo:hi])
                assert torch.equal(joints, ref[1]) and torch.equal(kp2d, ref[2]), f"step {step}"
        with open(os.path.join(tmp, f"ok{rank}"), "w") as f:
            f.write(ex.transport + ("" if ex.why_not_peer is None else " (" + ex.why_not_peer + ")"))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two visible GPUs")
@pytest.mark.parametrize("transport", ["auto", "dma"])
@pytest.mark.parametrize("n", [4096, 777])
def test_peer_store_exchange_two_gpus(tmp_path, n, transport):
    """'auto' = peer stores from a kernel; 'dma' = copy-engine copies + stream memory operations.

    The 'dma' cases passed on a 2-GPU box, but that transport's stream-memory-operation wait has no time bound and
    its one 8-GPU run did not finish: it is opt-in in the product and opt-in here
    (SMPLB200_TEST_DMA=1), so that an unattended suite can never sit in an unbounded wait.
    """
    if transport == "dma" and os.environ.get("SMPLB200_TEST_DMA") != "1":
        pytest.skip("opt-in transport: set SMPLB200_TEST_DMA=1")
    port = 29700 + (os.getpid() % 200) + (n % 7) + (13 if transport == "dma" else 0)
    mp.spawn(_gpu_worker, args=(2, port, n, str(tmp_path), transport), nprocs=2, join=True)
    got = [(tmp_path / f"ok{r}").read_text() for r in range(2)]
    print("exchange transport:", got)
    assert all(g.startswith(("peer", "dma", "collective")) for g in got)
    if transport == "dma":
        assert all(g.startswith("dma") for g in got)
