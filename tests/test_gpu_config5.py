"""-m gpu: BASELINE.json configs[4] end to end -- the reference's random-init DLA-34 producer
(oracle/dla34_ref.py, pinned to the unmodified `dla_net` by tests/test_dla_oracle.py) at 512 x 512,
batch 32 -> sigmoid(hm) -> `decode_gather` (K = 32 people per image) -> SMPL for the 1024 people.

What is asserted on the REAL head maps the network produces:
  * decode: scores / inds / clses / ys / xs and the gathered pose|shape|cam vectors equal, bit for bit, what
    the reference's `_nms` + `_topk` + `_transpose_and_gather_feat` give (oracle/decode_ref.py, itself pinned
    bit-exact to the unmodified reference functions) on the same head maps -- with BatchNorm statistics
    calibrated (oracle.dla34_ref.calibrate_batchnorm) so the heat map has distinct peaks; on the RAW
    random-init network, whose heat map collapses to its bias and ties by the thousand, the result is checked
    to be the reference's up to the order of equal scores (oracle.decode_ref.check_equivalent);
  * meshes: vertices / joints / kp2d of the 1024 decoded people match the CPU oracle (fp32 tolerance for
    joints and kp2d, 1e-5 m for the split-bf16 blendshapes);
  * USE_DCN=True variant (reference model.py:346-362): every one of the 16 deformable layers of the neck,
    run through the product's `DCN` module inside the network, matches oracle/dcn_ref.dcn_forward on the
    very tensors that layer saw.
"""
import pytest
import torch

from human_3d_reconstruction_b200 import DCN, SMPL, decode_gather, synthetic
from oracle.dcn_ref import dcn_forward
from oracle.decode_ref import check_equivalent, decode_gather as decode_ref
from oracle.dla34_ref import HEADS_HMR, calibrate_batchnorm, dla_net
from oracle.smpl_ref import smpl_forward_chunked

pytestmark = pytest.mark.gpu
SEED = 317                      # reference opts.py:37
DCN_RTOL = 5e-5                 # of the layer's largest |output| (split-bf16 operands, fp32 accumulate)


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


def images(batch, size, seed=11):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, size, size, generator=g)


def test_dla34_decode_smpl_batch32_512(dev):
    B, K = 32, 32
    x = images(B, 512).to(dev)
    net = dla_net(dict(HEADS_HMR), seed=SEED).eval().to(dev)

    def run():
        with torch.no_grad():
            out = net(x)[0]
            hm = torch.sigmoid(out["hm"])
            heads = [out["pose"], out["shape"], out["cam"]]
            assert hm.shape == (B, 1, 128, 128) and heads[0].shape == (B, 72, 128, 128)
            return hm, heads, decode_gather(hm, heads, K)

    def cpu(got):
        return tuple(t.cpu() for t in got[:5]) + ([t.cpu() for t in got[5]],)

    # (1) the raw random-init network, eval mode: a near-constant heat map full of exact ties
    hm, heads, got = run()
    ok, ties = check_equivalent(cpu(got), hm.cpu(), [h.cpu() for h in heads], K)
    assert ok, "decode on the raw random-init head maps is not a valid reference result"
    print(f"raw random-init network: {ties}/{B} images have tied scores among their top {K + 1}")
    # (2) BatchNorm statistics calibrated: distinct peaks => index-for-index equality with the reference
    calibrate_batchnorm(net, x)
    hm, heads, got = run()
    ref = decode_ref(hm.cpu(), [h.cpu() for h in heads], K)
    ok, ties = check_equivalent(cpu(got), hm.cpu(), [h.cpu() for h in heads], K)
    assert ok and ties == 0, f"calibrated network: ok={ok}, images with ties={ties}"
    for name, a, b in zip(("scores", "inds", "clses", "ys", "xs"), got[:5], ref[:5]):
        assert torch.equal(a.cpu(), b), name
    for name, a, b in zip(("pose", "shape", "cam"), got[5], ref[5]):
        assert torch.equal(a.cpu(), b), name
    # the decoded people through the SMPL kernels
    model = synthetic.make_model(0)
    layer = SMPL(model).to(dev)                                  # precision/lbs 'auto': 1024 bodies -> tcgen05 paths
    pose, shape, cam = (t.reshape(B * K, -1) for t in got[5])
    with torch.no_grad():
        v, j, k = layer(shape, pose, cam)
    rv, rj, rk = smpl_forward_chunked(model, shape.cpu().numpy(), pose.cpu().numpy(), cam.cpu().numpy(), chunk=256)
    assert torch.allclose(v.cpu(), rv, rtol=1e-5, atol=1e-5), (v.cpu() - rv).abs().max()
    assert torch.allclose(j.cpu(), rj, rtol=1e-5, atol=1e-6), (j.cpu() - rj).abs().max()
    assert torch.allclose(k.cpu(), rk, rtol=1e-5, atol=2e-6), (k.cpu() - rk).abs().max()


def test_dla34_neck_with_deformable_convolutions(dev):
    def deform(ci, co):
        return DCN(ci, co, kernel_size=(3, 3), stride=1, padding=1, dilation=1, deformable_groups=1)

    net = dla_net(dict(HEADS_HMR), seed=SEED, deform=deform).eval()
    layers = [m for m in net.modules() if isinstance(m, DCN)]
    assert len(layers) == 16
    g = torch.Generator().manual_seed(SEED)
    with torch.no_grad():                                       # upstream zero-initialises the offset conv: make it deform
        for m in layers:
            m.conv_offset_mask.weight.normal_(0.0, 0.6 / (m.in_channels * 9) ** 0.5, generator=g)
            m.conv_offset_mask.bias.normal_(0.0, 0.5, generator=g)
    net = net.to(dev)
    x = images(1, 256, seed=12).to(dev)
    calibrate_batchnorm(net, x)                                # O(1) activations at every layer (see its docstring)
    # the layer's own offset convolution must be plain fp32 for the comparison below (cuDNN would use TF32
    # by default, moving every sampling position by ~1e-3 px relative to the float64 oracle)
    tf32_was, torch.backends.cudnn.allow_tf32 = torch.backends.cudnn.allow_tf32, False
    seen = []
    hooks = [m.register_forward_hook(lambda mod, inp, outp: seen.append((mod, inp[0].detach().cpu(), outp.detach().cpu())))
             for m in layers]
    with torch.no_grad():
        out = net(x)[0]
        hm = torch.sigmoid(out["hm"])
        got = decode_gather(hm, [out["pose"], out["shape"], out["cam"]], 32)
    for h in hooks:
        h.remove()
    torch.backends.cudnn.allow_tf32 = tf32_was
    assert len(seen) == 16 and all(torch.isfinite(t).all() for t in got[5])
    worst = 0.0
    for mod, x, y in seen:
        ref = dcn_forward(x, mod.conv_offset_mask.weight.cpu(), mod.conv_offset_mask.bias.cpu(), mod.weight.cpu(),
                          mod.bias.cpu(), dtype=torch.float64)
        scale = ref.abs().max().item()
        err = (y.double() - ref).abs().max().item()
        worst = max(worst, err / scale)
        assert err <= DCN_RTOL * scale, f"DCN {mod.in_channels}->{mod.out_channels} @{tuple(x.shape[2:])}: {err:.3e} vs {scale:.3e}"
    print(f"worst DCN layer error / scale: {worst:.2e}")
