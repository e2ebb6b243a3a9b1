"""TEST INFRASTRUCTURE -- restatement of the reference's DLA-34 CenterNet-style producer network.

BASELINE.json configs[4] ("random-init DLA-34 CenterNet-style backbone at 512x512, batch 32, top-K
32 people per image decoded through the new SMPL kernels") needs the network that PRODUCES the head
maps the hot path consumes.  `/root/reference` does not travel to the GPU box, so the network is
restated here (plain PyTorch modules; cuDNN does the work on the GPU) and pinned to the reference:

  * follows reference src/lib/models/model.py:
      BasicBlock :31-60, Root :148-166, Tree :169-226, DLA (dla34) :229-318, DeformConv :346-362,
      IDAUp :365-390, DLAUp :393-416, DLASeg + heads :430-499, dla_net :501-516;
  * attribute names equal the reference's, so `state_dict()` keys match and reference weights load
    with strict=True; modules are constructed in the reference's order, so under the same
    `torch.manual_seed` the random initialisation consumes the RNG identically and the weights are
    EQUAL -- tests/golden/make_dla_golden.py (run in the build container, where the reference is
    importable) checks both and commits a small input/output fixture plus parameter checksums that
    tests/test_dla_oracle.py re-checks anywhere.

Only tests/, __graft_entry__.smoke() and bench.py import this file.  The product never does: its
part of configs[4] starts at the head maps (decode_gather -> SMPL, and the DCN module when the neck
runs with deformable convolutions).

`deform=None` builds the `not_use_dcn=True` variant (3x3 nn.Conv2d in every DeformConv);
`deform=callable(chi, cho)` supplies the deformable module (upstream's USE_DCN=True).
"""
from __future__ import annotations

import math

import torch
from torch import nn

BN_MOMENTUM = 0.1
HEADS_HMR = {"hm": 1, "wh": 2, "reg": 2, "pose": 72, "shape": 10, "cam": 3}   # reference opts.py:248-258 + SMPL heads


def _bn(c):
    return nn.BatchNorm2d(c, momentum=BN_MOMENTUM)


class BasicBlock(nn.Module):
    def __init__(self, cin, cout, stride=1, dilation=1):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, dilation, dilation, bias=False)
        self.bn1 = _bn(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, dilation, dilation, bias=False)
        self.bn2 = _bn(cout)

    def forward(self, x, residual=None):
        residual = x if residual is None else residual
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        y += residual
        return self.relu(y)


class Root(nn.Module):
    def __init__(self, cin, cout, kernel_size, residual):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, 1, 1, (kernel_size - 1) // 2, bias=False)
        self.bn = _bn(cout)
        self.relu = nn.ReLU(inplace=True)
        self.residual = residual

    def forward(self, *xs):
        y = self.bn(self.conv(torch.cat(xs, 1)))
        if self.residual:
            y += xs[0]
        return self.relu(y)


class Tree(nn.Module):
    def __init__(self, levels, cin, cout, stride=1, level_root=False, root_dim=0):
        super().__init__()
        root_dim = root_dim or 2 * cout
        if level_root:
            root_dim += cin
        if levels == 1:
            self.tree1 = BasicBlock(cin, cout, stride)
            self.tree2 = BasicBlock(cout, cout, 1)
            self.root = Root(root_dim, cout, 1, False)
        else:
            self.tree1 = Tree(levels - 1, cin, cout, stride)
            self.tree2 = Tree(levels - 1, cout, cout, root_dim=root_dim + cout)
        self.levels, self.level_root = levels, level_root
        self.downsample = nn.MaxPool2d(stride, stride=stride) if stride > 1 else None
        self.project = None
        if cin != cout:
            self.project = nn.Sequential(nn.Conv2d(cin, cout, 1, 1, bias=False), _bn(cout))

    def forward(self, x, residual=None, children=None):
        children = [] if children is None else children
        bottom = self.downsample(x) if self.downsample else x
        residual = self.project(bottom) if self.project else bottom
        if self.level_root:
            children.append(bottom)
        x1 = self.tree1(x, residual)
        if self.levels == 1:
            return self.root(self.tree2(x1), x1, *children)
        children.append(x1)
        return self.tree2(x1, children=children)


def _conv_level(cin, cout, stride):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride, 1, bias=False), _bn(cout), nn.ReLU(inplace=True))


class DLA34(nn.Module):
    channels = [16, 32, 64, 128, 256, 512]

    def __init__(self):
        super().__init__()
        c = self.channels
        self.base_layer = nn.Sequential(nn.Conv2d(3, c[0], 7, 1, 3, bias=False), _bn(c[0]), nn.ReLU(inplace=True))
        self.level0 = _conv_level(c[0], c[0], 1)
        self.level1 = _conv_level(c[0], c[1], 2)
        self.level2 = Tree(1, c[1], c[2], 2, level_root=False)
        self.level3 = Tree(2, c[2], c[3], 2, level_root=True)
        self.level4 = Tree(2, c[3], c[4], 2, level_root=True)
        self.level5 = Tree(1, c[4], c[5], 2, level_root=True)

    def forward(self, x):
        outs = []
        x = self.base_layer(x)
        for i in range(6):
            x = getattr(self, f"level{i}")(x)
            outs.append(x)
        return outs


def _bilinear_upsample_weights(up):
    """reference fill_up_weights (model.py:332-343): a fixed bilinear kernel in every channel."""
    w = up.weight.data
    f = math.ceil(w.size(2) / 2)
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    for i in range(w.size(2)):
        for j in range(w.size(3)):
            w[0, 0, i, j] = (1 - math.fabs(i / f - c)) * (1 - math.fabs(j / f - c))
    for ch in range(1, w.size(0)):
        w[ch, 0, :, :] = w[0, 0, :, :]


class DeformConv(nn.Module):
    def __init__(self, cin, cout, deform=None):
        super().__init__()
        self.actf = nn.Sequential(_bn(cout), nn.ReLU(inplace=True))
        self.conv = deform(cin, cout) if deform is not None else nn.Conv2d(cin, cout, 3, 1, 1)

    def forward(self, x):
        return self.actf(self.conv(x))


class IDAUp(nn.Module):
    def __init__(self, o, channels, up_f, deform=None):
        super().__init__()
        for i in range(1, len(channels)):
            f = int(up_f[i])
            proj = DeformConv(channels[i], o, deform)
            node = DeformConv(o, o, deform)
            up = nn.ConvTranspose2d(o, o, f * 2, stride=f, padding=f // 2, output_padding=0, groups=o, bias=False)
            _bilinear_upsample_weights(up)
            setattr(self, f"proj_{i}", proj)
            setattr(self, f"up_{i}", up)
            setattr(self, f"node_{i}", node)

    def forward(self, layers, startp, endp):
        for i in range(startp + 1, endp):
            k = i - startp
            layers[i] = getattr(self, f"up_{k}")(getattr(self, f"proj_{k}")(layers[i]))
            layers[i] = getattr(self, f"node_{k}")(layers[i] + layers[i - 1])


class DLAUp(nn.Module):
    def __init__(self, startp, channels, scales, deform=None):
        super().__init__()
        self.startp = startp
        channels, cin, scales = list(channels), list(channels), list(scales)
        for i in range(len(channels) - 1):
            j = -i - 2
            setattr(self, f"ida_{i}", IDAUp(channels[j], cin[j:], [s // scales[j] for s in scales[j:]], deform))
            scales[j + 1:] = [scales[j]] * len(scales[j + 1:])
            cin[j + 1:] = [channels[j]] * len(cin[j + 1:])

    def forward(self, layers):
        out = [layers[-1]]
        for i in range(len(layers) - self.startp - 1):
            getattr(self, f"ida_{i}")(layers, len(layers) - i - 2, len(layers))
            out.insert(0, layers[-1])
        return out


class DLASeg(nn.Module):
    """`dla_net(heads, num_layers=34, head_conv=256, down_ratio=4)` of the reference."""

    def __init__(self, heads=None, head_conv=256, down_ratio=4, last_level=5, deform=None):
        super().__init__()
        heads = dict(HEADS_HMR if heads is None else heads)
        self.first_level = int(math.log2(down_ratio))
        self.last_level = last_level
        self.base = DLA34()
        ch = self.base.channels
        scales = [2 ** i for i in range(len(ch[self.first_level:]))]
        self.dla_up = DLAUp(self.first_level, ch[self.first_level:], scales, deform)
        self.ida_up = IDAUp(ch[self.first_level], ch[self.first_level:self.last_level],
                            [2 ** i for i in range(self.last_level - self.first_level)], deform)
        self.heads = heads
        for head, classes in heads.items():
            fc = nn.Sequential(nn.Conv2d(ch[self.first_level], head_conv, 3, padding=1, bias=True),
                               nn.ReLU(inplace=True),
                               nn.Conv2d(head_conv, classes, 1, 1, 0, bias=True))
            if "hm" in head:
                fc[-1].bias.data.fill_(-2.19)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d) and m.bias is not None:
                        nn.init.constant_(m.bias, 0)
            setattr(self, head, fc)

    def forward(self, x):
        x = self.dla_up(self.base(x))
        y = [x[i].clone() for i in range(self.last_level - self.first_level)]
        self.ida_up(y, 0, len(y))
        return [{head: getattr(self, head)(y[-1]) for head in self.heads}]


def dla_net(heads=None, head_conv=256, down_ratio=4, deform=None, seed=None):
    """Random-init DLA-34 producer.  `seed` (reference default 317, opts.py:37) seeds torch first."""
    if seed is not None:
        torch.manual_seed(seed)
    return DLASeg(heads, head_conv=head_conv, down_ratio=down_ratio, deform=deform)


def calibrate_batchnorm(net: nn.Module, images: torch.Tensor) -> nn.Module:
    """Set every BatchNorm's running statistics to the batch statistics of `images` (one train-mode pass
    with momentum 1), then return the network in eval mode.

    A random-init DLA-34 in eval mode with the default running statistics (mean 0, var 1) lets the signal
    die out over its ~40 layers: the centre heat map collapses to its bias (-2.19 +- 2e-3) and thousands
    of pixels tie in fp32.  With calibrated statistics the eval-mode network on `images` equals its
    train-mode self, the heads have O(1) dynamic range and the decode stage sees distinct peaks -- the
    regime a trained network is in.  Weights stay the seeded random initialisation.
    """
    bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
    old = [m.momentum for m in bns]
    for m in bns:
        m.momentum = 1.0
    net.train()
    with torch.no_grad():
        net(images)
    for m, mom in zip(bns, old):
        m.momentum = mom
    return net.eval()
