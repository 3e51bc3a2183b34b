"""ResNet-34-D style encoder with mask-id embedding front-end and five shortcut branches, plus ASPP.

Reference: encoder/resnet.py:7-39 (BasicBlock), :42-153 (ResNet_D), :155-200 (ResShortCut_D),
:202-229 (ResMaskEmbedShortCut_D), module/aspp.py:4-57.  Attribute paths equal the reference's so that the
state dict is interchangeable.  Tensors flow as fp16 channels-last (NHWC in memory).
"""
import os

import torch
from torch import nn

from .. import dense, ops
from .layers import PlainConv, Slot, SNConv, seq

SIDE_SHORTCUTS = os.environ.get("MAGGIE_B200_NO_SIDE_SHORTCUTS", "0") != "1"


class EncBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1):
        super().__init__()
        self.stride = stride
        self.conv1 = SNConv(inplanes, planes, 3)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = SNConv(planes, planes, 3)
        self.bn2 = nn.BatchNorm2d(planes)
        nn.init.constant_(self.bn2.weight, 0)  # resnet.py:97-99
        self.downsample = None
        if stride != 1:
            # [AvgPool2d(2, stride), SN conv1x1, BN]  (resnet.py:111-116)
            self.downsample = seq(Slot(), SNConv(inplanes, planes, 1), nn.BatchNorm2d(planes))
            self.downsample[1].fold = True  # weight() yields the equivalent 2x2 stride-2 kernel (W / 4 on every tap)

    def forward(self, x):
        t = self.training
        idt, join = x, None
        if self.downsample is not None:
            # AvgPool2d(2,2) followed by a 1x1 conv == one 2x2 stride-2 conv with the 1x1 weight / 4 on every tap; the skip
            # path runs beside conv1 on a side stream
            idt, join = dense.side_branch(x, self.bn1, 5, lambda: ops.conv_bn_act(
                x, self.downsample[1].weight(), self.downsample[2], t, stride=2, padding=0, act=None))
        out = ops.conv_bn_act(x, self.conv1.weight(), self.bn1, t, stride=self.stride, act="relu")
        if join is not None:
            join()
        # conv2 -> bn2 -> (+identity) -> relu
        return ops.conv_bn_act(out, self.conv2.weight(), self.bn2, t, act="relu", residual=idt)


def _make_layer(inplanes, planes, blocks, stride):
    layers = [EncBlock(inplanes, planes, stride)]
    layers += [EncBlock(planes, planes) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


def _make_shortcut(inplane, planes):
    # conv -> ReLU -> BN -> conv -> ReLU -> BN  (note: activation BEFORE the norm, resnet.py:167-175)
    return seq(SNConv(inplane, planes, 3), Slot(), nn.BatchNorm2d(planes),
               SNConv(planes, planes, 3), Slot(), nn.BatchNorm2d(planes))


class ResMaskEmbedShortCutEncoder(nn.Module):
    """`res_shortcut_embed_29`: blocks [3,4,4,2], num_embed mask-embedding channels."""

    # packed input channels: 3 image + num_embed + zero padding.  32 (not 16): the two convolutions that read the packed
    # input then qualify for the halo-resident kernel K2b (forward and data gradient), which more than pays for the wider
    # tensor.
    IN_PAD = 32

    def __init__(self, num_mask=10, num_embed=3, **_):
        super().__init__()
        if num_mask != 10 or num_embed != 3:
            # K1 (mask-id embedding kernel) is built for the [11, 3] table of both live reference configs
            raise NotImplementedError(f"maggie_b200 supports num_mask=10, num_embed=3 (got {num_mask}, {num_embed})")
        self.num_embed = num_embed
        cin = 3 + num_embed
        self.conv1 = SNConv(cin, 32, 3)
        self.conv2 = SNConv(32, 32, 3)
        self.conv3 = SNConv(32, 64, 3)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm2d(32), nn.BatchNorm2d(32), nn.BatchNorm2d(64)
        self.layer1 = _make_layer(64, 64, 3, 1)
        self.layer2 = _make_layer(64, 128, 4, 2)
        self.layer3 = _make_layer(128, 256, 4, 2)
        self.layer_bottleneck = _make_layer(256, 512, 2, 2)
        with torch.no_grad():
            self.conv1.module.weight_bar[:, 3:] = 0  # resnet.py:102
        self.shortcut = nn.ModuleList([_make_shortcut(i, o) for i, o in
                                       ((cin, 32), (32, 32), (64, 64), (128, 128), (256, 256))])
        self.mask_embed_layer = nn.Embedding(num_mask + 1, num_embed)
        self.conv1.ci_pad_min = self.shortcut[0][0].ci_pad_min = self.IN_PAD   # the two layers that read the packed input

    def _shortcut(self, i, x):
        sc, t = self.shortcut[i], self.training
        x = ops.conv_bn_act(x, sc[0].weight(), sc[2], t, act="relu", act_first=True)
        return ops.conv_bn_act(x, sc[3].weight(), sc[5], t, act="relu", act_first=True)

    def forward(self, image, masks, slot_ids):
        """image [B,3,H,W] fp32; masks [B,M,H,W] {0,1} fp32 with slot_ids[M] (slot of each mask).
        Returns (os32 feature, (fea1..fea5)).

        The five shortcut branches are leaves of the trunk (nothing reads them before the decoder): on the GPU they run on
        a side stream, forked where their input appears and joined once at the end, so that they fill the SMs the
        latency-bound trunk layers leave idle.  Autograd replays each op's backward on its forward stream, so the same
        overlap happens in the backward pass; fork and join are plain event waits (CUDA-graph capturable)."""
        t = self.training
        hp = getattr(self, "precision", "fp16") == "high" and not t    # fp32-accurate evaluation (MaGGIe.set_precision)
        x = ops.mask_embed(image, masks, self.mask_embed_layer.weight, slot_ids, self.IN_PAD,
                           dtype=torch.float32 if hp else torch.float16)
        side = None
        # (with exchanged BatchNorm statistics all exchanges must stay in ONE stream order on every rank)
        if x.is_cuda and SIDE_SHORTCUTS and dense.sync_group(self.bn1) is None:
            side = dense.aux_stream(x.device, 1)
        main = torch.cuda.current_stream(x.device) if side is not None else None
        fea = [None] * 5

        def branch(i, f):
            if side is None:
                fea[i] = self._shortcut(i, f)
                return
            side.wait_stream(main)
            with torch.cuda.stream(side):
                fea[i] = self._shortcut(i, f)
            f.record_stream(side)

        branch(0, x)
        out = ops.conv_bn_act(x, self.conv1.weight(), self.bn1, t, stride=2)
        x1 = ops.conv_bn_act(out, self.conv2.weight(), self.bn2, t)
        branch(1, x1)
        out = ops.conv_bn_act(x1, self.conv3.weight(), self.bn3, t, stride=2)
        x2 = self.layer1(out)
        branch(2, x2)
        x3 = self.layer2(x2)
        branch(3, x3)
        x4 = self.layer3(x3)
        branch(4, x4)
        out = self.layer_bottleneck(x4)
        if side is not None:
            main.wait_stream(side)
            for f in fea:
                f.record_stream(main)
        return out, tuple(fea)


class ASPP(nn.Module):
    def __init__(self, in_channel=512, out_channel=512):
        super().__init__()
        mid = 256
        self.aspp1 = PlainConv(in_channel, mid, 1)
        self.aspp2, self.aspp3, self.aspp4 = (PlainConv(in_channel, mid, 3) for _ in range(3))
        self.aspp5 = PlainConv(in_channel, mid, 1)
        for i in range(1, 6):
            setattr(self, f"aspp{i}_bn", nn.BatchNorm2d(mid))
        self.conv2 = PlainConv(mid * 5, out_channel, 1)
        self.bn2 = nn.BatchNorm2d(out_channel)

    def forward(self, x):
        t = self.training
        # the three dilated branches are 32-CTA launches each: they run side by side on their own streams (see the encoder)
        par = x.is_cuda and SIDE_SHORTCUTS and dense.sync_group(self.aspp1_bn) is None
        main = torch.cuda.current_stream(x.device) if par else None
        ys = [ops.conv_bn_act(x, self.aspp1.w(), self.aspp1_bn, t, padding=0)]
        for i, d in ((2, 2), (3, 4), (4, 8)):
            if not par:
                ys.append(ops.conv_bn_act(x, getattr(self, f"aspp{i}").w(), getattr(self, f"aspp{i}_bn"), t,
                                          padding=d, dilation=d))
                continue
            side = dense.aux_stream(x.device, i)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ys.append(ops.conv_bn_act(x, getattr(self, f"aspp{i}").w(), getattr(self, f"aspp{i}_bn"), t,
                                          padding=d, dilation=d))
            x.record_stream(side)
        g = x.float().mean((2, 3), keepdim=True).to(x.dtype)
        g = ops.conv_bn_act(g, self.aspp5.w(), self.aspp5_bn, t, padding=0)
        ys.append(g.expand(-1, -1, x.shape[2], x.shape[3]))
        if par:
            for i in (2, 3, 4):
                main.wait_stream(dense.aux_stream(x.device, i))
                ys[i - 1].record_stream(main)
        y = torch.cat(ys, 1).contiguous(memory_format=torch.channels_last)
        return ops.conv_bn_act(y, self.conv2.w(), self.bn2, t, padding=0)
