"""MaGGIe decoder: OS32->OS8 dense blocks, mask-guided attention at OS8 (InstanceMatteDecoder), uncertainty
mask, sparse OS8->OS4->OS2->OS1 refinement on the active sites only, progressive fusion.

Reference: decoder/resnet.py:9-45 (BasicBlock), decoder/resnet_inst_matt_spconv.py:14-388,
module/instance_matte_decoder.py:9-306, module/mask_attention.py:9-206.  Attribute paths equal the
reference's (state-dict compatible).  No spconv, no cv2, no nonzero/boolean-index host syncs except the one
16-byte read of the site counts.
"""
import random

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import dense, ops, packs
from .layers import PlainConv, Slot, SNConv, SparseConvParams, seq


# ------------------------------------------------------------------------------------------- dense blocks
class DecBlock(nn.Module):
    """conv1 (ConvT 4x4 s2 when upsampling else 3x3) - BN - LReLU - conv2 3x3 - BN - (+skip) - LReLU."""

    def __init__(self, inplanes, planes, up=False):
        super().__init__()
        self.up = up
        self.conv1 = SNConv(inplanes, inplanes, 4 if up else 3, transposed=up)
        self.bn1 = nn.BatchNorm2d(inplanes)
        self.conv2 = SNConv(inplanes, planes, 3)
        self.bn2 = nn.BatchNorm2d(planes)
        self.upsample = None
        if up:
            # [UpsamplingNearest2d(2), SN conv1x1, BN]  (resnet_inst_matt_spconv.py:141-146)
            self.upsample = seq(Slot(), SNConv(inplanes, planes, 1), nn.BatchNorm2d(planes))

    def forward(self, x):
        t = self.training
        idt, join = x, None
        if self.up:
            # nearest x2 commutes with a 1x1 conv and leaves BN batch statistics unchanged, so the skip path runs
            # at the low resolution and is replicated afterwards (beside conv1, on a side stream)
            idt, join = dense.side_branch(x, self.bn1, 5, lambda: ops.conv_bn_act(
                x, self.upsample[1].weight(), self.upsample[2], t, padding=0, act=None))
        out = ops.conv_bn_act(x, self.conv1.weight(), self.bn1, t, act="lrelu", transposed=self.up)
        if join is not None:
            join()
        return ops.conv_bn_act(out, self.conv2.weight(), self.bn2, t, act="lrelu", residual=idt, res_up=self.up)


# ------------------------------------------------------------------------------------------- attention
class _FFN(nn.Module):
    def __init__(self, d, hidden, dropout=0.0):
        super().__init__()
        self.linear1, self.linear2 = nn.Linear(d, hidden), nn.Linear(hidden, d)
        self.dropout = nn.Dropout(dropout)  # p is read by the forward below
        self.norm = nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, x):
        p, t = self.dropout.p, self.training
        drop = (lambda h: F.dropout(h, p, t)) if (p > 0 and t) else (lambda h: h)
        if x.is_cuda and x.numel() // x.shape[-1] >= 1024:
            # site rows (inst_spec_layer on the OS8 sites): tcgen05 rows GEMM
            h = drop(F.relu(ops.linear_rows(x, self.linear1.weight, self.linear1.bias)))
            h = drop(ops.linear_rows(h, self.linear2.weight, self.linear2.bias))
        else:
            # instance tokens: fp32 library GEMMs on the master weights
            h = drop(ops.small_linear(x, self.linear1.weight, self.linear1.bias, relu=True))
            h = drop(ops.small_linear(h, self.linear2.weight, self.linear2.bias))
        return ops.layer_norm(x, self.norm, residual=h)


class _Attn(nn.Module):
    """Post-norm attention layer; `attn_name` is `multihead_attn` (cross) or `self_attn` (self)."""

    def __init__(self, d, attn_name):
        super().__init__()
        self.attn_name = attn_name
        setattr(self, attn_name, nn.MultiheadAttention(d, 1, dropout=0.0))
        self.norm = nn.LayerNorm(d)
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    def forward(self, tgt, mem, tgt_pos, mem_pos, key_padding=None, guidance=None):
        """tgt [B,L,E] (queries, residual stream), mem [B,S,E]; positional terms are added to Q and K only."""
        mha = getattr(self, self.attn_name)
        E = tgt.shape[-1]
        # q / k / v parts of the packed in_proj parameters as views; their gradients are re-assembled by ONE concatenation
        (wq, wk, wv), (bq, bk, bv) = ops.split_rows(mha.in_proj_weight, 3), ops.split_rows(mha.in_proj_bias, 3)
        q = ops.linear_rows(tgt, wq, bq, pos=tgt_pos)
        k = ops.linear_rows(mem, wk, bk, pos=mem_pos)
        v = ops.linear_rows(mem, wv, bv)
        o, stat = ops.attention(q, k, v, key_padding, guidance)
        o = ops.linear_rows(o, mha.out_proj.weight, mha.out_proj.bias)
        return ops.layer_norm(tgt, self.norm, residual=o), stat


class _MLP1(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.layers = nn.ModuleList([nn.Linear(din, dout)])


class InstanceMatteDecoder(nn.Module):
    def __init__(self, input_dim=128, attention_dim=128, n_block=2, output_dim=64, max_inst=10, use_id_pe=True, **_):
        super().__init__()
        d = attention_dim
        self.n_block, self.max_inst, self.use_id_pe = n_block, max_inst, use_id_pe
        self.feat_proj = _MLP1(input_dim, d)
        self.sa_layers = nn.ModuleList(_Attn(d, "self_attn") for _ in range(n_block))
        self.token_feat_ca_layers = nn.ModuleList(_Attn(d, "multihead_attn") for _ in range(n_block))
        self.mlp_layers = nn.ModuleList(_FFN(d, d) for _ in range(n_block))
        self.feat_token_ca_layers = nn.ModuleList(_Attn(d, "multihead_attn") for _ in range(n_block))
        self.final_token_feat_ca = _Attn(d, "multihead_attn")
        self.final_mlp = _MLP1(d, output_dim)
        self.decoder_norm = nn.LayerNorm(output_dim)
        self.query_feat = nn.Embedding(max_inst, d)
        self.id_embedding = nn.Embedding(max_inst + 1, d)
        nn.init.xavier_uniform_(self.id_embedding.weight)
        nn.init.xavier_uniform_(self.query_feat.weight)
        self.conv = seq(PlainConv(d, d, 3), nn.BatchNorm2d(d), Slot(), PlainConv(d, output_dim, 1),
                        nn.BatchNorm2d(output_dim), Slot())
        self._packs = packs.PackSet()

    def _smooth(self, x):
        t = self.training
        x = ops.conv_bn_act(x, self.conv[0].w(), self.conv[1], t, act="lrelu")
        return ops.conv_bn_act(x, self.conv[3].w(), self.conv[4], t, padding=0, act="lrelu")

    def forward(self, feat, mask_os8, gt_mask_os8=None, temporal_fn=None):
        """feat [b*n_f, C, h, w] channels-last; mask_os8 [b, n_f, n_i, h, w] bool (avg-pool>0 of the input masks);
        gt_mask_os8 (training) [b, n_f, n_i, h, w] bool (max-pool of gt alpha > 0).
        temporal_fn (video): [b, n_f, C, h, w] -> (propagated features, hidden states); the smoothing convs then run
        on the un-propagated features (-> out_feat) and on the propagated ones (-> logits), as the reference does.
        Returns logits [b*n_f, 10, h, w] fp32, out_feat [b*n_f, 64, h, w], tokens [b, 10, 64] fp32, loss (, hidden)."""
        if feat.is_cuda and feat.dtype == torch.float16 and packs.active() is not self._packs:
            # operand packs of every projection that runs on the pixel rows: one grouped preparation per forward
            ws = [self.feat_proj.layers[0].weight]
            for layer in (*self.token_feat_ca_layers, *self.feat_token_ca_layers, self.final_token_feat_ca):
                E = layer.multihead_attn.in_proj_weight.shape[1]
                wq = layer.multihead_attn.in_proj_weight
                ws += [wq[:E], wq[E:2 * E], wq[2 * E:], layer.multihead_attn.out_proj.weight]
            with self._packs.prepare(ws, need_bwd=torch.is_grad_enabled()):
                return self.forward(feat, mask_os8, gt_mask_os8, temporal_fn)
        b, n_f, n_i, h, w = mask_os8.shape
        hw, nq, t = h * w, self.max_inst, self.training
        dt = feat.dtype
        # key index = pixel * n_f + frame (instance_matte_decoder.py:177)
        x = feat.permute(0, 2, 3, 1).reshape(b, n_f, hw, -1).permute(0, 2, 1, 3).reshape(b, hw * n_f, -1)
        ids = torch.arange(1, n_i + 1, device=feat.device).view(1, 1, n_i, 1, 1)
        id_pos = (mask_os8 * ids).amax(2)                                                   # [b,n_f,h,w]
        emb = self.id_embedding.weight
        x_pos = ops.id_embedding(id_pos.reshape(b, n_f, hw).permute(0, 2, 1).reshape(b, hw * n_f), emb, dt)   # [b,S,E]
        x = ops.linear_rows(x, self.feat_proj.layers[0].weight, self.feat_proj.layers[0].bias)
        # the instance tokens (10 per sample) stay fp32 end to end: their layers read the fp32 master weights in place
        tdt = torch.float32 if feat.is_cuda else dt
        tok = self.query_feat.weight.to(tdt)[None].expand(b, -1, -1)
        tok_pos = emb[1:nq + 1].to(tdt)[None].expand(b, -1, -1)

        valid = mask_os8.flatten(3).any(3).any(1)                                           # [b,n_i]
        if n_i < nq:
            valid = torch.cat([valid, valid.new_zeros(b, nq - n_i)], 1)
        tok_pad = ~valid
        guidance = target = None
        if t:
            g = gt_mask_os8.reshape(b, n_f, n_i, hw).permute(0, 2, 3, 1).reshape(b, n_i, hw * n_f)
            if n_i < nq:
                g = torch.cat([g, g.new_zeros(b, nq - n_i, hw * n_f)], 1)
            guidance = g
            target = g.any(2).float()

        loss = 0.0
        pe = self.use_id_pe
        for i in range(self.n_block):
            tok, st = self.token_feat_ca_layers[i](tok, x, tok_pos if pe else None, x_pos if pe else None, None, guidance)
            if t:
                loss = loss + (target - st).sum() / (n_f * b)
            tok = self.mlp_layers[i](tok)
            tok, _ = self.sa_layers[i](tok, tok, tok_pos, tok_pos, tok_pad)
            x, _ = self.feat_token_ca_layers[i](x, tok, x_pos if pe else None, tok_pos if pe else None, tok_pad)
        tok, st = self.final_token_feat_ca(tok, x, tok_pos, x_pos, None, guidance)
        if t:
            loss = loss + (target - st).sum() / (n_f * b)
        loss = loss / (self.n_block + 1)

        x = x.reshape(b, hw, n_f, -1).permute(0, 2, 3, 1).reshape(b * n_f, -1, h, w)
        x = x.contiguous(memory_format=torch.channels_last)
        hidden = out_feat = None
        if temporal_fn is not None:
            prop, hidden = temporal_fn(x.reshape(b, n_f, *x.shape[1:]))
            out_feat = self._smooth(x)
            x = self._smooth(prop.flatten(0, 1).contiguous(memory_format=torch.channels_last))
        else:
            x = out_feat = self._smooth(x)
        tok = ops.small_linear(tok, self.final_mlp.layers[0].weight, self.final_mlp.layers[0].bias)
        tok = F.layer_norm(tok.float(), (tok.shape[-1],), self.decoder_norm.weight, self.decoder_norm.bias,
                           self.decoder_norm.eps)                                           # [b,10,64] fp32
        logits = ops.token_logits(tok, x, n_f)
        if temporal_fn is not None:
            return logits, out_feat, tok, loss, hidden
        return logits, out_feat, tok, loss


# ------------------------------------------------------------------------------------------- full decoder
def _scatter_slots(t, slots, n_slots):
    """[B, len(slots), ...] compact planes -> the reference's zero-padded [B, n_slots, ...] slot layout."""
    out = t.new_zeros((t.shape[0], n_slots) + tuple(t.shape[2:]))
    ops.put(out, 1, slots, t)
    return out


def _draw_widths(n, k_size, is_train):
    """Ellipse sizes for compute_unknown: same numpy draws, in the same order, as utils/utils.py:45-50."""
    # (one vectorised draw: the legacy RandomState yields the same stream as n scalar calls - tests/test_ops_host.py - at
    #  13 us instead of 370 us for the 80 slots of a C2 step, five times per step)
    return np.random.randint(1, k_size, size=n).tolist() if is_train else [k_size // 2] * n


class MaGGIeDecoder(nn.Module):
    """`res_shortcut_inst_matt_spconv_22`."""

    def __init__(self, atten_dim=128, atten_block=2, atten_head=1, final_channel=64, max_inst=10, use_id_pe=True,
                 warmup_mask_atten_iter=0, warmup_detail_iter=3000, detail_mask_dropout=0.2, **_):
        super().__init__()
        assert atten_head == 1, "the reference configs use one attention head"
        if warmup_mask_atten_iter > 0:
            # the reference's mask-attention warm-up (`use_mask_atten`) is off in both live configs and not built here
            raise NotImplementedError("maggie_b200 implements warmup_mask_atten_iter = 0 (both live reference configs)")
        fc = final_channel
        self.max_inst = max_inst
        self.warmup_mask_atten_iter, self.warmup_detail_iter = warmup_mask_atten_iter, warmup_detail_iter
        self.inst_spec_layer = _FFN(fc, fc, 0.1)
        self.layer1 = seq(DecBlock(512, 256, up=True), DecBlock(256, 256))
        self.layer2 = seq(DecBlock(256, 128, up=True), DecBlock(128, 128), DecBlock(128, 128))
        self.refine_OS8 = InstanceMatteDecoder(128, atten_dim, atten_block, fc, max_inst, use_id_pe)
        S, BN = SparseConvParams, nn.BatchNorm1d
        # index-only path of the reference (weights exist, never train, features are discarded)
        self.dummy_downscale = seq(S(3, 32, 3), S(32, 32, 3), S(32, 64, 3), S(64, 64, 3))
        self.layer3 = seq(S(fc, 64, 3), BN(64), Slot(), S(64, 64, 3))
        self.guidance_layer = seq(S(128, 64, 1), BN(64), Slot(), S(64, 64, 3, bias=True), Slot())
        self.layer3_smooth = seq(S(64, 64, 1, bias=True), Slot(), BN(64))
        self.layer4 = seq(S(64, 32, 3), BN(32), Slot(), S(32, 32, 1))
        self.layer4_smooth = seq(S(64, 32, 1, bias=True), Slot(), BN(32))
        self.layer5 = seq(S(32, 32, 3), BN(32), Slot(), S(32, 32, 3))
        self.layer5_smooth = seq(S(64, 32, 1, bias=True), Slot(), BN(32))
        self.refine_OS4 = seq(S(64, 32, 3), BN(32), Slot(), S(32, 1, 3, bias=True))
        self.refine_OS1 = seq(S(32, 32, 3), BN(32), Slot(), S(32, 1, 3, bias=True))
        for p in self.dummy_downscale.parameters():
            p.requires_grad_(True)  # as in the reference: trainable flag set, but they never receive a gradient
        self._packs = packs.PackSet()

    # -- sparse refinement ---------------------------------------------------------------------------
    def predict_details(self, os8_feat, roi, queries, fea1, fea2, fea3, T=None):
        """roi uint8 [B, n_i, H, W]; queries [B, n_i, 64] fp32; T: site tables of `roi` when already built.  Returns
        fp32 logit maps [B*n_i,1,H/4,W/4], [B*n_i,1,H,W] (-99 where inactive) and the site counts."""
        B, n_i, H, W = roi.shape
        slots = B * n_i
        if os8_feat.is_cuda and packs.active() is not self._packs:
            ws = [m.weight for m in self.modules() if isinstance(m, SparseConvParams) and m.weight.shape[-1] % 32 == 0]
            with self._packs.prepare(ws, need_bwd=torch.is_grad_enabled()):
                return self.predict_details(os8_feat, roi, queries, fea1, fea2, fea3, T)
        if T is None:
            T = ops.build_sites(roi.reshape(slots, H, W))
        c1, c2, c4, c8 = T.coords
        dt = os8_feat.dtype
        t = self.training

        nb4, nb1 = T.nbr[2], T.nbr[0]
        conv = lambda src, m, **kw: ops.rows_conv(src, m.weight, m.bias, training=t, **kw)
        subm = lambda src, m, nb, **kw: conv(src, m, table=nb, table_t=nb, mirror=True, **kw)

        x = ops.gather_dense(os8_feat, c8, n_i)
        slot = c8[:, 0].long()
        g = queries.reshape(-1, queries.shape[-1]).index_select(
            0, torch.div(slot, n_i, rounding_mode="floor") * queries.shape[1] + slot % n_i)
        x = self.inst_spec_layer(x * g.to(dt))
        # OS8 -> OS4: inverse conv + BN + LReLU, SubM 3x3
        x = conv(x, self.layer3[0], table=T.parent[2], table_t=T.child[3], bn=self.layer3[1], mode="bn_act", act="lrelu")
        x = subm(x, self.layer3[3], nb4)
        # instance-specific guidance: gate the dense detail features with sigma(conv(cat[detail, instance]))
        det = ops.gather_dense(fea3, c4, n_i)
        gd = conv(torch.cat([det, x], 1), self.guidance_layer[0], bn=self.guidance_layer[1], mode="bn_act", act="lrelu")
        gd = subm(gd, self.guidance_layer[3], nb4)
        x = det * torch.sigmoid(gd.float()).to(dt)
        x = conv(x, self.layer3_smooth[0], bn=self.layer3_smooth[2], mode="act_bn", act="relu")
        y = subm(x, self.refine_OS4[0], nb4, bn=self.refine_OS4[1], mode="bn_act", act="lrelu")
        os4 = ops.rows_head(y, self.refine_OS4[3].weight, self.refine_OS4[3].bias, nb4, c4, slots, H // 4, W // 4)
        # OS4 -> OS2
        x = conv(x, self.layer4[0], table=T.parent[1], table_t=T.child[2], bn=self.layer4[1], mode="bn_act", act="lrelu")
        x = conv(x, self.layer4[3])
        x = torch.cat([ops.gather_dense(fea2, c2, n_i), x], 1)
        x = conv(x, self.layer4_smooth[0], bn=self.layer4_smooth[2], mode="act_bn", act="relu")
        # OS2 -> OS1
        x = conv(x, self.layer5[0], table=T.parent[0], table_t=T.child[1], bn=self.layer5[1], mode="bn_act", act="lrelu")
        x = subm(x, self.layer5[3], nb1)
        x = torch.cat([ops.gather_dense(fea1, c1, n_i), x], 1)
        x = conv(x, self.layer5_smooth[0], bn=self.layer5_smooth[2], mode="act_bn", act="relu")
        y = subm(x, self.refine_OS1[0], nb1, bn=self.refine_OS1[1], mode="bn_act", act="lrelu")
        os1 = ops.rows_head(y, self.refine_OS1[3].weight, self.refine_OS1[3].bias, nb1, c1, slots, H, W)
        return os4, os1, T.counts

    # -- forward --------------------------------------------------------------------------------------
    def dense_stage(self, x, fea4, fea5, mask_os8, gt_os8):
        """OS32 -> OS8 dense blocks + mask-guided attention (static shapes, no host sync: CUDA-graph capturable).
        Returns OS8 logits [B,10,h,w] fp32, OS8 features [B,64,h,w], instance tokens [b,10,64], attention loss."""
        x = self.layer1(x) + fea5
        x = self.layer2(x) + fea4
        return self.refine_OS8(x, mask_os8, gt_os8)

    @staticmethod
    def pooled_masks(masks, gt_alphas, b, n_f, n_i, H, W, training, slots=None, n_slots=None):
        """mask -> OS8 by avg-pool > 0 (utils.py:16-21); GT alpha > 0 -> OS8 by max-pool (:11-15).  With `slots`
        (compact training planes) the small OS8 masks are scattered into the reference's `n_slots` slot layout."""
        def place(m):
            if slots is None:
                return m
            out = m.new_zeros((b, n_f, n_slots, H // 8, W // 8))
            ops.put(out, 2, slots, m)
            return out

        mask_os8 = place(F.avg_pool2d(masks.reshape(b * n_f, n_i, H, W), 8, 8).reshape(b, n_f, n_i, H // 8, W // 8) > 0)
        gt_os8 = None
        if training:
            gt_os8 = place(F.max_pool2d((gt_alphas > 0).float(), 8, 8).reshape(b, n_f, n_i, H // 8, W // 8) > 0)
        return mask_os8, gt_os8

    def _os8_alpha(self, os8_logits, masks, n_i, H, W, slots=None, status=None):
        if slots is not None:       # compact training planes: only the slots that hold an instance are upsampled
            os8_logits = ops.take(os8_logits, 1, slots)
        if self.training:
            valid = (masks.flatten(2).sum(2) > 0).float()          # [B, planes]: alpha of a plane without a mask is zeroed
            flag = status[ops.STATUS_ALL_ZERO:ops.STATUS_ALL_ZERO + 1] if status is not None else None
            return ops.upsample_tanh(os8_logits, size=(H, W), plane_scale=valid, all_zero=flag)
        return ops.upsample_tanh(os8_logits, size=(H, W))[:, :n_i]

    def _choose_guidance(self, iter):
        """The host part of the warm-up switch of resnet_inst_matt_spconv.py:311-316 (same python RNG draw for
        `iter < 3 * warmup`).  The third condition of the reference, `x_os8.sum() == 0`, is decided on the DEVICE (flag
        STATUS_ALL_ZERO of the step's status word, written by the OS8 head kernel): no host read in the middle of the
        step.  (In that degenerate case the reference skips the RNG draw; here it has already been made.)"""
        wd = self.warmup_detail_iter
        return self.training and (iter < wd or (iter < wd * 3 and random.random() < 0.5))

    def plan_roi(self, gt_alphas, iter, status=None):
        """Warm-up iterations take the uncertain region from the ground-truth alphas (resnet_inst_matt_spconv.py:311-316),
        i.e. from an INPUT: its mask and site tables can then be built before the dense stage is even launched, and the
        one host read of the step (site counts + status flags) no longer stalls the middle of the step.
        Returns (unk, site tables) or None."""
        if not (self.training and iter < self.warmup_detail_iter and gt_alphas is not None):
            return None
        unk = ops.unknown_mask(gt_alphas, _draw_widths(gt_alphas.shape[0] * gt_alphas.shape[1], 30, False))
        T = ops.build_sites(unk.reshape(-1, *unk.shape[-2:]), status)
        return (unk, T) if T.counts[0] > 0 else None   # empty set: the regular path handles the degenerate batch

    def _guidance_and_roi(self, a8, gt_alphas, iter, roi_plan, status=None):
        """-> (use_gt, uncertain-region mask, site tables).  Exactly one host read per step: the status word (site counts +
        flags), issued by the `build_sites` call below unless the input stage has already done it (`roi_plan`)."""
        if roi_plan is not None:
            return True, roi_plan[0], roi_plan[1]
        widths = _draw_widths(a8.shape[0] * a8.shape[1], 30, False)
        if self._choose_guidance(iter):
            unk, use_gt = ops.unknown_mask(gt_alphas, widths), True
            T = ops.build_sites(unk.reshape(-1, *unk.shape[-2:]), status)
        elif self.training and status is not None:
            # predicted alpha guides - unless it is all zero, which the device decides (no host read before the mask)
            flag = status[ops.STATUS_ALL_ZERO:ops.STATUS_ALL_ZERO + 1]
            unk = ops.unknown_mask(a8, widths, alt=gt_alphas, use_alt=flag)
            T = ops.build_sites(unk.reshape(-1, *unk.shape[-2:]), status)
            use_gt = bool(T.flags[ops.STATUS_ALL_ZERO - 4])
        else:
            # eval, or the host-side golden tests (CPU tensors, no status word: the reference's host test as it is)
            use_gt = bool(self.training and status is None and float(a8.sum()) == 0)
            unk = ops.unknown_mask(gt_alphas if use_gt else a8, widths)
            T = ops.build_sites(unk.reshape(-1, *unk.shape[-2:]), status)
        return use_gt, unk, T

    def _refine_and_fuse(self, x, queries, fea, a8, unk, use_gt, gt_alphas, b, n_f, H, W, slots=None, n_slots=None, T=None):
        """process_os4_os1 + fuse (resnet_inst_matt_spconv.py:346-366, 272-290, 333-340).  `slots`: the planes are
        compact (plane j = reference slot slots[j] of n_slots); random ellipse sizes are drawn for all reference
        slots, in the reference's order, and the compact planes pick theirs."""
        t = self.training
        fea1, fea2, fea3 = fea
        B = a8.shape[0]
        n_ref = n_slots if slots is not None else a8.shape[1]
        if T is None:
            T = ops.build_sites(unk.reshape(-1, H, W))   # (callers normally pass the tables: see _guidance_and_roi)
        if t and T.counts[0] == 0:
            if slots is not None:
                # degenerate batch: the reference paints the dummy patch into EVERY slot, the empty ones included
                # (their sites enter the sparse BatchNorm statistics) -> fall back to its full slot layout
                full = lambda v: None if v is None else _scatter_slots(v, slots, n_slots)
                ret = self._refine_and_fuse(x, queries, fea, full(a8), full(unk), use_gt, full(gt_alphas), b, n_f, H, W)
                return {k: (ops.take(v, 1, slots) if torch.is_tensor(v) else v) for k, v in ret.items()}
            unk[:, :, 200:250, 200:250] = 1
            T = ops.build_sites(unk.reshape(-1, H, W))
        if slots is not None:
            queries = ops.take(queries, 1, slots)
        pick = (lambda ws: ws) if slots is None else (lambda ws: [ws[f * n_ref + s] for f in range(B) for s in slots])
        widths = lambda k: pick(_draw_widths(B * n_ref, k, t))
        counts = [0, 0, 0, 0]
        if T.counts[0] > 0:
            q = queries[:, None].expand(-1, n_f, -1, -1).reshape(b * n_f, *queries.shape[1:])
            os4, os1, counts = self.predict_details(x, unk, q, fea1, fea2, fea3, T)
            os4 = os4.reshape(b * n_f, a8.shape[1], H // 4, W // 4)
            os1 = os1.reshape(b * n_f, a8.shape[1], H, W)
            a4 = ops.upsample_tanh(os4, scale=4.0)
            a1 = ops.upsample_tanh(os1)
        else:
            a4 = torch.zeros_like(a8)
            a1 = torch.zeros_like(a8)
        ret = dict(alpha_os1=a1, alpha_os4=a4, alpha_os8=a8)
        # progressive fusion (fuse(), :272-290): the weights are {0,1} masks, so `x*w + y*(1-w)` is a select
        # (K10: mask + blend of a stage in one kernel)
        a, w4 = ops.fuse_stage(a8, a4, a8, widths(27), unk)
        a, w1 = ops.fuse_stage(a, a1, a, widths(15), unk)
        ret["refined_masks"] = a
        if not use_gt:
            w4, w1 = w4.to(a8.dtype), w1.to(a8.dtype)
        if use_gt:
            w4 = ops.unknown_mask(gt_alphas, widths(30), and_mask=unk)
            w1 = ops.unknown_mask(gt_alphas, widths(15), and_mask=unk)
        ret["weight_os4"], ret["weight_os1"], ret["detail_mask"] = w4, w1, unk
        ret["site_counts"] = counts
        return ret

    def forward(self, dense_out, fea, image_hw, b, n_f, n_i, masks, iter, gt_alphas, slots=None, n_slots=None,
                roi_plan=None, status=None, **_):
        """dense_out: (os8_logits, os8_feat, queries, loss_atten) from `dense_stage`; fea: (fea1, fea2, fea3);
        masks [b*n_f, n_i, H, W] fp32 {0,1}; gt_alphas [b*n_f, n_i, H, W]; slots / n_slots: see `_refine_and_fuse`;
        status: the step's device status word (`ops.new_status`)."""
        H, W = image_hw
        os8_logits, x, queries, loss_atten = dense_out
        a8 = self._os8_alpha(os8_logits, masks, n_i, H, W, slots, status if roi_plan is None else None)
        use_gt, unk, T = self._guidance_and_roi(a8, gt_alphas, iter, roi_plan, status)
        ret = self._refine_and_fuse(x, queries, fea, a8, unk, use_gt, gt_alphas, b, n_f, H, W, slots, n_slots, T)
        if self.training and iter >= self.warmup_mask_atten_iter:
            ret["loss_max_atten"] = loss_atten
        return ret


# ------------------------------------------------------------------------------------------- video decoder
class MaGGIeTempDecoder(MaGGIeDecoder):
    """`res_shortcut_inst_matt_spconv_temp_22`: ConvGRU on the OS8 features (bidirectional), temporal-difference head,
    bidirectional alpha fusion, eval-time box cropping.  Reference: decoder/resnet_inst_matt_spconv_temp.py:14-203,
    module/conv_gru.py:4-70, utils/utils.py:61-84."""

    def __init__(self, temp_method="bi", **kw):
        super().__init__(**kw)
        self.temp_method = temp_method.split("_")[0]
        self.use_fusion = "fusion" in temp_method
        assert self.temp_method == "bi", "only the live 'bi_fusion' configuration is implemented"
        gru = nn.Module()
        gru.ih = seq(nn.Conv2d(256, 256, 3, padding=1), Slot())
        gru.hh = seq(nn.Conv2d(256, 128, 3, padding=1), Slot())
        self.os8_temp_module = gru
        self.diff_module = seq(SNConv(128, 64, 1), nn.BatchNorm2d(64), Slot(), SNConv(64, 32, 3), nn.BatchNorm2d(32), Slot(),
                               PlainConv(32, 1, 3))
        for m in self.diff_module:  # runs 2(n_f-1) times per forward, one power iteration each: not a bank layer
            m.bankable = False

    # -- ConvGRU -------------------------------------------------------------------------------------------
    def _gru_step(self, x, h):
        g = self.os8_temp_module
        return ops.gru_step(x, h, g.ih[0].weight, g.ih[0].bias, g.hh[0].weight, g.hh[0].bias)

    def propagate(self, feat, prev_h=None):
        """feat [b, n_f, C, h, w] -> (bidirectionally propagated features, forward hidden states) (conv_gru.py:50-70)."""
        n_f = feat.shape[1]
        cl = lambda t: t.contiguous(memory_format=torch.channels_last)
        h = cl(prev_h.to(feat.dtype)) if prev_h is not None else torch.zeros_like(cl(feat[:, 0]))
        fw = []
        for k in range(n_f):
            h = self._gru_step(cl(feat[:, k]), h)
            fw.append(h)
        hb, bw = fw[-1], [None] * (n_f - 1)
        for k in range(n_f - 2, -1, -1):
            hb = self._gru_step(cl(feat[:, k]), hb)
            bw[k] = hb
        out = [((fw[k].float() + bw[k].float()) / 2).to(feat.dtype) for k in range(n_f - 1)] + [fw[-1]]
        return torch.stack(out, 1), torch.stack(fw, 1)

    def dense_stage(self, x, fea4, fea5, mask_os8, gt_os8, mem_feat=None):
        x = self.layer1(x) + fea5
        x = self.layer2(x) + fea4
        return self.refine_OS8(x, mask_os8, gt_os8, temporal_fn=lambda f: self.propagate(f, mem_feat))

    # -- temporal difference + fusion ----------------------------------------------------------------------
    def _diff(self, x):
        dm, t = self.diff_module, self.training
        x = ops.conv_bn_act(x, dm[0].weight(), dm[1], t, padding=0, act="relu")
        x = ops.conv_bn_act(x, dm[3].weight(), dm[4], t, act="relu")
        d = ops.conv_bias(x, dm[6].weight, None).float()
        return F.interpolate(d, scale_factor=8.0, mode="bilinear", align_corners=False)

    def bidirectional_fusion(self, feat, preds):
        n_f = feat.shape[1]
        cat = lambda a, c: torch.cat([feat[:, a], feat[:, c]], 1).contiguous(memory_format=torch.channels_last)
        fd = [self._diff(cat(i - 1, i)) for i in range(1, n_f)]          # frame i-1 -> i
        # frame i -> i-1, in the reference's call order (every call advances the spectral-norm power iteration)
        bd = [self._diff(cat(i, i - 1)) for i in range(n_f - 1, 0, -1)][::-1]
        fd = torch.stack([torch.zeros_like(fd[0])] + fd, 1)
        bd = torch.stack(bd + [torch.zeros_like(bd[-1])], 1)
        # both alpha recurrences + their average in one pass (K11)
        return fd, bd, ops.temporal_fuse(fd, bd, preds.float())

    @staticmethod
    def _gaussian_smoothing(x, sigma=3):
        """utils.py:61-84 (g*g broadcast kernel, un-normalised; crop; bilinear resize back)."""
        ks, pad = sigma * 2 + 1, sigma
        grid = torch.arange(ks, device=x.device).float() - ks // 2
        g = torch.exp(-grid ** 2 / (2 * sigma ** 2))
        g = g / g.sum()
        k = (g * g).view(1, 1, 1, ks).expand(x.shape[1], 1, ks, ks).contiguous()
        y = F.conv2d(F.pad(x, (pad, pad, pad, pad)), k, groups=x.shape[1])[:, :, pad:-pad, pad:-pad]
        return F.interpolate(y, size=x.shape[-2:], mode="bilinear", align_corners=False)

    @staticmethod
    def _loss_dtssd(pred, gt, mask, pad_ratio=1.0):
        """loss.py:9-16.  The reference sums `mask + 1e-6` over all of its planes; `pad_ratio` = reference planes per
        plane given here (compact training planes leave the all-zero slots out)."""
        diff = ((pred[:, 1:] - pred[:, :-1]) - (gt[:, 1:] - gt[:, :-1])) ** 2 * mask[:, 1:]
        if pad_ratio == 1.0:
            return diff.sum() / torch.sum(mask[:, 1:] + 1e-6)
        return diff.sum() / (torch.sum(mask[:, 1:]) + 1e-6 * pad_ratio * mask[:, 1:].numel())

    def forward(self, dense_out, fea, image_hw, b, n_f, n_i, masks, iter, gt_alphas, spar_gt=None, slots=None,
                n_slots=None, roi_plan=None, status=None, **_):
        H, W = image_hw
        t = self.training
        os8_logits, x, queries, loss_atten, hidden = dense_out
        feat_os8 = x.reshape(b, n_f, *x.shape[1:]).detach()
        a8 = self._os8_alpha(os8_logits, masks, n_i, H, W, slots, status if (t and roi_plan is None) else None)
        T = None
        if t:
            use_gt, unk, T = self._guidance_and_roi(a8, gt_alphas, iter, roi_plan, status)
        else:
            use_gt = False
            a8 = torch.where(a8 >= 0.95, torch.ones_like(a8), a8)
            unk = ops.unknown_mask(a8, _draw_widths(a8.shape[0] * a8.shape[1], 30, False))
        if not t:
            # keep only a +-30 px box around each instance's smoothed coarse alpha (one D2H of the box table)
            sm = self._gaussian_smoothing(a8, 3) > 0.1
            ys, xs = sm.any(3), sm.any(2)                                           # [B, n_i, H], [B, n_i, W]
            first = lambda m: m.float().argmax(-1)
            last = lambda m: m.shape[-1] - 1 - m.flip(-1).float().argmax(-1)
            box = torch.stack([first(ys), last(ys), first(xs), last(xs), ys.any(-1).long()], -1).cpu()
            keep = torch.ones(a8.shape, dtype=torch.bool)
            for i in range(a8.shape[0]):
                for j in range(a8.shape[1]):
                    y0, y1, x0, x1, ok = (int(v) for v in box[i, j])
                    if not ok:
                        continue
                    keep[i, j] = False
                    keep[i, j, max(0, y0 - 30):min(y1 + 30, H), max(0, x0 - 30):min(x1 + 30, W)] = True
            keep = keep.to(a8.device)
            unk = unk * keep
            a8 = a8 * keep
        ret = self._refine_and_fuse(x, queries, fea, a8, unk, use_gt, gt_alphas, b, n_f, H, W, slots, n_slots, T)
        ret["mem_feat"] = hidden
        a = ret["refined_masks"]
        fd, bd, fused = self.bidirectional_fusion(feat_os8, a.reshape(b, n_f, *a.shape[1:]))
        ret["temp_alpha"], ret["diff_forward"], ret["diff_backward"] = fused, torch.sigmoid(fd), torch.sigmoid(bd)
        if t:
            ret["loss_max_atten"] = loss_atten
            # the reference reads slot 0 of its scattered transition maps (resnet_inst_matt_spconv_temp.py:189-197): with
            # compact planes that is the plane that was assigned slot 0, or zeros when no instance got that slot
            j0 = 0 if slots is None else (list(slots).index(0) if 0 in slots else -1)
            sg = spar_gt.reshape(fd.shape[0], -1, *spar_gt.shape[1:])[:, 1:, max(j0, 0):max(j0, 0) + 1]
            if j0 < 0:
                sg = torch.zeros_like(sg)
            bce = F.binary_cross_entropy_with_logits(fd[:, 1:, 0], sg[:, :, 0]) + \
                F.binary_cross_entropy_with_logits(bd[:, :-1, 0], sg[:, :, 0])
            ones = torch.ones_like(sg)
            dtf = self._loss_dtssd(torch.sigmoid(fd[:, 1:]), sg, ones)
            dtb = self._loss_dtssd(torch.sigmoid(bd[:, :-1]), sg, ones)
            ret.update(loss_temp_bce=bce, loss_temp_dtssd=dtf + dtb, loss_temp=(bce + dtf + dtb) * 0.25)
        return ret
