"""Plain-PyTorch CPU fp32 restatement of the MaGGIe image hot path (forward; backward via autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py) - the checker for tests/, smoke() and bench.py's cpu_baseline.
Functional style: `P` is a flat mapping {state_dict name: tensor}; SpectralNorm u/v and BatchNorm running
statistics are mutated in place exactly as the reference mutates its parameters/buffers.
Reference files followed (relative to /root/reference/maggie):
  network/arch/maggie.py:63-368             forward, prepare_input, transform_output, compute_loss
  network/encoder/resnet.py:7-39,106-229    BasicBlock, layers, shortcuts, mask-id embedding
  network/module/aspp.py:35-57              ASPP
  network/module/spectral_norm.py:22-35     one power iteration on EVERY forward
  network/decoder/resnet.py:9-45            decoder BasicBlock
  network/decoder/resnet_inst_matt_spconv.py:161-388  decoder forward, sparse refinement, fuse
  network/module/instance_matte_decoder.py:101-306    mask-guided attention at OS8
  network/module/mask_attention.py:9-206    post-norm SA / CA / FFN / MLP around nn.MultiheadAttention
  network/loss.py:67-191                    Sobel gradient loss, Laplacian pyramid loss
  utils/utils.py:7-55                       resizeAnyShape, compute_unknown  (-> oracle/unknown.py)
Pinned against the unmodified reference by oracle/make_golden.py / tests/test_oracle_golden.py.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import unknown as U
from .spconv_torch import _index_map

BN_EPS, BN_MOM, LN_EPS = 1e-5, 0.1, 1e-5


# ----------------------------------------------------------------------------- primitives
def _l2n(v, eps=1e-12):
    return v / (v.norm() + eps)


def sn_weight(P, prefix):
    """spectral_norm.py:22-35. prefix e.g. 'encoder.conv1' -> keys prefix.module.weight_{bar,u,v}."""
    w, u, v = P[prefix + ".module.weight_bar"], P[prefix + ".module.weight_u"], P[prefix + ".module.weight_v"]
    h = w.shape[0]
    wm = w.detach().view(h, -1)
    v.data.copy_(_l2n(torch.mv(wm.t(), u.data)))
    u.data.copy_(_l2n(torch.mv(wm, v.data)))
    sigma = u.detach().dot(w.view(h, -1).mv(v.detach()))
    return w / sigma.expand_as(w)


def bn(P, prefix, x, training):
    return F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"], P[prefix + ".weight"],
                        P[prefix + ".bias"], training, BN_MOM, BN_EPS)


def lrelu(x):
    return F.leaky_relu(x, 0.2)


# ----------------------------------------------------------------------------- encoder + ASPP
def enc_block(P, pre, x, stride, training):
    out = F.conv2d(x, sn_weight(P, pre + ".conv1"), stride=stride, padding=1)
    out = F.relu(bn(P, pre + ".bn1", out, training))
    out = F.conv2d(out, sn_weight(P, pre + ".conv2"), padding=1)
    out = bn(P, pre + ".bn2", out, training)
    idt = x
    if pre + ".downsample.1.module.weight_bar" in P:
        idt = F.avg_pool2d(x, 2, stride)
        idt = bn(P, pre + ".downsample.2", F.conv2d(idt, sn_weight(P, pre + ".downsample.1")), training)
    return F.relu(out + idt)


def shortcut(P, pre, x, training):
    x = F.relu(F.conv2d(x, sn_weight(P, pre + ".0"), padding=1))
    x = bn(P, pre + ".2", x, training)
    x = F.relu(F.conv2d(x, sn_weight(P, pre + ".3"), padding=1))
    return bn(P, pre + ".5", x, training)


def mask_embed(P, x13):
    """encoder/resnet.py:211-229."""
    inp, masks = x13[:, :3], x13[:, 3:]
    ids = torch.arange(1, masks.shape[1] + 1)[None, :, None, None]
    m = (masks * ids).long()
    emb = F.embedding(m, P["encoder.mask_embed_layer.weight"])
    on = (m > 0).float().unsqueeze(-1)
    emb = (emb * on).sum(1) / (on.sum(1) + 1e-6)
    return torch.cat([inp, emb.permute(0, 3, 1, 2)], dim=1)


def encoder(P, x13, training):
    x = mask_embed(P, x13)
    e = "encoder."
    out = F.relu(bn(P, e + "bn1", F.conv2d(x, sn_weight(P, e + "conv1"), stride=2, padding=1), training))
    x1 = F.relu(bn(P, e + "bn2", F.conv2d(out, sn_weight(P, e + "conv2"), padding=1), training))
    out = F.relu(bn(P, e + "bn3", F.conv2d(x1, sn_weight(P, e + "conv3"), stride=2, padding=1), training))
    feats = []
    for name, n, stride in (("layer1", 3, 1), ("layer2", 4, 2), ("layer3", 4, 2), ("layer_bottleneck", 2, 2)):
        for i in range(n):
            out = enc_block(P, f"{e}{name}.{i}", out, stride if i == 0 else 1, training)
        feats.append(out)
    x2, x3, x4, out = feats
    # reference order: trunk first, then the five shortcut branches (resnet.py:194-198)
    fea = [shortcut(P, f"{e}shortcut.{i}", t, training) for i, t in enumerate((x, x1, x2, x3, x4))]
    return out, fea, x[:, :3]


def aspp(P, x, training):
    a = "aspp."
    ys = [F.relu(bn(P, a + "aspp1_bn", F.conv2d(x, P[a + "aspp1.weight"]), training))]
    for i, d in ((2, 2), (3, 4), (4, 8)):
        ys.append(F.relu(bn(P, f"{a}aspp{i}_bn", F.conv2d(x, P[f"{a}aspp{i}.weight"], padding=d, dilation=d), training)))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(bn(P, a + "aspp5_bn", F.conv2d(g, P[a + "aspp5.weight"]), training))
    ys.append(g.expand(-1, -1, x.shape[2], x.shape[3]))
    y = F.conv2d(torch.cat(ys, 1), P[a + "conv2.weight"])
    return F.relu(bn(P, a + "bn2", y, training))


# ----------------------------------------------------------------------------- decoder OS32 -> OS8
def dec_block(P, pre, x, up, training):
    if up:
        out = F.conv_transpose2d(x, sn_weight(P, pre + ".conv1"), stride=2, padding=1)
    else:
        out = F.conv2d(x, sn_weight(P, pre + ".conv1"), padding=1)
    out = lrelu(bn(P, pre + ".bn1", out, training))
    out = bn(P, pre + ".bn2", F.conv2d(out, sn_weight(P, pre + ".conv2"), padding=1), training)
    idt = x
    if up:
        idt = F.interpolate(x, scale_factor=2, mode="nearest")
        idt = bn(P, pre + ".upsample.2", F.conv2d(idt, sn_weight(P, pre + ".upsample.1")), training)
    return lrelu(out + idt)


# ----------------------------------------------------------------------------- attention (IMD)
def layer_norm(P, pre, x):
    return F.layer_norm(x, (x.shape[-1],), P[pre + ".weight"], P[pre + ".bias"], LN_EPS)


def mha(P, pre, q, k, v, key_padding_mask=None):
    """nn.MultiheadAttention, 1 head, seq-first [L,N,E]; returns (out [L,N,E], weights [N,L,S])."""
    E = q.shape[-1]
    w, b = P[pre + ".in_proj_weight"], P[pre + ".in_proj_bias"]
    qp = F.linear(q, w[:E], b[:E]).transpose(0, 1)          # [N,L,E]
    kp = F.linear(k, w[E:2 * E], b[E:2 * E]).transpose(0, 1)  # [N,S,E]
    vp = F.linear(v, w[2 * E:], b[2 * E:]).transpose(0, 1)
    s = torch.bmm(qp, kp.transpose(1, 2)) / math.sqrt(E)
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, :], float("-inf"))
    a = torch.softmax(s, dim=-1)
    o = torch.bmm(a, vp).transpose(0, 1)
    return F.linear(o, P[pre + ".out_proj.weight"], P[pre + ".out_proj.bias"]), a


def cross_attn(P, pre, tgt, mem, pos, qpos, kpm=None):
    if torch.isnan(tgt).any():
        raise ValueError("Mask is empty")  # mask_attention.py:95-98
    q = tgt if qpos is None else tgt + qpos
    k = mem if pos is None else mem + pos
    t2, a = mha(P, pre + ".multihead_attn", q, k, mem, kpm)
    return layer_norm(P, pre + ".norm", tgt + t2), a


def self_attn(P, pre, tgt, qpos, kpm):
    q = tgt + qpos
    t2, _ = mha(P, pre + ".self_attn", q, q, tgt, kpm)
    return layer_norm(P, pre + ".norm", tgt + t2)


def ffn(P, pre, x, p_drop=0.0, training=False):
    h = F.relu(F.linear(x, P[pre + ".linear1.weight"], P[pre + ".linear1.bias"]))
    h = F.dropout(h, p_drop, training)
    h = F.linear(h, P[pre + ".linear2.weight"], P[pre + ".linear2.bias"])
    h = F.dropout(h, p_drop, training)
    return layer_norm(P, pre + ".norm", x + h)


def pool_mask(x, stride, use_max):
    """utils.py:7-25 for the two pooled modes."""
    shp = x.shape
    x = x.reshape(-1, shp[-3], shp[-2], shp[-1]).float()
    x = F.max_pool2d(x, stride, stride) if use_max else (F.avg_pool2d(x, stride, stride) > 0).float()
    return x.view(*shp[:-2], *x.shape[-2:])


def imd(P, feat, mask, gt_mask, training, n_block=2, max_inst=10, use_id_pe=True, temporal_hook=None, captured=None):
    """instance_matte_decoder.py:112-306 with use_mask_atten=False, atten_stride=1, no temporal PE.
    feat [b*n_f,128,h,w]; mask [b,n_f,n_i,H,W] {0,1}. Returns (logits [b*n_f,10,h,w], out_feat [b*n_f,64,h,w],
    tokens [b,10,64], max_loss)."""
    pre = "decoder.refine_OS8."
    stride = int(1 / (feat.shape[-1] * 1.0 / mask.shape[-1]))
    mask = pool_mask(mask, stride, use_max=False)
    b, n_f = mask.shape[:2]
    h, w = feat.shape[-2:]
    C = feat.shape[1]
    feat = feat.view(b, n_f, 1, C, h * w)
    ids = torch.arange(1, mask.shape[2] + 1)[None, None, :, None, None]
    id_pos = (mask * ids).max(2)[0].long()                                   # [b,n_f,h,w]
    emb = P[pre + "id_embedding.weight"]
    feat_pos = F.embedding(id_pos, emb).permute(0, 4, 1, 2, 3)               # [b,c,n_f,h,w]
    feat_pos = feat_pos.permute(0, 2, 1, 3, 4).reshape(b, n_f, 1, -1, h * w)
    tokens = P[pre + "query_feat.weight"][None].repeat(b, 1, 1)             # [b,10,c]
    token_pos = emb[1:max_inst + 1][None].repeat(b, 1, 1)
    feat = feat.permute(4, 2, 1, 0, 3).reshape(h * w * n_f, b, -1)
    feat_pos = feat_pos.permute(4, 2, 1, 0, 3).reshape(h * w * n_f, b, -1)
    feat = F.linear(feat, P[pre + "feat_proj.layers.0.weight"], P[pre + "feat_proj.layers.0.bias"])
    n_i = max_inst
    tokens = tokens.permute(1, 0, 2)
    token_pos = token_pos.permute(1, 0, 2)

    guidance = None
    if training:
        gm = pool_mask(gt_mask, stride, use_max=True)                        # [b,n_f,n_i,h,w]
        m = gm.permute(1, 0, 2, 3, 4).reshape(n_f * b, -1, h * w)
        if m.shape[1] < n_i:
            m = torch.cat([m, torch.zeros(n_f * b, n_i - m.shape[1], h * w)], dim=1)
        guidance = (m > 0).reshape(n_f, b, n_i, -1).permute(1, 2, 3, 0).flatten(2, 3)  # [b,n_i,hw*n_f]

    valid = mask.sum((1, 3, 4)) > 0
    if valid.shape[1] < n_i:
        valid = torch.cat([valid, torch.zeros(b, n_i - valid.shape[1], dtype=torch.bool)], dim=1)
    tok_pad = ~valid

    def atten_loss(a):
        vals = (guidance * a).sum(2)
        gt = torch.ones_like(vals)
        gt[guidance.sum(2) == 0] = 0
        return (gt - vals).sum() / (n_f * b)

    max_loss = 0
    pe = use_id_pe
    for i in range(n_block):
        tokens, a = cross_attn(P, f"{pre}token_feat_ca_layers.{i}", tokens, feat,
                               feat_pos if pe else None, token_pos if pe else None)
        if training:
            max_loss = max_loss + atten_loss(a)
        tokens = ffn(P, f"{pre}mlp_layers.{i}", tokens)
        tokens = self_attn(P, f"{pre}sa_layers.{i}", tokens, token_pos, tok_pad)
        feat, _ = cross_attn(P, f"{pre}feat_token_ca_layers.{i}", feat, tokens,
                             token_pos if pe else None, feat_pos if pe else None, tok_pad)
    tokens, a = cross_attn(P, pre + "final_token_feat_ca", tokens, feat, feat_pos, token_pos)
    if training:
        max_loss = max_loss + atten_loss(a)
    max_loss = max_loss / (n_block + 1)

    feat = feat.reshape(h, w, n_f, b, -1).permute(3, 2, 4, 0, 1).reshape(b * n_f, -1, h, w)

    def smooth(z):
        z = F.conv2d(z, P[pre + "conv.0.weight"], padding=1)
        z = lrelu(bn(P, pre + "conv.1", z, training))
        z = F.conv2d(z, P[pre + "conv.3.weight"])
        return lrelu(bn(P, pre + "conv.4", z, training))

    out_feat = None
    if temporal_hook is not None:            # instance_matte_decoder.py:280-288
        temporal = temporal_hook(feat, b, n_f)
        out_feat = smooth(feat)
        feat = smooth(temporal)
    else:
        feat = smooth(feat)

    tokens = F.linear(tokens, P[pre + "final_mlp.layers.0.weight"], P[pre + "final_mlp.layers.0.bias"])
    tokens = layer_norm(P, pre + "decoder_norm", tokens.permute(1, 0, 2))      # [b,10,64]
    out = torch.einsum("bqc,btchw->btqhw", tokens, feat.reshape(b, n_f, -1, h, w)).flatten(0, 1)
    if temporal_hook is not None:
        return out, out_feat, tokens, max_loss, captured["hidden"]
    return out, feat, tokens, max_loss


# ----------------------------------------------------------------------------- sparse refinement
class Sites:
    """Active-site list at one scale: idx int64 [N,3] (slot,y,x), dense index map, spatial shape."""

    def __init__(self, idx, n_slots, H, W):
        self.idx, self.n_slots, self.H, self.W = idx.long(), n_slots, H, W
        self.imap = _index_map(self.idx, n_slots, H, W)
        self.N = idx.shape[0]


def subm(P, key, feats, S, k):
    """SubMConv2d (spconv_torch.SubMConv2d semantics), weights [Cout,k,k,Cin]."""
    w = P[key + ".weight"]
    if k == 1:
        out = feats @ w[:, 0, 0, :].t()
    else:
        c = k // 2
        out = feats.new_zeros((S.N, w.shape[0]))
        for ky in range(k):
            for kx in range(k):
                ny, nx = S.idx[:, 1] + ky - c, S.idx[:, 2] + kx - c
                ok = (ny >= 0) & (ny < S.H) & (nx >= 0) & (nx < S.W)
                rows = torch.full_like(ny, -1)
                rows[ok] = S.imap[S.idx[ok, 0], ny[ok], nx[ok]]
                sel = (rows >= 0).nonzero(as_tuple=True)[0]
                if sel.numel():
                    out = out.index_add(0, sel, feats[rows[sel]] @ w[:, ky, kx, :].t())
    b = P.get(key + ".bias")
    return out if b is None else out + b


def inverse_conv(P, key, feats, S_coarse, S_fine):
    """SparseInverseConv2d(k=3) undoing SparseConv2d(k3,s2,p1): out[p] = sum_{k: p = 2q-1+k} W[k] in[q]."""
    w = P[key + ".weight"]
    out = feats.new_zeros((S_fine.N, w.shape[0]))
    y, x = S_fine.idx[:, 1], S_fine.idx[:, 2]
    for ky in range(3):
        for kx in range(3):
            ty, tx = y + 1 - ky, x + 1 - kx
            ok = (ty % 2 == 0) & (tx % 2 == 0)
            qy, qx = torch.div(ty, 2, rounding_mode="floor"), torch.div(tx, 2, rounding_mode="floor")
            ok &= (qy >= 0) & (qy < S_coarse.H) & (qx >= 0) & (qx < S_coarse.W)
            sel = ok.nonzero(as_tuple=True)[0]
            if sel.numel():
                rows = S_coarse.imap[S_fine.idx[sel, 0], qy[sel], qx[sel]]  # always active by construction
                out = out.index_add(0, sel, feats[rows] @ w[:, ky, kx, :].t())
    return out


def bn1d(P, pre, x, training):
    if x.shape[0] == 0:
        return x
    return F.batch_norm(x, P[pre + ".running_mean"], P[pre + ".running_var"], P[pre + ".weight"], P[pre + ".bias"],
                        training, BN_MOM, BN_EPS)


def gather_dense(dense, S, n_i):
    """dense [B,C,H,W] -> [N,C] at S (frame = slot // n_i)."""
    return dense[torch.div(S.idx[:, 0], n_i, rounding_mode="floor"), :, S.idx[:, 1], S.idx[:, 2]]


def scatter_logits(vals, S, fill=-99.0):
    out = vals.new_full((S.n_slots, 1, S.H, S.W), 0.0)
    out[S.idx[:, 0], :, S.idx[:, 1], S.idx[:, 2]] = vals
    out = out - 99
    out[S.idx[:, 0], :, S.idx[:, 1], S.idx[:, 2]] += 99
    return out


def predict_details(P, os8_feat, roi, queries, fea1, fea2, fea3, training, p_drop=0.1):
    """resnet_inst_matt_spconv.py:196-270. roi uint8/bool [B,n_i,H,W]; queries [B,10,64]."""
    d = "decoder."
    B, n_i, H, W = roi.shape
    slots = B * n_i
    s1 = torch.from_numpy(U.active_sites(roi.reshape(slots, H, W).cpu().numpy()))
    s2, (H2, W2) = U.downscale_sites(s1.numpy(), H, W)
    s4, (H4, W4) = U.downscale_sites(s2, H2, W2)
    s8, (H8, W8) = U.downscale_sites(s4, H4, W4)
    S1, S2 = Sites(s1, slots, H, W), Sites(torch.from_numpy(s2), slots, H2, W2)
    S4, S8 = Sites(torch.from_numpy(s4), slots, H4, W4), Sites(torch.from_numpy(s8), slots, H8, W8)

    x = gather_dense(os8_feat, S8, n_i)
    g = queries[torch.div(S8.idx[:, 0], n_i, rounding_mode="floor"), S8.idx[:, 0] % n_i]
    x = ffn(P, d + "inst_spec_layer", x * g, p_drop, training)
    # layer3: inverse conv OS8->OS4, BN, LReLU, SubM3x3
    x = inverse_conv(P, d + "layer3.0", x, S8, S4)
    x = subm(P, d + "layer3.3", lrelu(bn1d(P, d + "layer3.1", x, training)), S4, 3)
    # instance-specific guidance
    det = gather_dense(fea3, S4, n_i)
    gde = subm(P, d + "guidance_layer.0", torch.cat([det, x], 1), S4, 1)
    gde = subm(P, d + "guidance_layer.3", lrelu(bn1d(P, d + "guidance_layer.1", gde, training)), S4, 3)
    x = det * torch.sigmoid(gde)
    x = bn1d(P, d + "layer3_smooth.2", F.relu(subm(P, d + "layer3_smooth.0", x, S4, 1)), training)
    # OS4 head
    y = subm(P, d + "refine_OS4.0", x, S4, 3)
    y = subm(P, d + "refine_OS4.3", lrelu(bn1d(P, d + "refine_OS4.1", y, training)), S4, 3)
    os4 = scatter_logits(y, S4)
    # OS2
    x = inverse_conv(P, d + "layer4.0", x, S4, S2)
    x = subm(P, d + "layer4.3", lrelu(bn1d(P, d + "layer4.1", x, training)), S2, 1)
    x = torch.cat([gather_dense(fea2, S2, n_i), x], 1)
    x = bn1d(P, d + "layer4_smooth.2", F.relu(subm(P, d + "layer4_smooth.0", x, S2, 1)), training)
    # OS1
    x = inverse_conv(P, d + "layer5.0", x, S2, S1)
    x = subm(P, d + "layer5.3", lrelu(bn1d(P, d + "layer5.1", x, training)), S1, 3)
    x = torch.cat([gather_dense(fea1, S1, n_i), x], 1)
    x = bn1d(P, d + "layer5_smooth.2", F.relu(subm(P, d + "layer5_smooth.0", x, S1, 1)), training)
    y = subm(P, d + "refine_OS1.0", x, S1, 3)
    y = subm(P, d + "refine_OS1.3", lrelu(bn1d(P, d + "refine_OS1.1", y, training)), S1, 3)
    os1 = scatter_logits(y, S1)
    return os4, os1, dict(N1=S1.N, N2=S2.N, N4=S4.N, N8=S8.N)


def unknown(alpha, k_size, is_train=False):
    """utils.py:28-55 incl. the numpy RNG draws of the training mode (one randint per slice, in order)."""
    n = int(np.prod(alpha.shape[:-2]))
    widths = [np.random.randint(1, k_size) for _ in range(n)] if is_train else [k_size // 2] * n
    return torch.from_numpy(U.compute_unknown(alpha.detach().cpu().numpy(), widths))


def decoder(P, emb, fea, image, b, n_f, n_i, masks, it, gt_alphas, training, cfg, p_drop=0.1):
    """resnet_inst_matt_spconv.py:292-388."""
    d = "decoder."
    da = cfg["decoder_args"]
    fea1, fea2, fea3, fea4, fea5 = fea
    masks5 = masks.reshape(b, n_f, n_i, masks.shape[2], masks.shape[3])
    valid = (masks5.flatten(0, 1).sum((2, 3), keepdim=True) > 0)
    gt_masks = None
    if training:
        gt_masks = (gt_alphas > 0).reshape(b, n_f, n_i, gt_alphas.shape[2], gt_alphas.shape[3])
    x = dec_block(P, d + "layer1.0", emb, True, training)
    x = dec_block(P, d + "layer1.1", x, False, training) + fea5
    x = dec_block(P, d + "layer2.0", x, True, training)
    x = dec_block(P, d + "layer2.1", x, False, training)
    x = dec_block(P, d + "layer2.2", x, False, training) + fea4
    h, w = image.shape[-2:]

    x_os8, x, queries, loss_atten = imd(P, x, masks5, gt_masks, training, da["atten_block"], da["max_inst"], da["use_id_pe"])
    stages = dict(os8_logits=x_os8, os8_feat=x, queries=queries)
    x_os8 = F.interpolate(x_os8, size=(h, w), mode="bilinear", align_corners=False)
    x_os8 = (torch.tanh(x_os8) + 1.0) / 2.0
    x_os8 = x_os8 * valid if training else x_os8[:, :n_i]

    guided = x_os8.clone()
    use_gt = False
    wd = da["warmup_detail_iter"]
    if training and (it < wd or x_os8.sum() == 0 or (it < wd * 3 and random.random() < 0.5)):
        guided, use_gt = gt_alphas.clone(), True
    unk = unknown(guided, 30)
    if unk.max() == 0 and training:
        unk[:, :, 200:250, 200:250] = 1
    if unk.sum() > 0 or training:
        q = queries[:, None].expand(-1, n_f, -1, -1).reshape(b * n_f, *queries.shape[1:])
        os4, os1, counts = predict_details(P, x, unk, q, fea1, fea2, fea3, training, p_drop)
        stages.update(counts)
        os4 = os4.reshape(b * n_f, guided.shape[1], *os4.shape[-2:])
        os1 = os1.reshape(b * n_f, guided.shape[1], *os1.shape[-2:])
        stages.update(os4_logits=os4, os1_logits=os1)
        os4 = F.interpolate(os4, scale_factor=4.0, mode="bilinear", align_corners=False)
        os4 = (torch.tanh(os4) + 1.0) / 2.0
        os1 = (torch.tanh(os1) + 1.0) / 2.0
    else:
        os4 = torch.zeros((b * n_f, x_os8.shape[1], h, w))
        os1 = torch.zeros_like(os4)
    ret = dict(alpha_os1=os1, alpha_os4=os4, alpha_os8=x_os8)
    # fuse (:272-290)
    a = x_os8
    w4 = ((unknown(a, 27, training) * unk) > 0).type(a.dtype)
    a = os4 * w4 + a * (1 - w4)
    w1 = ((unknown(a, 15, training) * unk) > 0).type(a.dtype)
    a = os1 * w1 + a * (1 - w1)
    ret["refined_masks"] = a
    if use_gt:
        w4 = unknown(gt_alphas, 30, training) * unk
        w1 = unknown(gt_alphas, 15, training) * unk
    ret["weight_os4"], ret["weight_os1"], ret["detail_mask"] = w4, w1, unk
    if training and it >= da["warmup_mask_atten_iter"]:
        ret["loss_max_atten"] = loss_atten
    return ret, stages


# ----------------------------------------------------------------------------- losses
def regression_loss(logit, target, weight):
    loss = F.l1_loss(logit * weight, target * weight, reduction="none")
    return loss.sum() / (torch.sum(weight) + 1e-8)


_SOBEL_X = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]]) / 8.0


def sobel(x, eps=1e-6):
    n, c, h, w = x.shape
    xp = F.pad(x.reshape(n * c, 1, h, w), [1, 1, 1, 1], mode="replicate")
    gx = F.conv2d(xp, _SOBEL_X[None, None])
    gy = F.conv2d(xp, _SOBEL_X.t()[None, None])
    return torch.sqrt(gx * gx + gy * gy + eps).reshape(n, c, h, w)


def grad_loss(logit, label, mask, eps=1e-6):
    return torch.sum(F.l1_loss(sobel(logit * mask), sobel(label * mask), reduction="none")) / (mask.sum() + eps)


_GAUSS = torch.tensor([[1., 4., 6., 4., 1.], [4., 16., 24., 16., 4.], [6., 24., 36., 24., 6.],
                       [4., 16., 24., 16., 4.], [1., 4., 6., 4., 1.]]) / 256.0


def _conv_gauss(img, kern):
    img = F.pad(img, (2, 2, 2, 2), mode="reflect")
    return F.conv2d(img, kern[None, None].repeat(img.shape[1], 1, 1, 1), groups=img.shape[1])


def _lap_pyramid(img, levels=3):
    cur, pyr = img, []
    for _ in range(levels):
        down = _conv_gauss(cur, _GAUSS)[:, :, ::2, ::2]
        up = torch.zeros(down.shape[0], down.shape[1], down.shape[2] * 2, down.shape[3] * 2)
        up[:, :, ::2, ::2] = down
        up = _conv_gauss(up, 4 * _GAUSS)
        pyr.append(cur - up)
        cur = down
    return pyr


def lap_loss(inp, tgt, weight, levels=3):
    pi, pt = _lap_pyramid(inp, levels), _lap_pyramid(tgt, levels)
    total, wcur = 0, weight
    for i in range(levels):
        # LapLoss() is built with channels=3 (loss.py:170-173) but fed 1-channel images: the [3,1,5,5] kernel with
        # groups=1 yields 3 identical channels, so every level's weighted L1 sum is counted 3x.
        total = total + 3.0 * (F.l1_loss(pi[i], pt[i], reduction="none") * wcur).sum() / (wcur.sum() + 1e-6)
        wcur = wcur[:, :, ::2, ::2]
    return total


def compute_loss(pred, w4, w1, alphas, cfg):
    """arch/maggie.py:268-368 (image model: dtSSD weight 0)."""
    a1, a4, a8 = pred["alpha_os1"], pred["alpha_os4"], pred["alpha_os8"]
    L = {}
    w8 = torch.ones_like(a8) * (alphas.sum((2, 3), keepdim=True) > 0)
    if cfg["loss_reweight_os8"]:
        ug = (alphas <= 254.0 / 255.0) & (alphas >= 1.0 / 255.0)
        up = (a8 <= 254.0 / 255.0) & (a8 >= 1.0 / 255.0)
        w8 = (ug | up).type(w8.dtype) + w8
    total = 0
    r1, r4, r8 = regression_loss(a1, alphas, w1), regression_loss(a4, alphas, w4), regression_loss(a8, alphas, w8)
    L.update(loss_rec_os1=r1, loss_rec_os4=r4, loss_rec_os8=r8, loss_rec=r1 * 2 + r4 + r8)
    total = total + L["loss_rec"] * cfg["loss_alpha_w"]
    h, w = a8.shape[-2:]
    v = lambda t: t.reshape(-1, 1, h, w)
    l1, l4, l8 = lap_loss(v(a1), v(alphas), v(w1)), lap_loss(v(a4), v(alphas), v(w4)), lap_loss(v(a8), v(alphas), v(w8))
    L.update(loss_lap_os1=l1, loss_lap_os4=l4, loss_lap_os8=l8, loss_lap=l1 * 2 + l4 + l8)
    total = total + L["loss_lap"] * cfg["loss_alpha_lap_w"]
    g1, g4, g8 = grad_loss(a1, alphas, w1), grad_loss(a4, alphas, w4), grad_loss(a8, alphas, w8)
    L.update(loss_grad_os1=g1, loss_grad_os4=g4, loss_grad_os8=g8, loss_grad=g1 * 2 + g4 + g8)
    total = total + L["loss_grad"] * cfg["loss_alpha_grad_w"]
    L["total"] = total
    return L


# ----------------------------------------------------------------------------- top level
def forward(P, batch, training, cfg, return_stages=False, p_drop=0.1):
    """MaGGIe.forward (arch/maggie.py:63-139). Returns eval: output dict; train: (output, loss_dict)."""
    num_masks = cfg["encoder_args"]["num_mask"]
    x, masks = batch["image"], batch["mask"]
    alphas, trans = batch.get("alpha"), batch.get("transition")
    b, n_f, _, h, w = x.shape
    n_i = masks.shape[2]
    x = x.view(-1, 3, h, w)
    if masks.shape[-1] != w:
        masks = F.interpolate(masks.flatten(0, 1), size=(h, w), mode="nearest")
    else:
        masks = masks.view(-1, n_i, h, w)
    chosen = None
    inp_masks = masks
    if num_masks - n_i > 0:
        if not training:
            inp_masks = torch.cat([masks, torch.zeros(b * n_f, num_masks - n_i, h, w)], dim=1)
        else:
            chosen = np.random.choice(num_masks, n_i, replace=False)
            inp_masks = torch.zeros(b * n_f, num_masks, h, w)
            inp_masks[:, chosen] = masks
            masks = inp_masks
            if alphas is not None:
                na = torch.zeros(b, n_f, num_masks, h, w)
                na[:, :, chosen] = alphas
                alphas = na
            if trans is not None:
                nt = torch.zeros(b, n_f, num_masks, h, w)
                nt[:, :, chosen] = trans
                trans = nt
            n_i = num_masks
    inp = torch.cat([x, inp_masks], dim=1)
    if alphas is not None:
        alphas = alphas.view(-1, n_i, h, w)
    if trans is not None:
        trans = trans.view(-1, n_i, h, w)
    emb, fea, image = encoder(P, inp, training)
    emb = aspp(P, emb, training)
    pred, stages = decoder(P, emb, fea, image, b, n_f, n_i, masks, batch.get("iter", 0), alphas, training, cfg, p_drop)
    stages["aspp"] = emb

    alpha_pred = pred.pop("refined_masks")
    w4 = w1 = pred["detail_mask"].type(alpha_pred.dtype)
    if training and np.random.rand() < 0.75:
        w4, w1 = pred.pop("weight_os4"), pred.pop("weight_os1")
    n_out = num_masks if (training and num_masks > 0) else n_i
    out = {k: pred[k][:, :n_out].view(b, n_f, n_out, h, w) for k in ("alpha_os1", "alpha_os4", "alpha_os8")}
    out["refined_masks"] = alpha_pred[:, :n_out].view(b, n_f, n_out, h, w)
    out["detail_mask"] = pred["detail_mask"][:, :n_out].view(b, n_f, n_out, h, w)
    if training:
        valid = (trans.sum((2, 3), keepdim=True) > 0).float()
        for k, v in pred.items():
            if "loss" in k or "mem_" in k:
                continue
            pred[k] = v * valid
        L = compute_loss(pred, w4, w1, alphas, cfg)
        if "loss_max_atten" in pred and cfg["loss_atten_w"] > 0:
            L["loss_max_atten"] = pred["loss_max_atten"]
            L["total"] = L["total"] + L["loss_max_atten"] * cfg["loss_atten_w"]
        if chosen is not None:
            out = {k: v[:, :, chosen] for k, v in out.items()}
        return (out, L, stages) if return_stages else (out, L)
    out = {k: v[:, :, :n_i] for k, v in out.items()}
    return (out, stages) if return_stages else out


# ============================================================================================== video model
# MaGGIe_Temp (arch/maggie_temp.py), ResShortCut_InstMattSpconv_BiTempSpar_Dec
# (decoder/resnet_inst_matt_spconv_temp.py), ConvGRU (module/conv_gru.py), gaussian_smoothing (utils/utils.py:61-84),
# loss_dtSSD (loss.py:7-16,41-44).
def conv_gru_step(P, pre, x, h):
    """conv_gru.py:23-28."""
    C = x.shape[1]
    rz = torch.sigmoid(F.conv2d(torch.cat([x, h], 1), P[pre + ".ih.0.weight"], P[pre + ".ih.0.bias"], padding=1))
    r, z = rz.split(C, dim=1)
    c = torch.tanh(F.conv2d(torch.cat([x, r * h], 1), P[pre + ".hh.0.weight"], P[pre + ".hh.0.bias"], padding=1))
    return (1 - z) * h + z * c


def propagate_bi(P, pre, feat, prev_h):
    """conv_gru.py:50-70 with temp_method='bi'. feat [b,n_f,C,h,w] -> (feat', hidden states [b,n_f,C,h,w])."""
    b, n_f = feat.shape[:2]
    h = prev_h if prev_h is not None else torch.zeros_like(feat[:, 0])
    fw = []
    for t in range(n_f):
        h = conv_gru_step(P, pre, feat[:, t], h)
        fw.append(h)
    hidden = torch.stack(fw, 1)
    hb, bw = fw[-1], [None] * (n_f - 1)
    for t in range(n_f - 2, -1, -1):
        hb = conv_gru_step(P, pre, feat[:, t], hb)
        bw[t] = hb
    out = [(fw[t] + bw[t]) / 2 for t in range(n_f - 1)] + [fw[-1]]
    return torch.stack(out, 1), hidden


def imd_temp(P, feat, mask, gt_mask, training, mem_feat, **kw):
    """instance_matte_decoder.py:112-306 with aggregate_mem_fn: the smoothing convs run twice (no-temporal features,
    then ConvGRU-propagated features)."""
    captured = {}

    def hook(x, b, n_f):
        captured["no_temp"] = x
        f5, hidden = propagate_bi(P, "decoder.os8_temp_module", x.reshape(b, n_f, *x.shape[1:]), mem_feat)
        captured["hidden"] = hidden
        return f5.flatten(0, 1)

    return imd(P, feat, mask, gt_mask, training, temporal_hook=hook, captured=captured, **kw)


def gaussian_smoothing(x, sigma=3):
    """utils.py:61-84 (not a true 2-D Gaussian: the kernel is g*g broadcast over rows, un-normalised; then the
    result is cropped by the padding once more and bilinearly resized back)."""
    ks = sigma * 2 + 1
    pad = ks // 2
    grid = torch.arange(ks).float() - ks // 2
    g = torch.exp(-grid ** 2 / (2 * sigma ** 2))
    g = g / g.sum()
    k = (g.view(1, 1, -1) * g.view(1, 1, -1)).expand(x.shape[1], 1, ks, ks).type_as(x)
    y = F.conv2d(F.pad(x, (pad, pad, pad, pad)), k, groups=x.shape[1])
    y = y[:, :, pad:-pad, pad:-pad]
    return F.interpolate(y, size=x.shape[-2:], mode="bilinear", align_corners=False)


def diff_module(P, x, training):
    d = "decoder.diff_module."
    x = F.relu(bn(P, d + "1", F.conv2d(x, sn_weight(P, d + "0")), training))
    x = F.relu(bn(P, d + "4", F.conv2d(x, sn_weight(P, d + "3"), padding=1), training))
    return F.conv2d(x, P[d + "6.weight"], padding=1)


def bidirectional_fusion(P, feat, preds, training):
    """resnet_inst_matt_spconv_temp.py:35-79. feat [b,n_f,64,h,w] (detached), preds [b,n_f,n_i,H,W]."""
    n_f = feat.shape[1]
    up = lambda d: F.interpolate(d, scale_factor=8.0, mode="bilinear", align_corners=False)
    fdiffs, fpreds = [], [preds[:, 0]]
    for i in range(1, n_f):
        d = up(diff_module(P, torch.cat([feat[:, i - 1], feat[:, i]], 1), training))
        fdiffs.append(d)
        fpreds.append(fpreds[-1] * (1 - d.sigmoid()) + preds[:, i] * d.sigmoid())
    fdiffs = torch.stack([torch.zeros_like(fdiffs[0])] + fdiffs, 1)
    bdiffs, bpreds = [], [preds[:, n_f - 1]]
    for i in range(n_f - 1, 0, -1):
        d = up(diff_module(P, torch.cat([feat[:, i], feat[:, i - 1]], 1), training))
        bdiffs.append(d)
        bpreds.append(bpreds[-1] * (1 - d.sigmoid()) + preds[:, i - 1] * d.sigmoid())
    bpreds, bdiffs = bpreds[::-1], bdiffs[::-1]
    bdiffs = torch.stack(bdiffs + [torch.zeros_like(bdiffs[-1])], 1)
    fused = [fpreds[0]] + [(fpreds[i] + bpreds[i]) / 2 for i in range(1, n_f - 1)] + [bpreds[n_f - 1]]
    return fdiffs, bdiffs, torch.stack(fused, 1)


def loss_dtssd(pred, gt, mask):
    """loss.py:7-16."""
    diff = ((pred[:, 1:] - pred[:, :-1]) - (gt[:, 1:] - gt[:, :-1])) ** 2 * mask[:, 1:]
    return diff.sum() / torch.sum(mask[:, 1:] + 1e-6)


def decoder_temp(P, emb, fea, image, b, n_f, n_i, masks, it, gt_alphas, spar_gt, training, cfg, mem_feat=None, p_drop=0.1):
    """resnet_inst_matt_spconv_temp.py:81-203."""
    d = "decoder."
    da = cfg["decoder_args"]
    fea1, fea2, fea3, fea4, fea5 = fea
    masks5 = masks.reshape(b, n_f, n_i, masks.shape[2], masks.shape[3])
    valid = (masks5.flatten(0, 1).sum((2, 3), keepdim=True) > 0)
    gt_masks = (gt_alphas > 0).reshape(b, n_f, n_i, *gt_alphas.shape[2:]) if training else None
    x = dec_block(P, d + "layer1.0", emb, True, training)
    x = dec_block(P, d + "layer1.1", x, False, training) + fea5
    x = dec_block(P, d + "layer2.0", x, True, training)
    x = dec_block(P, d + "layer2.1", x, False, training)
    x = dec_block(P, d + "layer2.2", x, False, training) + fea4
    h, w = image.shape[-2:]
    x_os8, x, queries, loss_atten, hidden = imd_temp(P, x, masks5, gt_masks, training, mem_feat, n_block=da["atten_block"],
                                                     max_inst=da["max_inst"], use_id_pe=da["use_id_pe"])
    feat_os8 = x.view(b, n_f, *x.shape[1:]).detach()
    x_os8 = (torch.tanh(F.interpolate(x_os8, scale_factor=8.0, mode="bilinear", align_corners=False)) + 1.0) / 2.0
    x_os8 = x_os8 * valid if training else x_os8[:, :n_i]
    guided, use_gt = x_os8, False
    wd = da["warmup_detail_iter"]
    if training and (it < wd or x_os8.sum() == 0 or (it < wd * 3 and random.random() < 0.5)):
        guided, use_gt = gt_alphas.clone(), True
    if not training:
        x_os8[x_os8 >= 0.95] = 1.0
    unk = unknown(guided, 30)
    if not training:
        smooth = gaussian_smoothing(x_os8, sigma=3)
        for i in range(smooth.shape[0]):
            for j in range(n_i):
                ys, xs = torch.nonzero(smooth[i, j] > 0.1, as_tuple=True)
                if len(ys) == 0:
                    continue
                y0, y1 = max(0, int(ys.min()) - 30), min(int(ys.max()) + 30, h)
                x0, x1 = max(0, int(xs.min()) - 30), min(int(xs.max()) + 30, w)
                tm = torch.zeros((h, w), dtype=torch.bool)
                tm[y0:y1, x0:x1] = True
                unk[i, j] = unk[i, j] * tm
                x_os8[i, j] = x_os8[i, j] * tm
    if unk.max() == 0 and training:
        unk[:, :, 200:250, 200:250] = 1
    if unk.sum() > 0 or training:
        q = queries[:, None].expand(-1, n_f, -1, -1).reshape(b * n_f, *queries.shape[1:])
        os4, os1, _ = predict_details(P, x, unk, q, fea1, fea2, fea3, training, p_drop)
        os4 = os4.reshape(b * n_f, guided.shape[1], *os4.shape[-2:])
        os1 = os1.reshape(b * n_f, guided.shape[1], *os1.shape[-2:])
        os4 = (torch.tanh(F.interpolate(os4, scale_factor=4.0, mode="bilinear", align_corners=False)) + 1.0) / 2.0
        os1 = (torch.tanh(os1) + 1.0) / 2.0
    else:
        os4 = torch.zeros((b * n_f, x_os8.shape[1], h, w))
        os1 = torch.zeros_like(os4)
    ret = dict(alpha_os1=os1, alpha_os4=os4, alpha_os8=x_os8)
    a = x_os8
    w4 = ((unknown(a, 27, training) * unk) > 0).type(a.dtype)
    a = os4 * w4 + a * (1 - w4)
    w1 = ((unknown(a, 15, training) * unk) > 0).type(a.dtype)
    a = os1 * w1 + a * (1 - w1)
    ret["refined_masks"], ret["detail_mask"], ret["mem_feat"] = a, unk, hidden
    if use_gt:
        w4 = unknown(gt_alphas, 30, training) * unk
        w1 = unknown(gt_alphas, 15, training) * unk
    ret["weight_os4"], ret["weight_os1"] = w4, w1
    fd, bd, fused = bidirectional_fusion(P, feat_os8, a.view(b, n_f, *a.shape[1:]), training)
    ret["temp_alpha"], ret["diff_forward"], ret["diff_backward"] = fused, fd.sigmoid(), bd.sigmoid()
    if training:
        ret["loss_max_atten"] = loss_atten
        sg = spar_gt.view(fd.shape[0], -1, *spar_gt.shape[1:])
        bce = F.binary_cross_entropy_with_logits(fd[:, 1:, 0], sg[:, 1:, 0]) + \
            F.binary_cross_entropy_with_logits(bd[:, :-1, 0], sg[:, 1:, 0])
        ones = torch.ones_like(sg[:, 1:, 0:1])
        dtf = loss_dtssd(fd[:, 1:].sigmoid(), sg[:, 1:, 0:1], ones)
        dtb = loss_dtssd(bd[:, :-1].sigmoid(), sg[:, 1:, 0:1], ones)
        ret.update(loss_temp_bce=bce, loss_temp_dtssd=dtf + dtb, loss_temp=(bce + dtf + dtb) * 0.25)
    return ret


def forward_video(P, batch, training, cfg, mem_feat=None, prev_pred=None, p_drop=0.1):
    """MaGGIe_Temp.forward (arch/maggie_temp.py:34-77 on top of arch/maggie.py:63-139)."""
    num_masks = cfg["encoder_args"]["num_mask"]
    x, masks = batch["image"], batch["mask"]
    alphas, trans = batch.get("alpha"), batch.get("transition")
    b, n_f, _, h, w = x.shape
    n_i = masks.shape[2]
    x = x.view(-1, 3, h, w)
    masks = F.interpolate(masks.flatten(0, 1), size=(h, w), mode="nearest") if masks.shape[-1] != w else masks.view(-1, n_i, h, w)
    chosen, inp_masks = None, masks
    if num_masks - n_i > 0:
        if not training:
            inp_masks = torch.cat([masks, torch.zeros(b * n_f, num_masks - n_i, h, w)], 1)
        else:
            chosen = np.random.choice(num_masks, n_i, replace=False)
            inp_masks = torch.zeros(b * n_f, num_masks, h, w)
            inp_masks[:, chosen] = masks
            masks = inp_masks
            na, nt = torch.zeros(b, n_f, num_masks, h, w), torch.zeros(b, n_f, num_masks, h, w)
            na[:, :, chosen], nt[:, :, chosen] = alphas, trans
            alphas, trans, n_i = na, nt, num_masks
    if alphas is not None:
        alphas, trans = alphas.view(-1, n_i, h, w), trans.view(-1, n_i, h, w)
    emb, fea, image = encoder(P, torch.cat([x, inp_masks], 1), training)
    emb = aspp(P, emb, training)
    pred = decoder_temp(P, emb, fea, image, b, n_f, n_i, masks, batch.get("iter", 0), alphas, trans, training, cfg, mem_feat, p_drop)
    alpha_pred = pred.pop("refined_masks")
    w4 = w1 = pred["detail_mask"].type(alpha_pred.dtype)
    if training and np.random.rand() < 0.75:
        w4, w1 = pred.pop("weight_os4"), pred.pop("weight_os1")
    n_out = num_masks if (training and num_masks > 0) else n_i
    v5 = lambda t: t[:, :n_out].view(b, n_f, n_out, h, w)
    out = {k: v5(pred[k]) for k in ("alpha_os1", "alpha_os4", "alpha_os8")}
    out["refined_masks"], out["detail_mask"] = v5(alpha_pred), v5(pred["detail_mask"])
    db, df, ta = pred.pop("diff_backward"), pred.pop("diff_forward"), pred.pop("temp_alpha")
    out["diff_pred_backward"], out["diff_pred_forward"], out["temp_alpha"] = db.repeat(1, 1, n_i, 1, 1), df.repeat(1, 1, n_i, 1, 1), ta
    if training:
        valid = (trans.sum((2, 3), keepdim=True) > 0).float()
        for k, v in pred.items():
            if "loss" in k or "mem_" in k:
                continue
            pred[k] = v * valid
        L = compute_loss(pred, w4, w1, alphas, cfg)
        if cfg["loss_dtSSD_w"] > 0:
            shp = (b, n_f, num_masks, h, w)
            r = lambda t: t.reshape(*shp)
            d1 = loss_dtssd(r(pred["alpha_os1"]), r(alphas), r(w1))
            d4 = loss_dtssd(r(pred["alpha_os4"]), r(alphas), r(w4))
            w8 = torch.ones_like(pred["alpha_os8"]) * (alphas.sum((2, 3), keepdim=True) > 0)
            if cfg["loss_reweight_os8"]:
                ug = (alphas <= 254.0 / 255.0) & (alphas >= 1.0 / 255.0)
                up = (pred["alpha_os8"] <= 254.0 / 255.0) & (pred["alpha_os8"] >= 1.0 / 255.0)
                w8 = (ug | up).type(w8.dtype) + w8
            d8 = loss_dtssd(r(pred["alpha_os8"]), r(alphas), r(w8))
            L.update(loss_dtSSD_os1=d1, loss_dtSSD_os4=d4, loss_dtSSD_os8=d8, loss_dtSSD=d1 * 2 + d4 + d8)
            L["total"] = L["total"] + L["loss_dtSSD"] * cfg["loss_dtSSD_w"]
        if cfg["loss_atten_w"] > 0:
            L["loss_max_atten"] = pred["loss_max_atten"]
            L["total"] = L["total"] + L["loss_max_atten"] * cfg["loss_atten_w"]
        L.update(loss_temp_bce=pred["loss_temp_bce"], loss_temp=pred["loss_temp"], loss_temp_dtssd=pred["loss_temp_dtssd"])
        L["total"] = L["total"] + pred["loss_temp"]
        if chosen is not None:
            out = {k: v[:, :, chosen] for k, v in out.items()}
        return out, L
    out = {k: v[:, :, :n_i] for k, v in out.items()}
    out["mem_feat"] = pred["mem_feat"]
    # alpha-matte level aggregation over the 3-frame window (maggie_temp.py:37-75)
    al = out["refined_masks"]
    prev = al[:, 0] if prev_pred is None else prev_pred
    nxt = al[:, -1]
    dfw = (out["diff_pred_forward"] > 0.5).float()
    dbw = (out["diff_pred_backward"] > 0.5).float()
    p01 = prev * (1 - dfw[:, 1]) + al[:, 1] * dfw[:, 1]
    p21 = nxt * (1 - dbw[:, 1]) + al[:, 1] * dbw[:, 1]
    diff = torch.abs(p01 - p21)
    p01[diff > 0.0] = al[:, 1][diff > 0.0]
    al[:, 1] = p01
    al[:, 2] = p01 * (1 - dfw[:, 2]) + nxt * dfw[:, 2]
    return out
