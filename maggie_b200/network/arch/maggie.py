"""`MaGGIe` top module: the reference's `model(batch, **kwargs)` contract on top of the B200-native ops.

Reference: arch/maggie.py:18-368.  Input/return contract:
  batch['image'] [b,n_f,3,H,W] float, batch['mask'] [b,n_f,n_i,H,W] or [...,H/8,W/8] {0,1};
  training additionally 'alpha', 'transition' [b,n_f,n_i,H,W] and 'iter'.
  eval  -> dict(refined_masks, alpha_os1, alpha_os4, alpha_os8, detail_mask) each [b,n_f,n_i,H,W]
  train -> (output dict, loss dict with scalar tensors; loss['total'] carries the autograd graph).
numpy / python RNG draws are made in the same order as the reference so that equal seeds give equal draws.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

try:  # keeps from_pretrained / save_pretrained / push_to_hub (arch/maggie.py:18)
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

from ...config import as_cfg
from ... import ops
from .. import loss as losses
from ..decoder import MaGGIeDecoder
from ..encoder import ASPP, ResMaskEmbedShortCutEncoder

ENCODERS = {"res_shortcut_embed_29": ResMaskEmbedShortCutEncoder}
DECODERS = {"res_shortcut_inst_matt_spconv_22": MaGGIeDecoder}


class _DenseStage(nn.Module):
    """encoder + ASPP + OS32->OS8 decoder blocks + mask-guided attention as ONE callable with static shapes and no
    host synchronisation, so that forward and backward can each be replayed as a CUDA graph.  Holds references to
    the model's sub-modules (it is deliberately not registered inside the model: the state dict is unchanged)."""

    def __init__(self, model):
        super().__init__()
        self.encoder, self.aspp, self.decoder = model.encoder, model.aspp, model.decoder

    def forward(self, image, masks, slot_ids, mask_os8, gt_os8):
        emb, fea = self.encoder(image, masks, slot_ids)
        emb = self.aspp(emb)
        logits, feat, queries, loss = self.decoder.dense_stage(emb, fea[3], fea[4], mask_os8 > 0,
                                                               (gt_os8 > 0) if self.training else None)
        if not torch.is_tensor(loss):
            loss = logits.new_zeros(())
        return fea[0], fea[1], fea[2], logits, feat, queries, loss


class MaGGIe(nn.Module, PyTorchModelHubMixin):
    def __init__(self, cfg):
        super().__init__()
        cfg = as_cfg(cfg)
        self.cfg = cfg
        self.num_masks = cfg.encoder_args.num_mask
        if cfg.encoder not in ENCODERS or cfg.decoder not in DECODERS:
            raise NotImplementedError(f"maggie_b200 implements {list(ENCODERS)} + {list(DECODERS)}; "
                                      f"got {cfg.encoder} + {cfg.decoder}")
        self.encoder = ENCODERS[cfg.encoder](**cfg.encoder_args)
        self.aspp = ASPP(cfg.aspp.in_channels, cfg.aspp.out_channels)
        self.decoder = DECODERS[cfg.decoder](**cfg.decoder_args)
        self._stage = [_DenseStage(self)]          # in a list: not a registered sub-module
        self._graphs, self._use_graphs = {}, False
        self.replayed_native_launches = 0          # native kernels executed through graph replays (bench accounting)
        for module in (self.aspp, self.decoder):  # arch/maggie.py:41-49
            for _, p in module.named_parameters():
                if p.dim() > 1:
                    nn.init.xavier_uniform_(p)

    # -------------------------------------------------------------------------------------------
    def _prepare(self, batch):
        x, masks = batch["image"], batch["mask"]
        alphas, trans = batch.get("alpha"), batch.get("transition")
        b, n_f, _, h, w = x.shape
        n_i = masks.shape[2]
        x = x.reshape(b * n_f, 3, h, w)
        masks = masks.reshape(b * n_f, n_i, *masks.shape[-2:]).float()
        if masks.shape[-1] != w:
            masks = F.interpolate(masks, size=(h, w), mode="nearest")
        slot_ids, chosen = list(range(n_i)), None
        dec_masks = masks
        if self.num_masks - n_i > 0 and self.training:
            chosen = np.random.choice(self.num_masks, n_i, replace=False)
            slot_ids = [int(c) for c in chosen]

            def scatter(t):
                out = t.new_zeros((b * n_f, self.num_masks, h, w))
                out[:, chosen] = t.reshape(b * n_f, n_i, h, w)
                return out

            dec_masks = scatter(masks)
            alphas = scatter(alphas.float()) if alphas is not None else None
            trans = scatter(trans.float()) if trans is not None else None
            n_i = self.num_masks
        else:
            alphas = alphas.reshape(b * n_f, n_i, h, w).float() if alphas is not None else None
            trans = trans.reshape(b * n_f, n_i, h, w).float() if trans is not None else None
        return x, masks, slot_ids, dec_masks, alphas, trans, chosen, (b, n_f, n_i, h, w)

    def enable_cuda_graphs(self, on=True):
        """Replay the dense stage (encoder, ASPP, dense decoder blocks, attention; forward AND backward) as CUDA graphs.
        Training mode only; one graph pair per input shape.  The sparse stage stays eager (its shapes follow the
        number of active sites)."""
        self._use_graphs = bool(on)
        return self

    def _dense(self, x, masks, slot_ids, mask_os8, gt_os8):
        stage = self._stage[0]
        stage.train(self.training)
        ids = ops.slot_ids_tensor(slot_ids, x.device)
        args = (x.float().contiguous(), masks.contiguous(), ids, mask_os8.float(),
                gt_os8.float() if gt_os8 is not None else mask_os8.float())
        if not (self._use_graphs and self.training and torch.is_grad_enabled()):
            return stage(*args)
        key = tuple((tuple(a.shape), a.dtype) for a in args)
        entry = self._graphs.get(key)
        if entry is None:
            from ... import _lib
            sample = tuple(a.clone() for a in args)
            before = _lib.launch_count()
            fn = torch.cuda.make_graphed_callables(stage, sample, num_warmup_iters=3, allow_unused_input=True)
            # 3 warm-up iterations + 1 capture, each one forward + backward of the stage
            entry = self._graphs[key] = (fn, (_lib.launch_count() - before) // 4)
        self.replayed_native_launches += entry[1]
        return entry[0](*args)

    def forward(self, batch, **kwargs):
        x, masks, slot_ids, dec_masks, alphas, trans, chosen, (b, n_f, n_i, h, w) = self._prepare(batch)
        mask_os8, gt_os8 = self.decoder.pooled_masks(dec_masks, alphas, b, n_f, n_i, h, w, self.training)
        fea1, fea2, fea3, logits, feat, queries, loss_atten = self._dense(x, masks, slot_ids, mask_os8, gt_os8)
        pred = self.decoder((logits, feat, queries, loss_atten), (fea1, fea2, fea3), (h, w), b=b, n_f=n_f, n_i=n_i,
                            masks=dec_masks, iter=batch.get("iter", 0), gt_alphas=alphas, spar_gt=trans, **kwargs)
        self.last_site_counts = pred.pop("site_counts", None)

        alpha_pred = pred.pop("refined_masks")
        w4 = w1 = pred["detail_mask"].to(alpha_pred.dtype)
        if self.training and np.random.rand() < 0.75:
            w4, w1 = pred.pop("weight_os4"), pred.pop("weight_os1")
        n_out = self.num_masks if (self.training and self.num_masks > 0) else n_i
        view = lambda t: t[:, :n_out].reshape(b, n_f, n_out, h, w)
        output = {k: view(pred[k]) for k in ("alpha_os1", "alpha_os4", "alpha_os8")}
        output["refined_masks"] = view(alpha_pred)
        output["detail_mask"] = view(pred["detail_mask"])

        if self.training:
            valid = (trans.sum((2, 3), keepdim=True) > 0).float()
            for k in list(pred):
                if "loss" in k or "mem_" in k:
                    continue
                pred[k] = pred[k] * valid
            loss_dict = losses.compute_loss(pred, w4, w1, alphas, self.cfg)
            if "loss_max_atten" in pred and self.cfg.loss_atten_w > 0:
                loss_dict["loss_max_atten"] = pred["loss_max_atten"]
                loss_dict["total"] = loss_dict["total"] + loss_dict["loss_max_atten"] * self.cfg.loss_atten_w
            if chosen is not None:
                output = {k: v[:, :, chosen] for k, v in output.items()}
            return output, loss_dict

        output = {k: v[:, :, :n_i] for k, v in output.items()}
        for k in pred:
            if k.startswith("mem_"):
                output[k] = pred[k]
        return output
