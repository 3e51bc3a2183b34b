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
from ... import dense, ops
from ...weights import WeightBank
from .. import loss as losses
from ..decoder import MaGGIeDecoder, MaGGIeTempDecoder
from ..encoder import ASPP, ResMaskEmbedShortCutEncoder

ENCODERS = {"res_shortcut_embed_29": ResMaskEmbedShortCutEncoder}
DECODERS = {"res_shortcut_inst_matt_spconv_22": MaGGIeDecoder, "res_shortcut_inst_matt_spconv_temp_22": MaGGIeTempDecoder}


class _DenseStage(nn.Module):
    """encoder + ASPP + OS32->OS8 decoder blocks + mask-guided attention as ONE callable with static shapes and no
    host synchronisation, so that forward and backward can each be replayed as a CUDA graph.  Holds references to
    the model's sub-modules (it is deliberately not registered inside the model: the state dict is unchanged)."""

    def __init__(self, model):
        super().__init__()
        self.encoder, self.aspp, self.decoder = model.encoder, model.aspp, model.decoder
        self.bank = WeightBank().attach(self)

    def forward(self, image, masks, slot_ids, mask_os8, gt_os8, mem_feat=None):
        with ops.step_scope("dense_stage", image.device):
            if getattr(self.encoder, "precision", "fp16") == "high" and not self.training:
                # fp32-accurate evaluation: fp32 weights straight from the containers (per-layer spectral norm), no fp16 packs
                return self._forward(image, masks, slot_ids, mask_os8, gt_os8, mem_feat)
            ops.prepare_weights(self.bank)  # K0: spectral norm + operand packs of every dense conv in one grouped op
            try:
                return self._forward(image, masks, slot_ids, mask_os8, gt_os8, mem_feat)
            finally:
                self.bank.release()

    def _forward(self, image, masks, slot_ids, mask_os8, gt_os8, mem_feat=None):
        emb, fea = self.encoder(image, masks, slot_ids)
        emb = self.aspp(emb)
        extra = {} if mem_feat is None else {"mem_feat": mem_feat}
        out = list(self.decoder.dense_stage(emb, fea[3], fea[4], mask_os8 > 0, (gt_os8 > 0) if self.training else None,
                                            **extra))
        if not torch.is_tensor(out[3]):
            out[3] = out[0].new_zeros(())
        return (fea[0], fea[1], fea[2], *out)


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

    @property
    def bank(self):
        """The dense stage's weight bank (K0): `FlatGradAllReduce(model.parameters(), bank=model.bank)` lets its grouped
        backward write the conv-weight gradients straight into the data-parallel flat buffer."""
        return self._stage[0].bank

    def _prepare(self, batch):
        """Flatten the batch dict.  Training with fewer instances than mask slots draws the reference's random slot
        assignment (arch/maggie.py:206-229) but keeps every pixel-sized tensor COMPACT (one plane per real instance):
        the reference's scattered `[B, num_masks, H, W]` tensors are zero in every other slot, and zero planes
        contribute exactly nothing to alphas, weights, active sites or losses.  `chosen[j]` is the slot of plane j."""
        x, masks = batch["image"], batch["mask"]
        alphas, trans = batch.get("alpha"), batch.get("transition")
        b, n_f, _, h, w = x.shape
        n_i = masks.shape[2]
        x = x.reshape(b * n_f, 3, h, w)
        masks = masks.reshape(b * n_f, n_i, *masks.shape[-2:]).float()
        if masks.shape[-1] != w:
            masks = F.interpolate(masks, size=(h, w), mode="nearest")
        slot_ids, chosen, n_slots, unsort = list(range(n_i)), None, n_i, None
        alphas = alphas.reshape(b * n_f, n_i, h, w).float() if alphas is not None else None
        trans = trans.reshape(b * n_f, n_i, h, w).float() if trans is not None else None
        if self.num_masks - n_i > 0 and self.training:
            draw = [int(c) for c in np.random.choice(self.num_masks, n_i, replace=False)]
            # On the host (golden tests) planes are kept in ascending slot order, so that active sites are enumerated
            # exactly as in the reference's scattered layout (row order decides which rows a dropout mask hits) and
            # `unsort` restores instance order.  On the GPU the dropout stream differs from the reference's anyway and
            # every other result is independent of the site order, so the planes stay where they are (no gathers).
            order = sorted(range(n_i), key=lambda j: draw[j]) if not x.is_cuda else list(range(n_i))
            chosen = [draw[j] for j in order]
            if order != list(range(n_i)):
                masks, alphas, trans = (None if t is None else ops.take(t, 1, order) for t in (masks, alphas, trans))
                unsort = [order.index(j) for j in range(n_i)]
            slot_ids, n_slots = chosen, self.num_masks
        return x, masks, slot_ids, alphas, trans, chosen, n_slots, unsort, (b, n_f, n_i, h, w)

    def set_precision(self, precision="fp16"):
        """'fp16' (default): fp16 activations, fp32 accumulation - the width of the reference's autocast TRAINING
        (engine/train.py:227).  'high': the dense stage (encoder, ASPP, dense decoder, mask-guided attention, OS8 head) of
        an EVALUATION forward keeps fp32 activations and runs its contractions on the tensor cores with split-fp16
        operands (x = hi + lo; three MMAs per product), i.e. at fp32-level accuracy like the reference's fp32 evaluation
        (engine/test.py:131 has no autocast); alpha_os8 then matches the reference to ~1e-5.  The sparse refinement stage
        stays fp16.  Image model only; training is unaffected."""
        if precision not in ("fp16", "high"):
            raise ValueError(f"precision must be 'fp16' or 'high' (got {precision!r})")
        if precision == "high" and type(self) is not MaGGIe:
            raise NotImplementedError("precision='high' is implemented for the image model")
        self.precision = self.encoder.precision = precision
        return self

    def enable_cuda_graphs(self, on=True):
        """Replay the dense stage (encoder, ASPP, dense decoder blocks, attention) as CUDA graphs: forward AND backward in
        training mode, the forward alone under `torch.no_grad()` in evaluation mode; one graph (pair) per input shape.
        The sparse stage stays eager (its shapes follow the number of active sites).  The warm-up passes a capture needs
        do not count: the stage's running statistics and spectral-norm vectors are restored after the capture, so the
        first replay starts from the state an eager run would have started from."""
        self._use_graphs = bool(on)
        return self

    def _graph_flags(self):
        """What a captured graph bakes in besides the input shapes: the statistics-exchange setting (`set_sync_bn`, or
        BatchNorm containers converted to `nn.SyncBatchNorm` after a capture) and the stream / kernel-chain switches."""
        return (dense._SYNC_EPOCH, type(self.encoder.bn1).__name__, dense.SIDE_BRANCHES, dense.AUX_WGRAD, dense.FUSED_BN_APPLY)

    @staticmethod
    def _volatile_state(stage):
        """The tensors a forward of the stage updates in place: BatchNorm running statistics / counters, spectral-norm u, v."""
        keep = ("running_mean", "running_var", "num_batches_tracked", "weight_u", "weight_v")
        return {k: v for k, v in stage.state_dict(keep_vars=True).items() if k.endswith(keep)}

    def _eval_graph(self, stage, args):
        """Capture (once per shape / precision) and replay the evaluation forward of the dense stage."""
        key = ("eval", getattr(self.encoder, "precision", "fp16")) + self._graph_flags() + tuple((tuple(a.shape), a.dtype) for a in args)
        entry = self._graphs.get(key)
        if entry is None:
            from ... import _lib
            static_in = tuple(a.clone() for a in args)
            vol = self._volatile_state(stage)
            saved = {k: v.detach().clone() for k, v in vol.items()}
            cur = torch.cuda.current_stream(args[0].device)
            side = torch.cuda.Stream(device=args[0].device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(2):
                    stage(*static_in)
            cur.wait_stream(side)
            torch.cuda.synchronize(args[0].device)
            before = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = stage(*static_in)
            with torch.no_grad():
                for k, v in vol.items():
                    v.copy_(saved[k])
            entry = self._graphs[key] = (g, static_in, static_out, _lib.launch_count() - before)
        g, static_in, static_out, n = entry
        for dst, src in zip(static_in, args):
            dst.copy_(src)
        g.replay()
        self.replayed_native_launches += n
        return tuple(o.clone() if torch.is_tensor(o) else o for o in static_out) if isinstance(static_out, tuple) else static_out

    def _dense(self, x, masks, slot_ids, mask_os8, gt_os8, mem_feat=None):
        stage = self._stage[0]
        if stage.training != self.training:
            stage.train(self.training)
        ids = ops.slot_ids_tensor(slot_ids, x.device)
        args = (x.float().contiguous(), masks.contiguous(), ids, mask_os8.float(),
                gt_os8.float() if gt_os8 is not None else mask_os8.float())
        if mem_feat is not None:
            args = args + (mem_feat,)
        if self._use_graphs and not self.training and not torch.is_grad_enabled() and x.is_cuda:
            return self._eval_graph(stage, args)
        if not (self._use_graphs and self.training and torch.is_grad_enabled()) or dense.sync_bn_needs_eager(self, x.device):
            return stage(*args)      # (the collective fallback of the BatchNorm statistics exchange cannot be captured)
        key = self._graph_flags() + tuple((tuple(a.shape), a.dtype) for a in args)
        entry = self._graphs.get(key)
        if entry is None:
            from ... import _lib
            sample = tuple(a.clone() for a in args)
            vol = self._volatile_state(stage)
            saved = {k: v.detach().clone() for k, v in vol.items()}
            before = _lib.launch_count()
            fn = torch.cuda.make_graphed_callables(stage, sample, num_warmup_iters=3, allow_unused_input=True)
            # 3 warm-up iterations + 1 capture, each one forward + backward of the stage; their in-place updates of the
            # running statistics / power-iteration vectors are undone (the first batch must count once, not five times)
            with torch.no_grad():
                for k, v in vol.items():
                    v.copy_(saved[k])
            entry = self._graphs[key] = (fn, (_lib.launch_count() - before) // 4)
        self.replayed_native_launches += entry[1]
        return entry[0](*args)

    # hooks overridden by the video model
    def _extra_outputs(self, pred, output, n_i):
        pass

    def _extra_losses(self, pred, loss_dict, w4, w1, alphas, shape5, pad_ratio=1.0):
        pass

    def _extra_decoder_losses(self, pred, loss_dict):
        pass

    def _input_stage(self, batch):
        """Everything that depends on the inputs only: flattening, the pooled OS8 masks and - in warm-up iterations - the
        uncertain-region mask with its site tables (the step's one host read)."""
        prep = self._prepare(batch)
        x, masks, slot_ids, alphas, trans, chosen, n_slots, unsort, (b, n_f, n_i, h, w) = prep
        mask_os8, gt_os8 = self.decoder.pooled_masks(masks, alphas, b, n_f, n_i, h, w, self.training, chosen, n_slots)
        # the step's status word (site counts + device-side flags, ONE host read per step, see ops.new_status); the
        # "some sample has no mask at all" flag stands in for the reference's per-layer NaN check + ValueError
        # (module/mask_attention.py:95-98, five host synchronisations per forward there)
        status = ops.new_status(x.device) if x.is_cuda else None
        if status is not None:
            empty = ~mask_os8.flatten(1).any(1)                       # [b]: no mask pixel in any frame / slot of the sample
            status[ops.STATUS_EMPTY_MASK:ops.STATUS_EMPTY_MASK + 1].copy_(empty.any().to(torch.int32))
        roi_plan = self.decoder.plan_roi(alphas, batch.get("iter", 0), status)
        return prep, mask_os8, gt_os8, roi_plan, status

    def _input_stage_async(self, batch, ready_event):
        """`batch['ready_event']` (optional, a `torch.cuda.Event`): the caller produced the inputs on another stream (a
        copy stream, as prefetching loaders do) and they are complete once the event has fired.  The input stage then
        runs on a side stream that waits for that event only, NOT for the work already queued on the compute stream (the
        previous step's backward): its host read no longer drains the GPU queue, and the CPU keeps running ahead."""
        main = torch.cuda.current_stream()
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != main.device:
            self._side_stream = torch.cuda.Stream(device=main.device, priority=-1)
        side = self._side_stream
        side.wait_event(ready_event)
        with torch.cuda.stream(side):
            out = self._input_stage(batch)
        main.wait_event(ready_event)
        main.wait_stream(side)

        def keep(o):   # tensors allocated on the side stream are consumed on the compute stream
            if torch.is_tensor(o) and o.is_cuda:
                o.record_stream(main)
            elif isinstance(o, (tuple, list)):
                for v in o:
                    keep(v)
        keep(out)
        return out

    def forward(self, batch, **kwargs):
        ev = batch.get("ready_event")
        if ev is not None and batch["image"].is_cuda:
            prep, mask_os8, gt_os8, roi_plan, status = self._input_stage_async(batch, ev)
        else:
            prep, mask_os8, gt_os8, roi_plan, status = self._input_stage(batch)
        x, masks, slot_ids, alphas, trans, chosen, n_slots, unsort, (b, n_f, n_i, h, w) = prep
        mem_feat = kwargs.pop("mem_feat", None)
        it = batch.get("iter", 0)
        fea1, fea2, fea3, *dense_out = self._dense(x, masks, slot_ids, mask_os8, gt_os8,
                                                   mem_feat if torch.is_tensor(mem_feat) else None)
        if dense_out[1].dtype == torch.float32 and dense_out[1].is_cuda:
            # precision='high': the OS8 logits / tokens stay fp32; the sparse refinement stage below takes fp16 features
            h16 = lambda t: t.to(torch.float16).contiguous(memory_format=torch.channels_last)
            fea1, fea2, fea3, dense_out[1] = h16(fea1), h16(fea2), h16(fea3), h16(dense_out[1])
        with ops.step_scope("sparse_stage", x.device):
            pred = self.decoder(tuple(dense_out), (fea1, fea2, fea3), (h, w), b=b, n_f=n_f, n_i=n_i, masks=masks,
                                iter=it, gt_alphas=alphas, spar_gt=trans, slots=chosen, n_slots=n_slots,
                                roi_plan=roi_plan, status=status, **kwargs)
        self.last_site_counts = pred.pop("site_counts", None)

        alpha_pred = pred.pop("refined_masks")
        w4 = w1 = pred["detail_mask"].to(alpha_pred.dtype)
        if self.training and np.random.rand() < 0.75:
            w4, w1 = pred.pop("weight_os4"), pred.pop("weight_os1")
        view = lambda t: t[:, :n_i].reshape(b, n_f, n_i, h, w)
        output = {k: view(pred[k]) for k in ("alpha_os1", "alpha_os4", "alpha_os8")}
        output["refined_masks"] = view(alpha_pred)
        output["detail_mask"] = view(pred["detail_mask"])
        self._extra_outputs(pred, output, n_i)

        if self.training:
            valid = (trans.sum((2, 3), keepdim=True) > 0).float()
            if type(self)._extra_losses is MaGGIe._extra_losses:
                # image model: the only consumer of `pred * valid` is the loss, which applies the factor itself
                loss_dict = losses.compute_loss(pred, w4, w1, alphas, self.cfg, valid=valid)
            else:
                for k in list(pred):
                    if "loss" in k or "mem_" in k:
                        continue
                    pred[k] = pred[k] * valid
                loss_dict = losses.compute_loss(pred, w4, w1, alphas, self.cfg)
            # planes per reference plane count (the reference sums an epsilon over its zero-padded slots too)
            self._extra_losses(pred, loss_dict, w4, w1, alphas, (b, n_f, n_i, h, w), n_slots / n_i)
            if "loss_max_atten" in pred and self.cfg.loss_atten_w > 0:
                loss_dict["loss_max_atten"] = pred["loss_max_atten"]
                loss_dict["total"] = loss_dict["total"] + loss_dict["loss_max_atten"] * self.cfg.loss_atten_w
            self._extra_decoder_losses(pred, loss_dict)
            if unsort is not None:
                output = {k: ops.take(v, 2, unsort) for k, v in output.items()}
            if x.is_cuda and torch.is_grad_enabled():
                dense.arm_backward_pool(x.device)   # one memset for the backward's small zeroed accumulators
            return output, loss_dict

        for k in pred:
            if k.startswith("mem_"):
                output[k] = pred[k]
        return output


class MaGGIe_Temp(MaGGIe):
    """Video model: temporal outputs / losses and the eval-time alpha-matte level aggregation over a 3-frame window.
    Reference: arch/maggie_temp.py:5-77, arch/maggie.py:348-365 (dtSSD)."""

    def _extra_outputs(self, pred, output, n_i):
        db, df, ta = pred.pop("diff_backward", None), pred.pop("diff_forward", None), pred.pop("temp_alpha", None)
        if db is not None:
            output["diff_pred_backward"] = db.repeat(1, 1, n_i, 1, 1)
            output["diff_pred_forward"] = df.repeat(1, 1, n_i, 1, 1)
            output["temp_alpha"] = ta

    def _extra_losses(self, pred, L, w4, w1, alphas, shape5, pad_ratio=1.0):
        if self.cfg.loss_dtSSD_w <= 0:
            return
        r = lambda t: t.reshape(*shape5).float()
        a8 = pred["alpha_os8"]
        w8 = (alphas.sum((2, 3), keepdim=True) > 0).to(a8.dtype).expand_as(a8)
        if self.cfg.loss_reweight_os8:
            lo, hi = 1.0 / 255.0, 254.0 / 255.0
            w8 = (((alphas <= hi) & (alphas >= lo)) | ((a8 <= hi) & (a8 >= lo))).to(a8.dtype) + w8
        dt = lambda p, g, m: MaGGIeTempDecoder._loss_dtssd(p, g, m, pad_ratio)
        d1 = dt(r(pred["alpha_os1"]), r(alphas), r(w1))
        d4 = dt(r(pred["alpha_os4"]), r(alphas), r(w4))
        d8 = dt(r(a8), r(alphas), r(w8))
        L.update(loss_dtSSD_os1=d1, loss_dtSSD_os4=d4, loss_dtSSD_os8=d8, loss_dtSSD=d1 * 2 + d4 + d8)
        L["total"] = L["total"] + L["loss_dtSSD"] * self.cfg.loss_dtSSD_w

    def _extra_decoder_losses(self, pred, L):
        if "loss_temp" in pred:
            L.update(loss_temp_bce=pred["loss_temp_bce"], loss_temp=pred["loss_temp"], loss_temp_dtssd=pred["loss_temp_dtssd"])
            L["total"] = L["total"] + pred["loss_temp"]

    def forward(self, batch, **kwargs):
        prev_pred = kwargs.pop("prev_pred", None)
        output = super().forward(batch, **kwargs)
        if self.training:
            return output
        al = output["refined_masks"]                                   # [1, 3, n_i, H, W]
        prev = al[:, 0] if prev_pred is None else prev_pred.to(al.device)
        nxt = al[:, -1]
        dfw = (output["diff_pred_forward"] > 0.5).float()
        dbw = (output["diff_pred_backward"] > 0.5).float()
        p01 = prev * (1 - dfw[:, 1]) + al[:, 1] * dfw[:, 1]             # propagate t-1 -> t
        p21 = nxt * (1 - dbw[:, 1]) + al[:, 1] * dbw[:, 1]              # propagate t+1 -> t
        p01 = torch.where((p01 - p21).abs() > 0.0, al[:, 1], p01)       # disagreement -> the model's own frame t
        al[:, 1] = p01
        al[:, 2] = p01 * (1 - dfw[:, 2]) + nxt * dfw[:, 2]
        return output
