"""Deterministic synthetic inputs and weights shared by the oracle, the golden generator, the tests, `bench.py` and the
developer scripts under `tools/`.  A data generator only: no model code, no oracle code (`oracle/synth.py` re-exports it).

Inputs follow SURVEY.md §8(d): N(0,1) images, soft-ellipse
alphas whose edge width `e` (px) is the active-fraction knob, mask = alpha > 0.5, transition = 0<alpha<1.
Weights are a pure function of (tensor name, shape, seed) so the same values can be loaded into the
unmodified reference (in the build container, to make goldens) and into the product model (on the GPU box,
where the reference does not exist).
"""
import math
import zlib

import numpy as np
import torch

MODEL_CFG = dict(
    arch="MaGGIe", weights="", sync_bn=False, having_unused_params=True, warmup_iters=3000,
    encoder="res_shortcut_embed_29",
    encoder_args=dict(num_embed=3, num_mask=10, pretrained=True),
    aspp=dict(in_channels=512, out_channels=512),
    decoder="res_shortcut_inst_matt_spconv_22",
    decoder_args=dict(atten_block=2, atten_dim=128, atten_head=1, atten_stride=1, detail_mask_dropout=0.1,
                      final_channel=64, freeze_detail_branch=False, head_channel=120, max_inst=10,
                      use_id_pe=True, warmup_detail_iter=3000, warmup_mask_atten_iter=0),
    loss_alpha_w=1.0, loss_alpha_type="l1", loss_alpha_grad_w=0.05, loss_alpha_lap_w=0.05,
    loss_atten_w=5.0, loss_reweight_os8=True, loss_dtSSD_w=0.0,
)  # configs/maggie_image.yaml:30-70


def video_cfg(**over):
    """configs/maggie_video.yaml:34-62."""
    cfg = model_cfg(arch="MaGGIe_Temp", decoder="res_shortcut_inst_matt_spconv_temp_22", loss_dtSSD_w=1.0, **over)
    cfg["decoder_args"]["temp_method"] = "bi_fusion"
    return cfg


def model_cfg(**over):
    import copy

    cfg = copy.deepcopy(MODEL_CFG)
    cfg.update(over)
    return cfg


def soft_ellipse_alphas(n_frames, n_inst, H, W, edge_px=6.0, shift=(3, 2), seed=1234):
    """[n_frames, n_inst, H, W] float32 alphas in [0,1]; instance centres spread over the frame,
    translated by `shift` px per frame (video)."""
    rng = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    out = np.zeros((n_frames, n_inst, H, W), np.float32)
    for i in range(n_inst):
        cy = H * (0.3 + 0.4 * rng.rand())
        cx = W * ((i + 0.5) / n_inst) * 0.8 + 0.1 * W
        ry = H * (0.18 + 0.12 * rng.rand())
        rx = W * (0.10 + 0.08 * rng.rand()) / max(1.0, n_inst / 3.0) * 1.5
        for f in range(n_frames):
            dy, dx = (yy - cy - shift[0] * f) / ry, (xx - cx - shift[1] * f) / rx
            rho = np.sqrt(dy * dy + dx * dx)
            # signed distance approximation in pixels along the radial direction
            dist = (rho - 1.0) * min(ry, rx)
            out[f, i] = np.clip(-dist / edge_px + 0.5, 0.0, 1.0)
    return torch.from_numpy(out)


def make_batch(b, n_f, n_i, H, W, edge_px=6.0, seed=1234, train=False, it=1, mask_os8=False, binary_alpha=False):
    """Batch dict in the reference's input contract (maggie/network/arch/maggie.py:63-78)."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(b, n_f, 3, H, W, generator=g)
    alphas = torch.stack([soft_ellipse_alphas(n_f, n_i, H, W, edge_px, seed=seed + 17 * k) for k in range(b)])
    if binary_alpha:  # no uncertain pixel at all (degenerate batch)
        alphas = (alphas > 0.5).float()
    mask = (alphas > 0.5).float()
    if mask_os8:
        mask = mask[..., ::8, ::8].contiguous()  # nearest downsample as dataloader/him.py:175-176
    batch = dict(image=image, mask=mask)
    if train:
        batch.update(alpha=alphas, transition=((alphas > 0) & (alphas < 1)).float(), iter=it,
                     fg=torch.zeros(b, n_f, 3, H, W), bg=torch.zeros(b, n_f, 3, H, W))
    return batch


def _uniform(shape, a, g):
    return (torch.rand(shape, generator=g) * 2 - 1) * a


def synth_tensor(name, ref, seed=1234):
    """Value for state-dict entry `name` shaped/typed like `ref`."""
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7FFFFFFF)
    shape = tuple(ref.shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=ref.dtype)
    if leaf in ("weight_u", "weight_v"):
        t = torch.randn(shape, generator=g)
        return t / (t.norm() + 1e-12)
    if leaf == "running_mean":
        return _uniform(shape, 0.2, g)
    if leaf == "running_var":
        return torch.rand(shape, generator=g) + 0.5
    if "mask_embed_layer" in name:
        return torch.randn(shape, generator=g)
    if len(shape) >= 2:
        if len(shape) == 4 and "decoder." in name and name.split(".")[1] in (
                "dummy_downscale", "layer3", "guidance_layer", "layer3_smooth", "refine_OS4", "layer4",
                "layer4_smooth", "layer5", "layer5_smooth", "refine_OS1"):
            fan_out, fan_in = shape[0] * shape[1] * shape[2], shape[3] * shape[1] * shape[2]  # [Cout,kh,kw,Cin]
        else:
            rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
            fan_out, fan_in = shape[0] * rf, shape[1] * rf
        return _uniform(shape, math.sqrt(6.0 / (fan_in + fan_out)), g)
    # 1-D: a 1-D `weight` is always a norm scale (BN / BN1d / LN); everything else is a bias
    if leaf == "weight":
        return torch.rand(shape, generator=g) + 0.5
    return _uniform(shape, 0.1, g)


def synth_state_dict(template, seed=1234):
    """template: mapping name -> tensor (e.g. model.state_dict()). Returns new tensors, same dtypes."""
    return {k: synth_tensor(k, v, seed).to(v.dtype) for k, v in template.items()}
