"""Example: a few training steps of MaGGIe through maggie_b200 on one GPU (or one process per GPU under torchrun), using every
piece the package offers around the model: uint8 input stage (K16), dense-stage CUDA graphs, flat-gradient all-reduce, optional
SyncBatchNorm-equivalent statistics (K15), fused unscale + clip + AdamW (K14), evaluation tail.

    python examples/train_steps.py                      # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_steps.py --sync-bn

The inputs are synthetic soft ellipses (there is no dataset in this repository); with real data the dataset stops before
`ToTensor()` and hands the decoded uint8 arrays to `io.prepare_batch`."""
import argparse
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maggie_b200 import io                                      # noqa: E402
from maggie_b200.config import CfgNode                          # noqa: E402
from maggie_b200.dp import FlatGradAllReduce, set_sync_bn       # noqa: E402
from maggie_b200.network import build_model                     # noqa: E402
from maggie_b200.optim import FusedAdamW                        # noqa: E402

MODEL = dict(  # the `model` section of the reference's configs/maggie_image.yaml
    arch="MaGGIe", weights="", sync_bn=False, having_unused_params=True, warmup_iters=3000,
    encoder="res_shortcut_embed_29", encoder_args=dict(num_embed=3, num_mask=10, pretrained=True),
    aspp=dict(in_channels=512, out_channels=512), decoder="res_shortcut_inst_matt_spconv_22",
    decoder_args=dict(atten_block=2, atten_dim=128, atten_head=1, atten_stride=1, detail_mask_dropout=0.1, final_channel=64,
                      freeze_detail_branch=False, head_channel=120, max_inst=10, use_id_pe=True, warmup_detail_iter=3000,
                      warmup_mask_atten_iter=0),
    loss_alpha_w=1.0, loss_alpha_type="l1", loss_alpha_grad_w=0.05, loss_alpha_lap_w=0.05, loss_atten_w=5.0,
    loss_reweight_os8=True, loss_dtSSD_w=0.0)


def synthetic_uint8(b, n_i, H, W, seed):
    """What a dataset would decode: frames uint8 [b,1,H,W,3], alphas / masks uint8 [b,1,n_i,H,W] (soft ellipses)."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.randint(0, 256, (b, 1, H, W, 3), generator=g, dtype=torch.uint8)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    alphas = torch.zeros((b, 1, n_i, H, W))
    for k in range(b):
        for i in range(n_i):
            cy, cx = H * (0.35 + 0.3 * torch.rand(1, generator=g)), W * (i + 0.5) / n_i
            ry, rx = H * 0.22, W * 0.35 / n_i
            rho = torch.sqrt(((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2)
            alphas[k, 0, i] = torch.clamp(-(rho - 1.0) * min(ry, rx) / 6.0 + 0.5, 0.0, 1.0)
    a8 = (alphas * 255.0).round().to(torch.uint8)
    return frames, a8, ((alphas > 0.5).to(torch.uint8) * 255)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--frames", type=int, default=4, help="frames per GPU")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--sync-bn", action="store_true")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        if args.sync_bn:
            set_sync_bn(True)                    # or: model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
    torch.manual_seed(1234)                      # identical initial weights on every rank
    model, _ = build_model(CfgNode(MODEL))
    model.to(dev).train()
    model.enable_cuda_graphs(True)
    flat = FlatGradAllReduce(model.parameters(), bank=model.bank)
    opt = FusedAdamW(flat, lr=1.5e-4, betas=(0.5, 0.999), weight_decay=0.01, clip_norm=0.01)   # engine/optim.py, train.py:270
    scale = 128.0                                # a GradScaler's current scale works the same way
    for it in range(1, args.steps + 1):
        frames, alphas, masks = synthetic_uint8(args.frames, 3, args.size, args.size, seed=1000 * rank + it)
        batch = io.prepare_batch(frames.to(dev), alphas.to(dev), masks.to(dev))          # image / alpha / mask (1/8 size)
        batch["transition"] = ((batch["alpha"] > 0) & (batch["alpha"] < 1)).float()      # the loader's gen_transition_gt
        batch["iter"] = it
        flat.zero()
        out, loss = model(batch, mem_feat=None)
        (loss["total"] * scale).backward()
        flat.allreduce()
        norm, found_inf = opt.step(grad_scale=scale).tolist()                            # (a host read, for the log line)
        if rank == 0:
            print(f"iter {it}: loss {float(loss['total']):.4f}  grad norm {norm:.3f}  skipped {bool(found_inf)}")
    model.eval()
    with torch.no_grad():
        frames, alphas, masks = synthetic_uint8(1, 3, args.size, args.size, seed=7)
        pred = model(io.prepare_batch(frames.to(dev), None, masks.to(dev)), mem_feat=None)
    alpha = io.finalize_alpha(pred["refined_masks"], [{"name": "resize", "ori_size": (args.size + 40, args.size + 24)}])
    if rank == 0:
        print("eval alpha", tuple(alpha.shape), "in", (float(alpha.min()), float(alpha.max())))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
