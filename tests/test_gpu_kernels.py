"""-m gpu: the hand-written kernels, called through the C ABI, against the oracle on the same seeded inputs.
Integer / bit work is compared bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ops_ref
from oracle import synth
from oracle import unknown as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def _alphas(shape, seed):
    rng = np.random.RandomState(seed)
    a = rng.rand(*shape).astype(np.float32)
    a[a < 0.6] = 0.0                      # mostly background
    a[a > 0.97] = 1.0                     # some saturated foreground
    flat = a.reshape(-1)
    flat[rng.randint(0, flat.size, 16)] = np.float32(1.0 / 255.0)    # exactly on the thresholds
    flat[rng.randint(0, flat.size, 16)] = np.float32(254.0 / 255.0)
    return a


@pytest.mark.parametrize("shape", [(3, 64, 64), (2, 2, 96, 160), (5, 40, 72), (1, 200, 33), (4, 8, 8), (1, 130, 1056)])
def test_unknown_mask_bit_exact(dev, shape):
    from maggie_b200 import ops
    a = _alphas(shape, seed=sum(shape))
    n = int(np.prod(shape[:-2]))
    rng = np.random.RandomState(1)
    for widths in ([15] * n, [13] * n, [7] * n, list(rng.randint(1, 30, n)), [1] * n, [29] * n, [2] * n):
        got = ops.unknown_mask(torch.from_numpy(a).to(dev), widths).cpu().numpy()
        ref = U.compute_unknown(a, widths)
        assert got.dtype == np.uint8 and (got == ref).all(), (shape, widths[:4], int((got != ref).sum()))


def test_unknown_mask_and_mask_and_soft_ellipses(dev):
    from maggie_b200 import ops
    al = synth.soft_ellipse_alphas(2, 3, 256, 320, edge_px=5.0).numpy()
    roi = U.compute_unknown(al, [15] * 6)
    got = ops.unknown_mask(torch.from_numpy(al).to(dev), [13] * 6, and_mask=torch.from_numpy(roi).to(dev)).cpu().numpy()
    assert (got == U.compute_unknown(al, [13] * 6) * roi).all()
    # empty input and all-inactive input
    assert ops.unknown_mask(torch.zeros(0, 16, 16, device=dev), []).shape == (0, 16, 16)
    assert int(ops.unknown_mask(torch.ones(2, 64, 64, device=dev), [15, 15]).sum()) == 0


def test_unknown_mask_properties_at_full_size(dev):
    """C2-size (80 x 512 x 512): dilation is extensive, monotone in the kernel size and idempotent on a full mask."""
    from maggie_b200 import ops
    al = torch.stack([synth.soft_ellipse_alphas(1, 10, 512, 512, edge_px=6.0, seed=s)[0] for s in range(8)]).to(dev)
    raw = ((al > 1 / 255) & (al < 254 / 255)).to(torch.uint8)
    u7, u15 = ops.unknown_mask(al, [7] * 80), ops.unknown_mask(al, [15] * 80)
    assert bool((u7 >= raw).all()) and bool((u15 >= u7).all())
    assert bool((ops.unknown_mask(al, [1] * 80) == raw).all())
    frac = float(u15.float().mean())
    assert 0.01 < frac < 0.5
    # checksum against the oracle on two slices only (the oracle takes ~0.1 s per slice at this size)
    ref = U.compute_unknown(al[3, 4:6].cpu().numpy(), [15, 15])
    assert (u15[3, 4:6].cpu().numpy() == ref).all()


@pytest.mark.parametrize("shape,p", [((2, 64, 64), 0.02), ((3, 96, 160), 0.2), ((1, 8, 8), 0.5), ((6, 128, 72), 0.001), ((2, 64, 64), 0.0)])
def test_site_tables_bit_exact(dev, shape, p):
    from maggie_b200 import ops
    rng = np.random.RandomState(int(p * 1000) + shape[0])
    roi = (rng.rand(*shape) < p).astype(np.uint8)
    if p == 0.2:
        roi[:, :, -1] = 1
        roi[:, 0, :] = 1
    T = ops.build_sites(torch.from_numpy(roi).to(dev))
    coords, nbr, parent, child, shapes = ops_ref.sites_tables_ref(roi)
    assert T.counts == [len(c) for c in coords]
    for l in range(4):
        assert (T.coords[l].cpu().numpy() == coords[l]).all(), f"coords level {l}"
        for name, got, ref in (("nbr", T.nbr[l], nbr[l]), ("parent", T.parent[l], parent[l]), ("child", T.child[l], child[l])):
            if ref is not None:
                assert (got.cpu().numpy() == ref).all(), f"{name} level {l}"


def test_site_tables_on_dilated_band_full_size(dev):
    from maggie_b200 import ops
    al = synth.soft_ellipse_alphas(1, 3, 512, 512, edge_px=6.0).numpy()[0]
    roi = U.compute_unknown(al, [15] * 3)
    T = ops.build_sites(torch.from_numpy(roi).to(dev))
    coords, nbr, parent, child, _ = ops_ref.sites_tables_ref(roi)
    assert T.counts == [len(c) for c in coords] and T.counts[0] > 10000
    assert (T.nbr[0].cpu().numpy() == nbr[0]).all() and (T.parent[0].cpu().numpy() == parent[0]).all()
    assert (T.child[3].cpu().numpy() == child[3]).all()
    # structural properties: lexicographic order, centre tap is the site itself, every site has a parent
    c = T.coords[0].cpu().numpy().astype(np.int64)
    key = (c[:, 0] * 512 + c[:, 1]) * 512 + c[:, 2]
    assert (np.diff(key) > 0).all()
    assert (T.nbr[0][:, 4].cpu().numpy() == np.arange(T.counts[0])).all()
    assert ((T.parent[0].cpu().numpy() >= 0).sum(1) >= 1).all()


@pytest.mark.parametrize("C", [6, 8, 16, 32])
def test_mask_embed_fwd_bwd(dev, C):
    from maggie_b200 import ops
    torch.manual_seed(0)
    B, M, H, W = 2, 3, 40, 56
    image = torch.randn(B, 3, H, W)
    masks = (torch.rand(B, M, H, W) > 0.6).float()
    table = torch.randn(11, 3, requires_grad=True)
    slot_ids = [7, 0, 4]
    ref = ops_ref.mask_embed(image, masks, table, slot_ids, C).float()
    gout = torch.randn(B, C, H, W)
    (ref * gout).sum().backward()
    tab_d = table.detach().to(dev).requires_grad_(True)
    got = ops.mask_embed(image.to(dev), masks.to(dev), tab_d, slot_ids, C)
    assert got.shape == (B, C, H, W) and got.dtype == torch.float16
    assert got.permute(0, 2, 3, 1).is_contiguous()
    assert (got.float().cpu() - ref.detach()).abs().max() < 2e-3 * max(1.0, float(ref.abs().max()))
    (got.float() * gout.to(dev).half().float()).sum().backward()
    assert (tab_d.grad.cpu() - table.grad).abs().max() < 2e-2 * float(table.grad.abs().max())
    assert float(tab_d.grad[[2, 3, 4, 6, 7, 9, 10]].abs().max()) == 0.0   # unused slots get no gradient


@pytest.mark.parametrize("planes,h,w,S,with_scale", [((8, 3), 64, 64, 8, True), ((24, 1), 128, 128, 4, False), ((2, 3), 40, 24, 1, True),
                                                     ((1, 2), 9, 7, 2, False), ((3,), 5, 6, 8, True)])
def test_upsample_tanh_fwd_bwd_matches_torch(dev, planes, h, w, S, with_scale):
    """K7 against F.interpolate(bilinear, align_corners=False) + (tanh + 1) / 2 (* per-plane factor) and its autograd."""
    from maggie_b200 import ops
    torch.manual_seed(h + S)
    x = (torch.randn(*planes, h, w, device=dev) * 2).requires_grad_(True)
    x.data[..., : h // 2, :] -= 99.0 * (torch.rand(*planes, h // 2, w, device=dev) > 0.7)   # the -99 fill of the logit maps
    ps = (torch.rand(*planes, device=dev) > 0.3).float() if with_scale else None
    y = ops.upsample_tanh(x, scale=float(S) if S > 1 else None, plane_scale=ps)
    g = torch.randn_like(y)
    y.backward(g)
    got = x.grad.clone()
    x2 = x.detach().clone().requires_grad_(True)
    xi = x2.reshape(-1, 1, h, w)
    up = F.interpolate(xi, scale_factor=float(S), mode="bilinear", align_corners=False) if S > 1 else xi
    ref = ((torch.tanh(up) + 1.0) / 2.0).reshape(y.shape)
    if ps is not None:
        ref = ref * ps.reshape(*planes, 1, 1)
    ref.backward(g)
    assert y.shape == ref.shape and (y - ref).abs().max() < 1e-6
    assert (got - x2.grad).abs().max() < 1e-5 * max(1.0, float(x2.grad.abs().max()))


def test_unknown_mask_device_side_source_select():
    """`use_alt` (a device flag) switches the source to `alt` without a host read (the reference's warm-up switch)."""
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(21)
    a = torch.rand(3, 2, 96, 128, generator=g).cuda()
    b = torch.rand(3, 2, 96, 128, generator=g).cuda()
    widths = [5, 9, 15, 3, 29, 1]
    for flag in (0, 1):
        f = torch.tensor([flag], dtype=torch.int32, device="cuda")
        got = ops.unknown_mask(a, widths, alt=b, use_alt=f)
        want = ops.unknown_mask(b if flag else a, widths)
        assert torch.equal(got, want)


def test_upsample_tanh_all_zero_flag():
    from maggie_b200 import ops
    x = torch.full((2, 3, 8, 8), -40.0, device="cuda")          # tanh == -1 exactly: alpha == 0 everywhere
    flag = torch.ones(1, dtype=torch.int32, device="cuda")
    ops.upsample_tanh(x, scale=8.0, all_zero=flag)
    assert int(flag) == 1
    x[1, 2, 3, 3] = 0.5
    ops.upsample_tanh(x, scale=8.0, all_zero=flag)
    assert int(flag) == 0
    # a zero plane factor keeps the flag set whatever the logits are
    flag.fill_(1)
    ops.upsample_tanh(x, scale=8.0, plane_scale=torch.zeros(2, 3, device="cuda"), all_zero=flag)
    assert int(flag) == 1


def test_fuse_stage_bit_exact_against_mask_then_blend():
    """K10: mask + blend of one progressive-fusion stage in one kernel == unknown_mask followed by torch.where."""
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(33)
    for shape in ((4, 3, 128, 160), (2, 2, 72, 100)):              # vectorised (W % 32 == 0) and scalar tails
        src = torch.rand(*shape, generator=g).cuda()
        src[src < 0.5] = 0.0
        fine, roi = torch.rand(*shape, generator=g).cuda(), (torch.rand(*shape, generator=g) > 0.3).to(torch.uint8).cuda()
        widths = [int(v) for v in torch.randint(1, 28, (shape[0] * shape[1],), generator=g)]
        a, w = ops.fuse_stage(src, fine, src, widths, roi)
        w_ref = ops.unknown_mask(src, widths, and_mask=roi)
        assert torch.equal(w, w_ref)
        assert torch.equal(a, torch.where(w_ref != 0, fine, src))
