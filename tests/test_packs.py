"""Host logic of the grouped operand packs (maggie_b200/packs.py): the gather indices reproduce, bit for bit, the
per-layer pack functions they replace (reshape / flip / permute / zero padding / fp16 cast)."""
import torch

from maggie_b200 import packs, sparse


def test_grouped_packs_equal_per_layer_packs():
    torch.manual_seed(0)
    in_proj = torch.randn(384, 128)
    ws = [torch.randn(32, 3, 3, 32), torch.randn(1, 3, 3, 32), torch.randn(64, 128), torch.randn(64, 3, 3, 64),
          torch.randn(32, 64), in_proj[:128], in_proj[128:256], in_proj[256:]]
    ps = packs.PackSet().prepare(ws)
    for w in ws:
        e = ps.lookup(w)
        assert e is not None
        assert torch.equal(ps.pack_fwd(e), sparse.pack_fwd(w)), tuple(w.shape)
        for mirror in (False, True):
            got, cop = ps.pack_bwd(e, mirror)
            ref, cop_ref = sparse.pack_bwd(w, mirror)
            assert cop == cop_ref and torch.equal(got, ref), (tuple(w.shape), mirror)
    # values are refreshed in place on the next prepare; the layout tables are reused
    idx = ps.idx_fwd
    ws[0].mul_(2.0)
    ps.prepare(ws)
    assert ps.idx_fwd is idx and torch.equal(ps.pack_fwd(ps.lookup(ws[0])), sparse.pack_fwd(ws[0]))
    assert ps.lookup(torch.randn(8, 8)) is None
