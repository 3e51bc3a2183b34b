"""Oracle = TEST INFRASTRUCTURE ONLY.

CPU restatements of the MaGGIe hot path (reference: hmchuong/MaGGIe, `maggie/network`), used as the
checker in `tests/`, in `__graft_entry__.smoke()` and as the `cpu_baseline` / `--impl reference` arm of
`bench.py`.  Nothing under `maggie_b200/` may import this package.  (The deterministic input / weight generator is the
repo-root `synthdata.py`, re-exported as `oracle/synth.py`; the GPU arm of `bench.py` and the scripts under `tools/` import
`synthdata` directly, so nothing on the measured GPU path imports this package.)

Pinning status: the reference ships no tests or golden vectors, so the oracle is pinned
against outputs of the *unmodified reference itself*, imported in the build container from
`/root/reference` through `oracle/ref_shims.py` (stubs for yacs/kornia/fvcore + the pure-torch
`spconv.pytorch` restatement in `oracle/spconv_torch.py`).  `oracle/make_golden.py` is the generating
script; the vectors live in `tests/golden/`.  The sparse-conv arithmetic itself lives in the third-party
`spconv-cu120` wheel (requirements.txt:4, unpinned, absent here) - for that boundary parity is
"unpinned" beyond the restated published semantics and the fully-active identities checked in
`tests/test_oracle_spconv.py`.
"""
