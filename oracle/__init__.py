"""CPU oracle for the LiDOG hot path (TEST INFRASTRUCTURE, never shipped).

This package is a plain numpy / torch-CPU restatement of what the reference's
hot path computes (voxelisation -> coordinate maps -> kernel maps -> sparse
convolution -> BEV projection).  It exists only so that `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs can check and time the CUDA product against it.  Nothing under
`lidog_b200/` or `MinkowskiEngine/` may import it.

PARITY STATUS
  * (a) voxelisation, (b) kernel maps, (c) sparse convolution: **parity
    unpinned**.  Their arithmetic lives in MinkowskiEngine==0.5.4
    (reference README.md:29), which is neither vendored under /root/reference
    nor installed nor fetchable, and the reference has no tests / golden
    vectors.  The conventions restated here are the ones SURVEY.md Appendix C
    fixes from the reference call sites; they are additionally anchored on
    ME-independent ground truths (dense torch conv3d on a densified grid,
    set/round-trip properties, fp64 gradcheck) in tests/.
  * (d) BEV projection: **pinned** against the reference's own
    `MinkUNetBaseBEV.sparse2super` (utils/models/minkunet_bev.py:169-230),
    executed on CPU by tests/golden/make_bev_golden.py; the vectors are
    committed under tests/golden/.
"""
