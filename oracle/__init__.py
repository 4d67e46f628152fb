"""CPU oracle for the CoAlign hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it, and only as the checker
or as the timed CPU baseline - never on the CUDA product path (``coalign_b200`` raises if its
CUDA library is missing; it has no CPU fallback).

Pinning status
--------------
* A3..A13 (PFN -> scatter -> ResNet BEV encoder -> warp + attention fusion -> decoder -> shrink
  -> heads): PINNED against outputs of the unmodified reference generated in the build container
  by ``tests/golden/gen_golden.py`` (fixtures in ``tests/golden/*.npz``).
* A1/A2 (spconv voxel generator): **parity unpinned** - spconv is a third-party dependency that
  is absent from /root/reference and from this image (docs pin v1.2.1,
  /root/reference/docs/md_files/installation.md:45-47).  ``voxelize.c`` / ``voxelize_np.py``
  restate its published ``points_to_voxel`` algorithm as consumed by
  /root/reference/opencood/data_utils/pre_processor/sp_voxel_preprocessor.py:62-85,145-174.
"""
