"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatements ("ports") of the NeuralRecon-W per-ray training hot path, used
as the parity checker for the CUDA library in ``neuralrecon-w_b200/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import anything from this package.  The product
path (``neuralrecon-w_b200/nrw``) never imports it and fails loudly when the
CUDA extension is missing.

Pinning status
--------------
* MLP / sampler / compositing path (``neuconw_port``): pinned against the
  UNMODIFIED reference modules imported from ``/root/reference`` in the build
  container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``; live
  comparison in ``tests/test_oracle_vs_reference.py`` when the tree is present).
* Octree path (``octree_port``): the reference delegates to NVIDIA Kaolin
  (unpinned fork ``git+https://github.com/Burningdust21/kaolin.git``,
  environment.yaml:18) whose source is not under ``/root/reference`` and which
  is not installable here.  **Parity unpinned** for that boundary: the port
  restates Kaolin's documented SPC semantics and the reference's own call
  sites (tools/prepare_data/generate_voxel.py:311-439).
"""
