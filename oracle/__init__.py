"""CPU oracle for the PAPC hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  ``papc_b200`` never
imports it and has no CPU fallback.

  layers_np.py    line-by-line NumPy restatement of pointnet2_basic_layers.py:17-281
  pillars_np.py   line-by-line NumPy restatement of point_cloud_ops.py / pillars.py
  papc_oracle.c   arithmetic-pinned C restatement of FPS / ball query / 3-NN / voxeliser
  capi.py         ctypes bindings of papc_oracle.c (built by oracle/build.py)
  ref_voxel.py    loader of the reference's own numba voxeliser (build container only)
"""
