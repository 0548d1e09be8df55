"""papc_b200 -- Blackwell-native (sm_100a) drop-in for ONE hot path of AgentMaker/PAPC:
the PointNet++ SetAbstraction forward and the PointPillars pillar encode.

  papc_b200.layers    mirror of PAPC/models/layers/pointnet2_basic_layers.py (A1-A8)
  papc_b200.pillars   mirror of pointpillars point_cloud_ops / voxel_generator / pillars (A9-A11)
  papc_b200.dist      batch sharding over one process per GPU + the single all-gather
  papc_b200.synth     seeded synthetic inputs (host, NumPy)
  papc_b200.csrc/     hand-written CUDA kernels + the C ABI declared in include/papc_b200.h

The CUDA library is the only back end: there is no CPU fallback (oracle/ is test
infrastructure and is never imported from here).
"""
__version__ = "0.1.0"
