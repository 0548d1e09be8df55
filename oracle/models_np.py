"""NumPy restatement of the reference's four PointNet++ model definitions -- TEST INFRASTRUCTURE,
NOT PRODUCT (only tests/ may import it).

  PointNet2_SSG_Clas / PointNet2_MSG_Clas   PAPC/models/classify/pointnet2/pointnet2.py:6-41, :43-78
  PointNet2_SSG_Seg  / PointNet2_MSG_Seg    PAPC/models/segment/pointnet2/pointnet2.py:6-51, :53-98
  Categorical                               PAPC/models/layers/pointnet2_basic_layers.py:7-14

built on the layer restatements of ``oracle/layers_np.py``.  The WIRING is pinned: the reference's own
model classes, executed unmodified over a NumPy stand-in for paddle (tests/golden/make_golden_models.py
-> tests/golden/models_ref.npz), give the same logits and head running statistics as these classes with
identical parameters (tests/test_oracle_vs_reference_source.py).  The arithmetic inside Paddle's layers
stays PARITY UNPINNED: the reference executes inside PaddlePaddle (absent here) and ships no vectors.

Paddle semantics restated for the heads: ``nn.Linear`` is ``x @ W + b`` with ``W`` [in,out];
``nn.BatchNorm1D`` (registered, so it follows train()/eval()): eps 1e-5, momentum 0.9, biased batch
variance in training, running statistics in eval; ``nn.Dropout`` (mode upscale_in_train) is the
identity in eval -- the models are compared in eval mode and, for the normalisation, in training mode
with the dropout probability set to 0 (Paddle's dropout mask is not reproducible outside Paddle).
"""
from __future__ import annotations

import numpy as np

from . import layers_np as LN

F32 = np.float32


def Categorical(y, num_class=16):
    """layers.py:7-14: y [B,1] int -> [B,num_class,1] float32."""
    new_y = np.eye(num_class)[np.asarray(y).reshape(-1, 1),]            # :11  [B,1,num_class]
    return new_y.transpose(0, 2, 1).astype(F32)                           # :12


class Linear:
    """paddle.nn.Linear(cin, cout): weight [cin,cout], bias [cout]."""

    def __init__(self, cin, cout, rng=None):
        rng = rng or np.random.default_rng(0)
        self.weight = (rng.standard_normal((cin, cout)) / np.sqrt(cin)).astype(F32)
        self.bias = np.zeros((cout,), dtype=F32)

    def __call__(self, x, acc=np.float64):
        return (x.astype(acc) @ self.weight.astype(acc) + self.bias.astype(acc)).astype(F32)


class _Model:
    training = False     # model.eval(); the SA / FP layers' BatchNorms stay on batch statistics regardless
    acc = np.float64

    def train(self, mode=True):
        self.training = mode
        for bn in self._registered_bns():
            bn.training = mode
        return self

    def eval(self):
        return self.train(False)


class _Clas(_Model):
    def _make_head(self, num_classes, rng):
        self.fc1 = Linear(1024, 512, rng)                                 # :17 / :54
        self.bn1 = LN.BatchNorm2D(512)                                    # nn.BatchNorm1D(512)
        self.fc2 = Linear(512, 256, rng)
        self.bn2 = LN.BatchNorm2D(256)
        self.fc3 = Linear(256, num_classes, rng)
        self.eval()

    def _registered_bns(self):
        return [self.bn1, self.bn2]

    def forward(self, inputs, start_idx=(None, None)):
        xyz = np.asarray(inputs, dtype=F32)                               # :26
        B = xyz.shape[0]
        if self.normal_channel:
            norm = xyz[:, 3:, :]                                          # :29
            xyz = xyz[:, :3, :]                                           # :30
        else:
            norm = None
        l1_xyz, l1_points = self.sa1(xyz, norm, start_idx=start_idx[0])   # :33
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        x = l3_points.reshape(B, 1024)                                    # :36
        x = LN.relu(self.bn1(self.fc1(x, self.acc), self.acc))            # :37 (dropout = identity)
        x = LN.relu(self.bn2(self.fc2(x, self.acc), self.acc))            # :38
        return self.fc3(x, self.acc)                                      # :39

    __call__ = forward


class PointNet2_SSG_Clas(_Clas):
    def __init__(self, num_classes=16, normal_channel=False, rng=None):
        in_channel = 6 if normal_channel else 3
        self.normal_channel = normal_channel
        self.sa1 = LN.PointNetSetAbstraction(512, 0.2, 32, in_channel, [64, 64, 128], False, rng)       # :11
        self.sa2 = LN.PointNetSetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256], False, rng)        # :13
        self.sa3 = LN.PointNetSetAbstraction(None, None, None, 256 + 3, [256, 512, 1024], True, rng)    # :15
        self._make_head(num_classes, rng)


class PointNet2_MSG_Clas(_Clas):
    def __init__(self, num_classes=16, normal_channel=False, rng=None):
        in_channel = 3 if normal_channel else 0
        self.normal_channel = normal_channel
        self.sa1 = LN.PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [16, 32, 128], in_channel,
                                                [[32, 32, 64], [64, 64, 128], [64, 96, 128]], rng)      # :48
        self.sa2 = LN.PointNetSetAbstractionMsg(128, [0.2, 0.4, 0.8], [32, 64, 128], 320,
                                                [[64, 64, 128], [128, 128, 256], [128, 128, 256]], rng)  # :49
        self.sa3 = LN.PointNetSetAbstraction(None, None, None, 640 + 3, [256, 512, 1024], True, rng)    # :50
        self._make_head(num_classes, rng)


class _Seg(_Model):
    def _make_head(self, num_parts, rng):
        self.conv1 = LN.Conv2D1x1(128, 128, rng)                          # nn.Conv1D(128,128,1)  :21
        self.bn1 = LN.BatchNorm2D(128)                                    # nn.BatchNorm1D(128)
        self.conv2 = LN.Conv2D1x1(128, num_parts, rng)                    # :24
        self.eval()

    def _registered_bns(self):
        return [self.bn1]

    def forward(self, inputs, start_idx=(None, None)):
        xyz = np.asarray(inputs[0], dtype=F32)                            # :27
        cls_label = Categorical(inputs[1], self.num_classes)              # :28
        B, C, N = xyz.shape
        l0_points = xyz                                                   # :32 / :35
        l0_xyz = xyz[:, :3, :] if self.normal_channel else xyz
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points, start_idx=start_idx[0])   # :37
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)        # :41
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)        # :42
        cls_label_one_hot = np.tile(cls_label.reshape(B, self.num_classes, 1), [1, 1, N])   # :43
        l0_points = self.fp1(l0_xyz, l1_xyz, np.concatenate([cls_label_one_hot, l0_xyz, l0_points], 1),
                             l1_points)                                   # :44
        x = l0_points[:, :, :, None]
        feat = LN.relu(self.bn1(self.conv1(x, self.acc), self.acc))       # :46
        x = self.conv2(feat, self.acc)[:, :, :, 0]                        # :47-48 (dropout = identity)
        return x.transpose(0, 2, 1)                                       # :49

    __call__ = forward


class PointNet2_SSG_Seg(_Seg):
    def __init__(self, num_classes=16, num_parts=50, normal_channel=False, rng=None):
        add = 3 if normal_channel else 0
        self.num_classes = num_classes
        self.normal_channel = normal_channel
        self.sa1 = LN.PointNetSetAbstraction(512, 0.2, 32, 6 + add, [64, 64, 128], False, rng)          # :15
        self.sa2 = LN.PointNetSetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256], False, rng)
        self.sa3 = LN.PointNetSetAbstraction(None, None, None, 256 + 3, [256, 512, 1024], True, rng)
        self.fp3 = LN.PointNetFeaturePropagation(1280, [256, 256], rng)                                  # :18
        self.fp2 = LN.PointNetFeaturePropagation(384, [256, 128], rng)
        self.fp1 = LN.PointNetFeaturePropagation(128 + 16 + 6 + add, [128, 128, 128], rng)
        self._make_head(num_parts, rng)


class PointNet2_MSG_Seg(_Seg):
    def __init__(self, num_classes=16, num_parts=50, normal_channel=False, rng=None):
        add = 3 if normal_channel else 0
        self.num_classes = num_classes
        self.normal_channel = normal_channel
        self.sa1 = LN.PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [32, 64, 128], 3 + add,
                                                [[32, 32, 64], [64, 64, 128], [64, 96, 128]], rng)      # :62
        self.sa2 = LN.PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], 128 + 128 + 64,
                                                [[128, 128, 256], [128, 196, 256]], rng)                # :63
        self.sa3 = LN.PointNetSetAbstraction(None, None, None, 512 + 3, [256, 512, 1024], True, rng)    # :64
        self.fp3 = LN.PointNetFeaturePropagation(1536, [256, 256], rng)                                  # :65
        self.fp2 = LN.PointNetFeaturePropagation(576, [256, 128], rng)
        self.fp1 = LN.PointNetFeaturePropagation(150 + add, [128, 128], rng)
        self._make_head(num_parts, rng)


# ----------------------------------------------------------------------------------
def conv_bn_lists(layer):
    """[(convs, bns), ...] of an SA / MSG / FP layer (oracle or product: same attribute names)."""
    if hasattr(layer, "conv_blocks"):
        return list(zip(layer.conv_blocks, layer.bn_blocks))
    return [(layer.mlp_convs, layer.mlp_bns)]
