"""The four PointNet++ model definitions of the reference, wired line for line on this library's
layers:

  PointNet2_SSG_Clas / PointNet2_MSG_Clas   PAPC/models/classify/pointnet2/pointnet2.py:6-41, :43-78
  PointNet2_SSG_Seg  / PointNet2_MSG_Seg    PAPC/models/segment/pointnet2/pointnet2.py:6-51, :53-98
  Categorical                               PAPC/models/layers/pointnet2_basic_layers.py:7-14

Constructor arguments, attribute names (sa1..sa3, fp1..fp3, fc1..fc3 / conv1, conv2, bn1, bn2, drop1,
drop2) and the forward signatures are the reference's.  The SetAbstraction / FeaturePropagation
layers are this library's CUDA layers (always on batch statistics, as the reference's unregistered
conv/bn lists run); the heads follow ``model.train()`` / ``model.eval()`` like the reference's
registered ``nn.BatchNorm1D`` / ``nn.Dropout``:

  * classification head: three dense layers over B rows (B x 1024 -> 512 -> 256 -> classes).  That is
    library-GEMM work on a 32-row matrix; it runs on torch (cuBLAS, fp32, TF32 off).  Parameters are
    ``torch.nn.Linear`` ones, i.e. ``weight`` is [out,in] -- the transpose of Paddle's [in,out].
  * segmentation head: ``conv1 -> bn1 -> relu`` over B*N point rows runs on this library's tcgen05
    pointwise-MLP kernels (``papc_pointwise_mlp_f32``) straight on the channels-last rows fp1 produced;
    ``conv2`` (128 -> num_parts, no norm) is one fp32 library GEMM whose [B,N,num_parts] result is already
    the transposed layout the reference returns (:49, :96).

``forward`` takes an optional ``start_idx=(sa1_start, sa2_start)`` (the reference draws the FPS start
with ``paddle.randint``, layers.py:76; parity tests pass seeded ones).
"""
from __future__ import annotations

import torch

from . import _lib as L
from .layers import (BatchNorm1DLayer, Conv1DLayer, PointNetFeaturePropagation, PointNetSetAbstraction,
                     PointNetSetAbstractionMsg, _SAMixin, pointwise_mlp_rows, precede)


def Categorical(y, num_class=16, device=None):
    """layers.py:7-14: one-hot encode integer labels y [B,1] (or [B]) -> float32 [B,num_class,1]."""
    y = torch.as_tensor(y, device=device).reshape(-1).long()
    if y.numel() and (int(y.min()) < 0 or int(y.max()) >= num_class):
        raise IndexError(f"Categorical: label outside [0,{num_class})")   # np.eye(num_class)[y,] raises too
    return torch.nn.functional.one_hot(y, num_class).to(torch.float32).unsqueeze(2)


class PaddleBatchNorm1D(torch.nn.Module):
    """``paddle.nn.BatchNorm1D(c)`` on [B,c] rows: epsilon 1e-5, momentum 0.9, the BIASED batch variance
    both normalises and enters the running ``_variance`` (torch's BatchNorm1d tracks the unbiased one)."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(num_features))
        self.bias = torch.nn.Parameter(torch.zeros(num_features))
        self.register_buffer("_mean", torch.zeros(num_features))
        self.register_buffer("_variance", torch.ones(num_features))
        self._momentum, self._epsilon = momentum, epsilon

    def forward(self, x):
        if self.training:
            mean, var = x.mean(0), x.var(0, unbiased=False)
            with torch.no_grad():
                self._mean.mul_(self._momentum).add_(mean, alpha=1.0 - self._momentum)
                self._variance.mul_(self._momentum).add_(var, alpha=1.0 - self._momentum)
        else:
            mean, var = self._mean, self._variance
        return (x - mean) * torch.rsqrt(var + self._epsilon) * self.weight + self.bias


class _ClsHead(torch.nn.Module):
    """fc1/bn1/drop1/fc2/bn2/drop2/fc3 of classify/pointnet2/pointnet2.py:17-23 (:36-39)."""

    def _make_head(self, num_classes, p2):
        self.fc1 = torch.nn.Linear(1024, 512)
        self.bn1 = PaddleBatchNorm1D(512)
        self.drop1 = torch.nn.Dropout(0.4)
        self.fc2 = torch.nn.Linear(512, 256)
        self.bn2 = PaddleBatchNorm1D(256)
        self.drop2 = torch.nn.Dropout(p2)
        self.fc3 = torch.nn.Linear(256, num_classes)

    def _head(self, l3_points, B):
        F = torch.nn.functional
        x = l3_points.reshape(B, 1024)
        x = self.drop1(F.relu(self.bn1(self.fc1(x))))
        x = self.drop2(F.relu(self.bn2(self.fc2(x))))
        return self.fc3(x)

    def _split(self, inputs):
        xyz = torch.as_tensor(inputs)
        L.require_cuda(xyz)
        if self.normal_channel:
            return xyz[:, :3, :], xyz[:, 3:, :]
        return xyz, None


class PointNet2_SSG_Clas(_ClsHead):
    """classify/pointnet2/pointnet2.py:6-41."""

    def __init__(self, name_scope='PointNet2_SSG_Clas_', num_classes=16, normal_channel=False):
        super().__init__()
        in_channel = 6 if normal_channel else 3
        self.normal_channel = normal_channel
        self.sa1 = PointNetSetAbstraction(npoint=512, radius=0.2, nsample=32, in_channel=in_channel,
                                          mlp=[64, 64, 128], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=128, radius=0.4, nsample=64, in_channel=128 + 3,
                                          mlp=[128, 128, 256], group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=256 + 3,
                                          mlp=[256, 512, 1024], group_all=True)
        self._make_head(num_classes, 0.4)

    def forward(self, inputs, start_idx=(None, None)):
        xyz, norm = self._split(inputs)
        B = xyz.shape[0]
        precede(start_idx[1])
        l1_xyz, l1_points = self.sa1(xyz, norm, start_idx=start_idx[0])
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        return self._head(l3_points, B)


class PointNet2_MSG_Clas(_ClsHead):
    """classify/pointnet2/pointnet2.py:43-78."""

    def __init__(self, name_scope='PointNet2_MSG_Clas_', num_classes=16, normal_channel=False):
        super().__init__()
        in_channel = 3 if normal_channel else 0
        self.normal_channel = normal_channel
        self.sa1 = PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [16, 32, 128], in_channel,
                                             [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
        self.sa2 = PointNetSetAbstractionMsg(128, [0.2, 0.4, 0.8], [32, 64, 128], 320,
                                             [[64, 64, 128], [128, 128, 256], [128, 128, 256]])
        self.sa3 = PointNetSetAbstraction(None, None, None, 640 + 3, [256, 512, 1024], True)
        self._make_head(num_classes, 0.5)

    forward = PointNet2_SSG_Clas.forward


class _SegHead(_SAMixin):
    """conv1/bn1/drop1/conv2 of segment/pointnet2/pointnet2.py:21-24 (:45-49).  conv1/bn1 are this
    library's parameter holders (they feed the pointwise-MLP kernel); bn1 follows train()/eval()."""

    def _make_head(self, num_parts):
        # registered sublayers, as in the reference (:21-24): their tensors are in state_dict() under
        # Paddle's keys (conv1.weight, bn1._mean, ...) and move with .to() like any module
        self.conv1 = Conv1DLayer(128, 128, 1)
        self.bn1 = BatchNorm1DLayer(128)
        self.drop1 = torch.nn.Dropout(0.5)
        self.conv2 = Conv1DLayer(128, num_parts, 1)

    def _holders(self):
        return []

    def _head(self, l0_points):
        B, C, N = l0_points.shape
        rows = L.f32c(l0_points.transpose(1, 2)).reshape(B * N, C)          # fp1's own buffer, no copy
        feat = pointwise_mlp_rows(rows, C, [self.conv1], [self.bn1],       # :45 relu(bn1(conv1(.)))
                                  "batch" if self.training else "running", update_running=self.training,
                                  sync=self._sync_group())
        x = self.drop1(feat)                                                # :46
        w2 = self.conv2.weight.reshape(self.conv2.weight.shape[0], -1)
        x = torch.addmm(self.conv2.bias, x, w2.t())                         # :47
        return x.reshape(B, N, -1)                                          # :48 [B,N,num_parts]

    def _inputs(self, inputs):
        xyz = torch.as_tensor(inputs[0])
        L.require_cuda(xyz)
        cls_label = Categorical(inputs[1], self.num_classes, device=xyz.device)   # :28 / :75  [B,16,1]
        B, C, N = xyz.shape
        if cls_label.shape[0] != B:
            raise ValueError("one class label per cloud expected")
        l0_points = xyz
        l0_xyz = xyz[:, :3, :] if self.normal_channel else xyz
        return B, N, l0_xyz, l0_points, cls_label

    def forward(self, inputs, start_idx=(None, None)):
        B, N, l0_xyz, l0_points, cls_label = self._inputs(inputs)
        precede(start_idx[1])
        l1_xyz, l1_points = self.sa1(l0_xyz, l0_points, start_idx=start_idx[0])
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        cls_label_one_hot = cls_label.reshape(B, self.num_classes, 1).expand(B, self.num_classes, N)
        l0_points = self.fp1(l0_xyz, l1_xyz, torch.cat([cls_label_one_hot, l0_xyz, l0_points], 1), l1_points)
        return self._head(l0_points)


class PointNet2_SSG_Seg(_SegHead):
    """segment/pointnet2/pointnet2.py:6-51."""

    def __init__(self, name_scope='PointNet2_SSG_Seg_', num_classes=16, num_parts=50, normal_channel=False):
        super().__init__()
        additional_channel = 3 if normal_channel else 0
        self.num_classes = num_classes
        self.normal_channel = normal_channel
        self.sa1 = PointNetSetAbstraction(npoint=512, radius=0.2, nsample=32, in_channel=6 + additional_channel,
                                          mlp=[64, 64, 128], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=128, radius=0.4, nsample=64, in_channel=128 + 3,
                                          mlp=[128, 128, 256], group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=256 + 3,
                                          mlp=[256, 512, 1024], group_all=True)
        self.fp3 = PointNetFeaturePropagation(in_channel=1280, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=384, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(in_channel=128 + 16 + 6 + additional_channel, mlp=[128, 128, 128])
        self._make_head(num_parts)


class PointNet2_MSG_Seg(_SegHead):
    """segment/pointnet2/pointnet2.py:53-98."""

    def __init__(self, name_scope='PointNet2_MSG_Seg_', num_classes=16, num_parts=50, normal_channel=False):
        super().__init__()
        additional_channel = 3 if normal_channel else 0
        self.num_classes = num_classes
        self.normal_channel = normal_channel
        self.sa1 = PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [32, 64, 128], 3 + additional_channel,
                                             [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
        self.sa2 = PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], 128 + 128 + 64,
                                             [[128, 128, 256], [128, 196, 256]])
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=512 + 3,
                                          mlp=[256, 512, 1024], group_all=True)
        self.fp3 = PointNetFeaturePropagation(in_channel=1536, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=576, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(in_channel=150 + additional_channel, mlp=[128, 128])
        self._make_head(num_parts)
