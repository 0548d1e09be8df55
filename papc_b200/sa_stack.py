"""The three SetAbstraction layers of PointNet2_SSG_Clas / PointNet2_MSG_Seg wired exactly as the
reference model definitions wire them (the FC heads are outside the metric):
  SSG: PAPC/models/classify/pointnet2/pointnet2.py:11-16, forward :33-35
  MSG: PAPC/models/segment/pointnet2/pointnet2.py:62-64, forward :84-86
"""
from __future__ import annotations

import numpy as np
import torch

from .layers import PointNetFeaturePropagation, PointNetSetAbstraction, PointNetSetAbstractionMsg, precede


def load_conv_bn(convs, bns, params):
    """Install explicit parameters (list of dicts with weight [cout,cin], bias, gamma, beta)."""
    for conv, bn, p in zip(convs, bns, params):
        w = torch.as_tensor(np.asarray(p["weight"]))
        conv.weight = w.reshape(w.shape[0], w.shape[1], 1, 1).to(conv.weight.device)
        conv.bias = torch.as_tensor(np.asarray(p["bias"])).to(conv.bias.device)
        bn.weight = torch.as_tensor(np.asarray(p["gamma"])).to(bn.weight.device)
        bn.bias = torch.as_tensor(np.asarray(p["beta"])).to(bn.bias.device)


class SSGSetAbstractionStack(torch.nn.Module):
    """sa1 -> sa2 -> sa3 of PointNet2_SSG_Clas (normal_channel=False): [B,3,N] -> l3_points [B,1024,1]."""

    def __init__(self, in_channel=3):
        super().__init__()
        self.sa1 = PointNetSetAbstraction(npoint=512, radius=0.2, nsample=32, in_channel=in_channel,
                                          mlp=[64, 64, 128], group_all=False)
        self.sa2 = PointNetSetAbstraction(npoint=128, radius=0.4, nsample=64, in_channel=128 + 3,
                                          mlp=[128, 128, 256], group_all=False)
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=256 + 3,
                                          mlp=[256, 512, 1024], group_all=True)

    def layers_(self):
        return [self.sa1, self.sa2, self.sa3]

    def forward(self, xyz, norm=None, start_idx=(None, None)):
        precede(start_idx[1])   # ready here, before sa1 is enqueued: sa2's overlapped sampling waits for this only
        l1_xyz, l1_points = self.sa1(xyz, norm, start_idx=start_idx[0])
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        return l3_xyz, l3_points


class MSGSegSetAbstractionStack(torch.nn.Module):
    """sa1 -> sa2 -> sa3 of PointNet2_MSG_Seg (normal_channel=False; features = xyz, 3 ch)."""

    def __init__(self, additional_channel=0):
        super().__init__()
        self.sa1 = PointNetSetAbstractionMsg(512, [0.1, 0.2, 0.4], [32, 64, 128], 3 + additional_channel,
                                             [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
        self.sa2 = PointNetSetAbstractionMsg(128, [0.4, 0.8], [64, 128], 128 + 128 + 64,
                                             [[128, 128, 256], [128, 196, 256]])
        self.sa3 = PointNetSetAbstraction(npoint=None, radius=None, nsample=None, in_channel=512 + 3,
                                          mlp=[256, 512, 1024], group_all=True)

    def forward(self, xyz, points, start_idx=(None, None)):
        precede(start_idx[1])
        l1_xyz, l1_points = self.sa1(xyz, points, start_idx=start_idx[0])
        l2_xyz, l2_points = self.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.sa3(l2_xyz, l2_points)
        return l3_xyz, l3_points


class MSGSegEncoderDecoder(torch.nn.Module):
    """SetAbstraction + FeaturePropagation part of PointNet2_MSG_Seg (normal_channel=False), wired as
    segment/pointnet2/pointnet2.py:62-67 and :84-93: sa1 -> sa2 -> sa3 -> fp3 -> fp2 -> fp1 -> l0_points
    [B,128,N].  ``cls_one_hot`` is the ``Categorical`` label tile [B,16,N] (layers.py:7-14) the caller
    supplies; the FC head (conv1/bn1/conv2, :68-71) is outside this library's scope."""

    def __init__(self, num_classes=16, additional_channel=0):
        super().__init__()
        self.enc = MSGSegSetAbstractionStack(additional_channel)
        self.fp3 = PointNetFeaturePropagation(in_channel=1536, mlp=[256, 256])
        self.fp2 = PointNetFeaturePropagation(in_channel=576, mlp=[256, 128])
        self.fp1 = PointNetFeaturePropagation(in_channel=128 + num_classes + 6 + additional_channel, mlp=[128, 128])

    def forward(self, xyz, cls_one_hot, start_idx=(None, None)):
        l0_xyz, l0_points = xyz, xyz
        precede(start_idx[1])
        l1_xyz, l1_points = self.enc.sa1(l0_xyz, l0_points, start_idx=start_idx[0])
        l2_xyz, l2_points = self.enc.sa2(l1_xyz, l1_points, start_idx=start_idx[1])
        l3_xyz, l3_points = self.enc.sa3(l2_xyz, l2_points)
        l2_points = self.fp3(l2_xyz, l3_xyz, l2_points, l3_points)
        l1_points = self.fp2(l1_xyz, l2_xyz, l1_points, l2_points)
        l0_points = self.fp1(l0_xyz, l1_xyz, torch.cat([cls_one_hot, l0_xyz, l0_points], dim=1), l1_points)
        return l0_points


class GraphedForward:
    """A forward pass captured once in a CUDA graph and replayed: the ~25 launches, event forks and
    memsets of a SetAbstraction stack become one ``cudaGraphLaunch`` (the side-stream sampling
    overlap and the programmatic dependent launches are captured as graph edges).  Inputs live in
    static buffers: ``g(x)`` copies ``x`` into them, replays, and returns the static outputs (valid
    until the next replay).  Shapes and parameters are frozen at capture time.

        g = GraphedForward(lambda x: model(x, None, start_idx=(st1, st2)), xyz_example)
        l3_xyz, l3_points = g(xyz)
    """

    def __init__(self, fn, *example_inputs, warmup=3):
        from . import _lib as L
        self.inputs = tuple(t.clone() for t in example_inputs)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):      # warm up off the capture: one-time kernel attribute calls,
            for _ in range(warmup):        # allocator pools, lazy module loads
                fn(*self.inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        lib = L.lib()
        n0 = lib.papc_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn(*self.inputs)
        self.kernels_per_replay = int(lib.papc_launch_count() - n0)   # library kernels inside the graph

    def replay(self):
        self.graph.replay()
        return self.outputs

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            dst.copy_(src, non_blocking=True)
        return self.replay()
