"""Host-side mirror of PAPC/models/detect/pointpillars/libs/ops/non_max_suppression/nms_gpu.py over the sm_100a
kernels (papc_b200/csrc/nms.cu).

Same callables and argument meaning as the reference: NumPy in, Python list / NumPy out.  CUDA torch tensors are
accepted too and then nothing leaves the device (``nms_device`` / ``rotate_iou_device`` are the allocation-aware
forms the detector's post-processing would call).  No CPU path: the kernels are the only implementation.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L


def _dev(device_id):
    if not torch.cuda.is_available():
        raise L.PapcError("papc_b200.nms needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", device_id)


def nms_device(dets, nms_overlap_thresh):
    """dets: CUDA float32 [n,5] (x1,y1,x2,y2,score) or [n,6] (cx,cy,w,h,angle,score) ->
    (keep int32 [n] -- original indices in visiting order, -1 padded --, num int32 [1]), both on the device."""
    L.require_cuda(dets)
    dets = L.f32c(dets)
    n, d = dets.shape
    if d not in (5, 6):
        raise ValueError("dets must be [n,5] (axis aligned) or [n,6] (rotated)")
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=dets.device)
    num = torch.zeros((1,), dtype=torch.int32, device=dets.device)
    lib = L.lib()
    wsb = lib.papc_nms_workspace_bytes(n)
    ws = torch.empty(max(int(wsb), 256), dtype=torch.uint8, device=dets.device)
    L.check(lib.papc_nms_f32(L.ptr(dets), n, d, float(np.float32(nms_overlap_thresh)), L.ptr(keep), L.ptr(num),
                             L.ptr(ws), wsb, L.stream_ptr(dets.device)), "nms")
    return keep[:n], num


def _nms_host(dets, thresh, device_id, width):
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    if dets.ndim != 2 or dets.shape[1] != width:
        raise ValueError(f"dets must be [n,{width}]")
    if dets.shape[0] == 0:
        return []
    keep, num = nms_device(torch.from_numpy(dets).to(_dev(device_id)), thresh)
    return keep[:int(num.item())].cpu().numpy().tolist()


def nms_gpu(dets, nms_overlap_thresh, device_id=0):
    """nms_gpu.py:133-164.  dets [n,5] (x1,y1,x2,y2,score) -> list of kept indices, best score first."""
    return _nms_host(dets, nms_overlap_thresh, device_id, 5)


nms_gpu_cc = nms_gpu   # nms_gpu.py:167-176 (the pybind11 build of cc/nms/nms_kernel.cu.cc): same contract


def rotate_nms_gpu(dets, nms_overlap_thresh, device_id=0):
    """nms_gpu.py:453-488.  dets [n,6] (cx,cy,w,h,angle,score) -> list of kept indices."""
    return _nms_host(dets, nms_overlap_thresh, device_id, 6)


def rotate_iou_device(boxes, query_boxes, criterion=-1):
    """CUDA float32 boxes [N,5], query_boxes [K,5] -> [N,K] on the device."""
    L.require_cuda(boxes, query_boxes)
    boxes, query_boxes = L.f32c(boxes), L.f32c(query_boxes)
    N, K = boxes.shape[0], query_boxes.shape[0]
    out = torch.zeros((N, K), dtype=torch.float32, device=boxes.device)
    L.check(L.lib().papc_rotate_iou_f32(L.ptr(boxes), N, L.ptr(query_boxes), K, int(criterion), L.ptr(out),
                                        L.stream_ptr(boxes.device)), "rotate_iou")
    return out


def rotate_iou_gpu_eval(boxes, query_boxes, criterion=-1, device_id=0):
    """nms_gpu.py:603-653 (criterion -1: IoU, 0: inter / area(query), 1: inter / area(box), 2: inter)."""
    dtype = np.asarray(boxes).dtype
    b = np.ascontiguousarray(boxes, dtype=np.float32)
    q = np.ascontiguousarray(query_boxes, dtype=np.float32)
    if b.shape[0] == 0 or q.shape[0] == 0:
        return np.zeros((b.shape[0], q.shape[0]), dtype=np.float32)
    dev = _dev(device_id)
    out = rotate_iou_device(torch.from_numpy(b).to(dev), torch.from_numpy(q).to(dev), criterion)
    return out.cpu().numpy().astype(dtype)


def rotate_iou_gpu(boxes, query_boxes, device_id=0):
    """nms_gpu.py:518-553."""
    return rotate_iou_gpu_eval(boxes, query_boxes, -1, device_id)
