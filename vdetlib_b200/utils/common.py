"""Drop-in for the hot-path part of vdetlib's ``utils.common``: ``iou`` (utils/common.py:451-468).

Everything else in the reference module (pickle / image crops / MATLAB and Caffe glue) is out
of scope.
"""
import numpy as np
import torch

from .. import ops


def iou(boxes1, boxes2):
    """Dense IoU matrix [A,B], float64, "+1" pixel convention -- utils/common.py:451-468.

    Accepts any array-like (cast to float64 exactly as the reference's ``astype('float')``),
    returns a float64 ndarray.  Computed by the float64 CUDA kernel, bit-identical to NumPy.
    """
    b1 = np.asarray(boxes1).astype('float')
    b2 = np.asarray(boxes2).astype('float')
    if b1.ndim != 2 or b2.ndim != 2:
        # the reference fails on its first fancy index boxes[:, [0]] (common.py:455)
        raise IndexError("too many indices for array")
    if b1.shape[1] < 4 or b2.shape[1] < 4:
        raise IndexError("index 3 is out of bounds for axis 1")
    dev = ops.default_device()
    a = torch.from_numpy(np.ascontiguousarray(b1[:, :4])).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(b2[:, :4])).to(dev)
    return ops.iou_matrix(a, b).cpu().numpy()
