"""Drop-in for vdetlib's compiled module ``utils.cython_nms`` (reference utils/nms.pyx).

Same three functions, same argument order and names, same return type (a fresh Python
``list`` of ``int`` indices), same errors:

* ``dets`` / ``tracks`` must be 2-D float32 ndarrays -- the reference's typed buffers
  (utils/nms.pyx:17,71,128-129) raise ``ValueError: Buffer dtype mismatch`` otherwise;
  non-contiguous arrays are accepted;
* ``ZeroDivisionError('float division')`` when a visited pair has union == 0 (nms.pyx:64,
  Cython's cdivision=False).

The arithmetic runs on the GPU (libvdet_b200.so); there is no CPU fallback.
"""
import numpy as np
import torch

from .. import ops


def _typed_buffer(a, name, ncol):
    if not isinstance(a, np.ndarray):
        raise TypeError("Argument '%s' has incorrect type (expected numpy.ndarray, got %s)"
                        % (name, type(a).__name__))
    if a.ndim != 2:
        raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % a.ndim)
    if a.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' but got '%s'" % a.dtype.name)
    if a.shape[1] < ncol:
        raise IndexError("index %d is out of bounds for axis 1 with size %d" % (ncol - 1, a.shape[1]))
    return torch.from_numpy(np.ascontiguousarray(a)).to(ops.default_device())


def nms(dets, thresh):
    """Greedy NMS of one image.  utils/nms.pyx:17-68.  dets [N,5] = (x1,y1,x2,y2,score)."""
    d = _typed_buffer(dets, "dets", 5)
    return ops.nms(d, float(thresh)).cpu().tolist()


def vid_nms(dets, thresh):
    """Whole-video NMS, suppression only inside a frame.  utils/nms.pyx:71-125.
    dets [M,6] = (frame,x1,y1,x2,y2,score); keep order = global descending score."""
    d = _typed_buffer(dets, "dets", 6)
    return ops.vid_nms(d, float(thresh)).cpu().tolist()


def track_det_nms(tracks, dets, thresh):
    """Suppress dets overlapping same-frame track boxes, then vid_nms the rest.
    utils/nms.pyx:128-189.  tracks [Q,5] = (frame,x1,y1,x2,y2); dets [K,6]."""
    t = _typed_buffer(tracks, "tracks", 5)
    d = _typed_buffer(dets, "dets", 6)
    return ops.track_det_nms(t, d, float(thresh)).cpu().tolist()
