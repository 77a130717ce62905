"""Drop-in for the post-CNN part of vdetlib's ``vdet.image_det`` (reference vdet/image_det.py).

Only ``apply_image_nms`` (:117-123) is on the hot path; the Caffe forward passes are out of scope.
"""
import logging

import numpy as np

from ..utils.cython_nms import nms


def apply_image_nms(boxes, scores, thres=0.3):
    """[N,4] boxes + [N] scores -> keep list.  vdet/image_det.py:117-123."""
    box_score = np.asarray(np.r_['-1', boxes, np.reshape(scores, (-1, 1))], dtype='float32')
    logging.info("Applying nms to image.")
    keep = nms(box_score, thres)
    logging.info("{} / {} boxes kept.".format(len(keep), len(boxes)))
    return keep
