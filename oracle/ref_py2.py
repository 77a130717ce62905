"""Run the reference's OWN Python-2 hot-path functions under Python 3, in memory.

TEST INFRASTRUCTURE ONLY (used by oracle/gen_golden.py and tests that pin the
NumPy restatement while /root/reference is mounted).  Never imported by the product.

vdetlib is Python 2 (print statements, xrange, implicit relative imports, cPickle,
scipy.misc, matlab ...) so its modules cannot be imported here.  Instead this loader

  1. reads a module's source text from /root/reference (read-only, never copied
     into the repository),
  2. cuts out ONE top-level function by name,
  3. applies a purely mechanical Python-2 -> 3 patch that does not touch arithmetic:
        * ``print <expr>`` statement        -> ``print(<expr>)``
        * ``window_size / 2`` (tubelet_cls.py:391, int/int floor division in Py2)
                                            -> ``window_size // 2``
     and provides Py2 built-ins through the exec namespace (``xrange``, list-returning
     ``map`` / ``zip`` / ``filter``, a silent ``print``),
  4. exec()s it in a namespace holding the names the function expects (numpy, copy,
     defaultdict, the other extracted functions, the compiled reference
     ``cython_nms`` from oracle/_ref).

The result is the reference's own code path, statement for statement, producing the
golden vectors in tests/golden/ (see oracle/gen_golden.py).
"""
import builtins
import copy
import logging
import os
import re
from collections import defaultdict
from operator import itemgetter

import numpy as np

REF_ROOT = "/root/reference"


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "utils", "nms.pyx"))


def _cut_function(src, name):
    lines = src.split("\n")
    start = None
    for i, ln in enumerate(lines):
        if re.match(r"def %s\(" % re.escape(name), ln):
            start = i
            break
    if start is None:
        raise KeyError(name)
    end = len(lines)
    for i in range(start + 1, len(lines)):
        ln = lines[i]
        if ln and not ln[0].isspace() and not ln.startswith("#"):
            end = i
            break
    return lines[start:end], start + 1


def _patch_py2(lines):
    out = []
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"^(\s*)print (.*)$", ln)
        if m:
            body = m.group(2)
            while body.count("(") > body.count(")") or body.rstrip().endswith("\\"):
                i += 1
                body = body.rstrip().rstrip("\\") + " " + lines[i].strip()
            ln = "%sprint(%s)" % (m.group(1), body)
        ln = ln.replace("half_window_size = window_size / 2", "half_window_size = window_size // 2")
        out.append(ln)
        i += 1
    return out


def _py2_builtins():
    return {
        "xrange": range,
        "map": lambda f, *a: list(builtins.map(f, *a)),
        "zip": lambda *a: list(builtins.zip(*a)),
        "filter": lambda f, a: list(builtins.filter(f, a)),
        "print": lambda *a, **k: None,
    }


class RefFunctions(object):
    """Namespace of reference functions, extracted lazily."""

    # (module path, function names)
    _WANTED = [
        ("utils/common.py", ["iou"]),
        ("utils/protocol.py", ["det_score", "top_detections", "frame_top_detections",
                               "tubelets_proto_from_tracks_proto", "tubelets_overlap",
                               "merge_score_protos", "tracks_proto_from_boxes",
                               "score_proto", "load_frame_to_det", "load_det_info",
                               "frame_path_at", "boxes_at_frame", "tubelet_box_at_frame"]),
        ("vdet/video_det.py", ["apply_vid_nms", "fast_rcnn_det_vid"]),
        ("vdet/image_det.py", ["apply_image_nms"]),
        ("vdet/tubelet_cls.py", ["do_score_completion", "dets_spatial_max_pooling",
                                 "raw_dets_spatial_max_pooling", "anchor_propagate",
                                 "score_proto_temporal_maxpool", "extrap1d",
                                 "score_proto_interpolation", "score_conv_cls",
                                 "rcnn_sampling_dets_scoring"]),
        ("vdet/track.py", ["greedily_track_from_det", "greedily_track_from_raw_dets"]),
    ]

    def __init__(self, cython_nms=None):
        if not available():
            raise RuntimeError("/root/reference is not mounted")
        if cython_nms is None:
            from . import build_ref
            build_ref.build()
            cython_nms = build_ref.load()
        quiet = logging.getLogger("vdetlib_ref")
        quiet.setLevel(logging.CRITICAL)
        with open(os.path.join(REF_ROOT, "misc", "imagenet_vdet_classes.txt")) as f:
            classes = [line.strip() for line in f.readlines()]        # utils/common.py:28-35
        from scipy.interpolate import interp1d
        import scipy.io as sio
        import hashlib

        def bbox_hash(video_name, frame_id, bbox):                       # utils/protocol.py:372-375 (py3 bytes)
            return hashlib.md5('{}_{}_{}_{}_{}_{}'.format(
                video_name, frame_id, bbox[0], bbox[1], bbox[2], bbox[3]).encode()).hexdigest()

        class Timer(object):                                              # utils/timer.py: timing only
            average_time = 0.0

            def tic(self):
                pass

            def toc(self):
                pass

        ns = {
            "Timer": Timer, "imread": lambda path: path,                    # images are never looked at on this path
            "np": np, "copy": copy, "defaultdict": defaultdict, "itemgetter": itemgetter,
            "logging": quiet, "interp1d": interp1d, "os": os, "sio": sio,
            "imagenet_vdet_classes": classes, "bbox_hash": bbox_hash,
            "nms": cython_nms.nms, "vid_nms": cython_nms.vid_nms,
            "track_det_nms": cython_nms.track_det_nms,
        }
        ns.update(_py2_builtins())
        self.ns = ns
        self.sources = {}
        for rel, names in self._WANTED:
            src = open(os.path.join(REF_ROOT, rel)).read()
            for name in names:
                lines, lineno = _cut_function(src, name)
                code = "\n".join(_patch_py2(lines))
                self.sources[name] = (rel, lineno, code)
                exec(compile("\n" * (lineno - 1) + code, os.path.join(REF_ROOT, rel), "exec"), ns)
        self.classes = classes

    def __getattr__(self, name):
        try:
            return self.ns[name]
        except KeyError:
            raise AttributeError(name)
