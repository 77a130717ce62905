#!/usr/bin/env python
"""Host-side cost of getting a det proto to the kernels: the reference's JSON (+gzip) files and dict
walk (utils/protocol.py:209-236, :323-327; vdet/video_det.py:53-56) against the packed container of
vdetlib_b200.utils.packed.  CPU only.

    python tools/packed_io_bench.py [frames] [boxes] [classes]  > profiles/rNN_packed_io.json
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vdetlib_b200 import synth                                   # noqa: E402
from vdetlib_b200.utils import packed, protocol                  # noqa: E402
from vdetlib_b200.vdet.dataset import imagenet_vdet_classes      # noqa: E402


def timed(fn, reps=1):
    best, out = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, out


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    b, s = synth.boxes_scores(T, N, C, seed=1)
    det = synth.det_proto(b, s, imagenet_vdet_classes)
    res = {"frames": T, "boxes_per_frame": N, "classes": C, "detections": T * N, "cpu_count": os.cpu_count()}
    with tempfile.TemporaryDirectory() as d:
        pj, pg, pk = os.path.join(d, "v.det"), os.path.join(d, "g.det.gz"), os.path.join(d, "v.vdetpk")
        res["json_dump_s"] = timed(lambda: protocol.proto_dump(det, pj))[0]
        res["json_gz_dump_s"] = timed(lambda: protocol.proto_dump(det, pg))[0]
        res["packed_dump_s"] = timed(lambda: protocol.proto_dump(det, pk))[0]
        res["json_bytes"], res["json_gz_bytes"], res["packed_bytes"] = (os.path.getsize(p) for p in (pj, pg, pk))
        res["json_load_s"], back = timed(lambda: protocol.proto_load(pj), 2)
        assert back == det
        res["json_gz_load_s"] = timed(lambda: protocol.proto_load(pg[:-3]), 2)[0]
        res["packed_load_to_dicts_s"], back = timed(lambda: protocol.proto_load(pk), 2)
        assert back == det
        # what the kernels need: the float32 matrix of every class
        def walk():
            return [np.asarray([[dd['frame']] + list(dd['bbox']) + [protocol.det_score(dd, c)]
                                for dd in det['detections']], dtype='float32') for c in (1, 2, 3)]
        t3, _ = timed(walk)
        res["dict_walk_to_f32_matrix_s_per_class"] = t3 / 3
        res["dict_walk_to_f32_matrix_s_all_classes_extrapolated"] = t3 / 3 * C
        res["packed_load_to_f32_arrays_s"], arrs = timed(lambda: packed.PackedDets.load(pk).grouped_f32(), 3)
        want = walk()[0]
        order = arrs[3]
        assert np.array_equal(arrs[0], want[order, 1:5]) and np.array_equal(arrs[1][:, 0], want[order, 5])
    res["speedup_file_to_kernel_inputs"] = ((res["json_load_s"] + res["dict_walk_to_f32_matrix_s_all_classes_extrapolated"])
                                            / res["packed_load_to_f32_arrays_s"])
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
