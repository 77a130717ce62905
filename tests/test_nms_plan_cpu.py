"""The CTA-shape planner of vdet_nms_frames_f32 (vdetlib_b200/csrc/nms_plan.h, plain C++): compiled with g++ and
checked against the B200 measurements it was derived from (profiles/r02_nms_shapes_T.jsonl, r02_nms_shapes.jsonl).
The planner only chooses between two kernels whose outputs are identical (the GPU sweeps assert "same"); what is
at stake here is time, so the test pins the decisions where the measured gap is clear and bounds the regret."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
#include <cstdio>
#include <cstdlib>
#include "nms_plan.h"
int main(int argc, char** argv) {
    for (int i = 1; i + 2 < argc; i += 3)
        std::printf("%d\n", vdet::nms_prefer_wide(std::atoi(argv[i]), std::atoi(argv[i + 1]), std::atoi(argv[i + 2])) ? 1 : 0);
    std::printf("tail %.4f %.4f %.4f %.4f\n", vdet::nms_tail_cost(0.0), vdet::nms_tail_cost(0.35), vdet::nms_tail_cost(0.68),
                vdet::nms_tail_cost(1.0));
    return 0;
}
"""


@pytest.fixture(scope="module")
def planner(tmp_path_factory):
    d = tmp_path_factory.mktemp("nms_plan")
    src = d / "driver.cpp"
    src.write_text(DRIVER)
    exe = d / "driver"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "vdetlib_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True)

    def ask(cases):
        args = [str(v) for c in cases for v in c]
        out = subprocess.run([str(exe)] + args, check=True, capture_output=True, text=True).stdout.split("\n")
        return [int(v) for v in out[:len(cases)]], out[len(cases)]
    return ask


def _rows(name):
    path = os.path.join(ROOT, "profiles", name)
    return [json.loads(l) for l in open(path) if l.strip()]


def test_planner_follows_the_measured_table(planner):
    rows = [r for r in _rows("r02_nms_shapes_T.jsonl") + _rows("r02_nms_shapes.jsonl") if r["N"] == 300]
    picks, tail = planner([(r["T"], r["C"], 148) for r in rows])
    assert tail.split() == ["tail", "0.0000", "0.4950", "1.0000", "1.0000"]
    regret = []
    for r, wide in zip(rows, picks):
        t_d, t_w = r["threads=256"], r["threads=320"]
        chosen, best = (t_w if wide else t_d), min(t_d, t_w)
        regret.append(chosen / best - 1.0)
        gap = abs(t_d - t_w) / best
        if gap >= 0.05 and 2 * r["T"] > 592:          # a clear winner outside the class-split regime: must be picked
            assert (t_w < t_d) == bool(wide), r
    assert max(regret) <= 0.10, regret                # worst case: the 125-frame launch (class-split regime, 9 %)
    assert sum(regret) / len(regret) <= 0.015


def test_planner_edges(planner):
    picks, _ = planner([(1000, 30, 148), (1000, 30, 146), (0, 30, 148), (1, 30, 148), (296, 30, 148), (1000, 9, 148),
                        (1000, 1, 148), (1184, 30, 148), (100000, 30, 148), (1000, 200, 148)])
    assert picks[0] == 1 and picks[1] == 1            # BASELINE config 2, one rank / with two SMs reserved for NCCL
    assert picks[2] == 0 and picks[3] == 0 and picks[4] == 0      # empty, tiny and class-split launches: default
    assert picks[5] == 0 and picks[6] == 0            # fewer classes than the wide shape has warps
    assert picks[7] == 0                              # exactly two rounds of the default grid
    assert picks[8] == 0                              # long launches: the default's steady state
    assert picks[9] in (0, 1)                         # (200 classes never fit the wide shape: the caller's `fits` guard)
