"""Single-pair latency probe (not part of the product): frame prep and nicp_align wall time at 640x480."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from g2o_frontend_b200 import capi, synth  # noqa: E402


def main():
    raws_cur, raws_cand, pairs, guesses = bench.make_workload(1, 2, 0)
    ctx = capi.Context(0)
    C = bench.CONF
    proj = capi.make_projector(synth.K_KINECT, bench.ROWS, bench.COLS, C["minD"], C["maxD"])
    sp = capi.make_stats_params(C["worldRadius"], C["minImageRadius"], C["maxImageRadius"], C["minPoints"],
                                C["curvatureThreshold"], C["omegaCurvatureThreshold"])
    ap = capi.make_align_params(C["inlierDistanceThreshold"], C["inlierNormalAngularThreshold"], C["flatCurvatureThreshold"],
                                C["inlierCurvatureRatioThreshold"], C["inlierMaxChi2"], True, 10, 1)
    clouds = [ctx.new_cloud(bench.ROWS * bench.COLS) for _ in range(3)]
    raws = raws_cur + raws_cand
    for rep in range(3):
        ctx.synchronize()
        t0 = time.perf_counter()
        n = 50
        for i in range(n):
            ctx.raw_depth_to_cloud(raws[i % 3], proj, sp, cloud=clouds[i % 3])
        ctx.synchronize()
        t_prep = (time.perf_counter() - t0) / n
        t0 = time.perf_counter()
        for i in range(n):
            r = ctx.align(clouds[1], clouds[0], proj, ap, guess=guesses[0])
        t_align = (time.perf_counter() - t0) / n
        t0 = time.perf_counter()
        for i in range(n):
            ctx.raw_depth_to_cloud(raws[0], proj, sp, cloud=clouds[0])
            r = ctx.align(clouds[1], clouds[0], proj, ap, guess=guesses[0])
        t_track = (time.perf_counter() - t0) / n
        print("rep %d: frame prep %.1f us | single align %.1f us (%.0f/s) | tracking step (prep+align) %.1f us (%.0f frames/s) | inliers %d" %
              (rep, t_prep * 1e6, t_align * 1e6, 1 / t_align, t_track * 1e6, 1 / t_track, r.inliers))
    ctx.close()


if __name__ == "__main__":
    main()
