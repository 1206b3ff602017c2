import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import bench
from g2o_frontend_b200 import capi, synth
raws_cur, raws_cand, pairs, guesses = bench.make_workload(1, 2, 0)
ctx = capi.Context(0)
C = bench.CONF
proj = capi.make_projector(synth.K_KINECT, bench.ROWS, bench.COLS, C["minD"], C["maxD"])
sp = capi.make_stats_params(C["worldRadius"], C["minImageRadius"], C["maxImageRadius"], C["minPoints"], C["curvatureThreshold"], C["omegaCurvatureThreshold"])
ap = capi.make_align_params(C["inlierDistanceThreshold"], C["inlierNormalAngularThreshold"], C["flatCurvatureThreshold"], C["inlierCurvatureRatioThreshold"], C["inlierMaxChi2"], True, 10, 1)
clouds = [ctx.raw_depth_to_cloud(r, proj, sp)[0] for r in raws_cur + raws_cand]
for _ in range(5): ctx.align(clouds[1], clouds[0], proj, ap, guess=guesses[0])
ctx.set_kernel_timing(True)
n = 50
t0 = time.perf_counter()
for _ in range(n): ctx.align(clouds[1], clouds[0], proj, ap, guess=guesses[0])
dt = (time.perf_counter() - t0) / n
kt = ctx.kernel_timing()
print("no-graph align %.1f us | corr_lin %.2f us/launch | project %.2f us/launch" % (dt * 1e6, kt["corr_lin_ms"] / kt["corr_lin_launches"] * 1e3, kt["project_ms"] / kt["project_launches"] * 1e3))
