"""Batched frame preparation (nicp_raw_depth_to_cloud_batch): device time per 640x480 frame and the fraction of the HBM
roofline by SURVEY.md 8d's algorithmic bytes (2 P raw + 84 P + 76 N: the index image is not materialised in a batch).
Measurement tool, not part of the product."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from g2o_frontend_b200 import capi, synth  # noqa: E402


def main():
    F = int(os.environ.get("PREP_FRAMES", 64))
    reps = int(os.environ.get("PREP_REPS", 10))
    C = bench.CONF
    rng = np.random.default_rng(3)
    poses = [synth.perturbed_pose(rng, np.eye(4), 0.25, 6.0) for _ in range(8)]
    base = [synth.render_depth_u16(p, seed=i) for i, p in enumerate(poses)]
    pinned = torch.empty((F, bench.ROWS, bench.COLS), dtype=torch.int16).pin_memory()
    host = pinned.numpy().view(np.uint16)
    for i in range(F):
        host[i] = base[i % len(base)]
    frames = [host[i] for i in range(F)]
    ctx = capi.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=dev)
    proj = capi.make_projector(synth.K_KINECT, bench.ROWS, bench.COLS, C["minD"], C["maxD"])
    sp = capi.make_stats_params(C["worldRadius"], C["minImageRadius"], C["maxImageRadius"], C["minPoints"],
                                C["curvatureThreshold"], C["omegaCurvatureThreshold"])
    clouds = [ctx.new_cloud(bench.ROWS * bench.COLS) for _ in range(F)]
    out = {"frames": F}
    for name, fn in (("batch", lambda: ctx.raw_depth_to_cloud_batch(frames, proj, sp, clouds=clouds)),
                     ("single", lambda: [ctx.raw_depth_to_cloud(frames[i], proj, sp, cloud=clouds[i]) for i in range(F)])):
        for _ in range(3):
            fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        ctx.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * F)
        n_pts = float(np.mean([c.size() for c in clouds[:8]]))
        P = bench.ROWS * bench.COLS
        byts = 2 * P + 84 * P + 76 * n_pts
        out[name] = {"us_per_frame": us, "algorithmic_MB_per_frame": byts / 1e6, "GBps": byts / us / 1e3,
                     "frac_of_hbm_peak": byts / us / 1e3 / bench.measured_peak()[0]}
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
