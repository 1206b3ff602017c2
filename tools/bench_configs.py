"""Secondary measurements for BASELINE configs 1, 2, 3 and 5 (config 4 is bench.py).  Wall-clock through the C-ABI
(host buffers in, result record out), after warm-up.  Not part of the product."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from g2o_frontend_b200 import capi, synth  # noqa: E402

C = bench.CONF


def params(ctx, step, minr, maxr, minp, dist, outer=10):
    K = synth.scaled_K(synth.K_KINECT, np.float32(1.0) / np.float32(step))
    proj = capi.make_projector(K, 480 // step, 640 // step, C["minD"], C["maxD"])
    sp = capi.make_stats_params(C["worldRadius"], minr, maxr, minp, C["curvatureThreshold"], C["omegaCurvatureThreshold"])
    ap = capi.make_align_params(dist, C["inlierNormalAngularThreshold"], C["flatCurvatureThreshold"],
                                C["inlierCurvatureRatioThreshold"], C["inlierMaxChi2"], True, outer, 1)
    return proj, sp, ap


def timeit(fn, n, warm=3):
    for _ in range(warm):
        fn()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    return (time.perf_counter() - t0) / n


def main():
    out = {}
    ctx = capi.Context(0)
    rawA = synth.render_depth_u16(synth.POSE_A, seed=1)
    rawB = synth.render_depth_u16(synth.POSE_B, seed=2)
    # ---- config 1: one pair, 640x480, 10 iterations
    proj, sp, ap = params(ctx, 1, 10, 30, 50, 1.0)
    cA, cB = ctx.new_cloud(480 * 640), ctx.new_cloud(480 * 640)
    ctx.raw_depth_to_cloud(rawA, proj, sp, cloud=cA)
    ctx.raw_depth_to_cloud(rawB, proj, sp, cloud=cB)
    out["config1_align_us"] = timeit(lambda: ctx.align(cA, cB, proj, ap), 100) * 1e6
    out["config1_frame_prep_us"] = timeit(lambda: (ctx.raw_depth_to_cloud(rawB, proj, sp, cloud=cB), ctx.synchronize()), 100) * 1e6
    out["config1_prep2_plus_align_us"] = timeit(lambda: (ctx.raw_depth_to_cloud(rawA, proj, sp, cloud=cA),
                                                         ctx.raw_depth_to_cloud(rawB, proj, sp, cloud=cB),
                                                         ctx.align(cA, cB, proj, ap)), 100) * 1e6
    # ---- config 2: 3-level pyramid (160x120 / 320x240 / 640x480), clouds built per level from the raw frames
    levels = [(4, 3, 6, 10, 0.5), (2, 5, 15, 25, 0.5), (1, 10, 30, 50, 1.0)]
    lv = [(s,) + params(ctx, s, a, b, c, d) for s, a, b, c, d in levels]
    clouds = [(ctx.new_cloud(480 * 640 // (s * s)), ctx.new_cloud(480 * 640 // (s * s))) for s, *_ in levels]

    def pyramid():
        T = np.eye(4, dtype=np.float32)
        for (s, pj, st, al), (ca, cb) in zip(lv, clouds):
            ctx.raw_depth_to_cloud(rawA, pj, st, step=s, cloud=ca)
            ctx.raw_depth_to_cloud(rawB, pj, st, step=s, cloud=cb)
            T = capi.result_T(ctx.align(ca, cb, pj, al, guess=T))
        return T

    out["config2_pyramid_us"] = timeit(pyramid, 50) * 1e6
    T = pyramid()
    out["config2_translation_error_m"] = float(np.abs(T[:3, 3] - synth.POSE_B[:3, 3]).max())
    # ---- config 3: keyframe tracking over a synthetic sequence (PwnTracker::processFrame logic), 640x480
    n_frames = 200
    poses = synth.trajectory(n_frames, seed=0)
    raws = [synth.render_depth_u16(p, seed=100 + i) for i, p in enumerate(poses)]
    key, cur = ctx.new_cloud(480 * 640), ctx.new_cloud(480 * 640)

    def track():
        nonlocal key, cur
        globalT = np.eye(4)
        keyT = np.eye(4)
        ctx.raw_depth_to_cloud(raws[0], proj, sp, cloud=key)
        nkey = 1
        for i in range(1, n_frames):
            ctx.raw_depth_to_cloud(raws[i], proj, sp, cloud=cur)
            guess = (np.linalg.inv(keyT) @ globalT).astype(np.float32)
            r = ctx.align(key, cur, proj, ap, guess=guess)
            if r.inliers > 0:
                globalT = keyT @ capi.result_T(r).astype(np.float64)
            if r.inliers / float(480 * 640) < 0.4:
                key, cur = cur, key
                keyT = globalT.copy()
                nkey += 1
        return globalT, nkey

    track()
    t0 = time.perf_counter()
    G, nkey = track()
    dt = time.perf_counter() - t0
    gt = np.linalg.inv(poses[0]) @ poses[-1]
    out["config3_frames_per_s"] = (n_frames - 1) / dt
    out["config3_keyframes"] = nkey
    out["config3_final_translation_error_m"] = float(np.abs(G[:3, 3] - gt[:3, 3]).max())
    # ---- config 5: 4 x 1280x960 MultiPointProjector rig
    cams = synth.make_rig(4, 1280, 960)
    gm = capi.make_multi_projector(cams)
    poseA = synth.make_pose((0.1, -0.05, 0.2), (0, 1, 0), 10.0)
    poseB = poseA @ synth.make_pose((0.03, -0.01, 0.04), (0.2, 1.0, 0.1), 2.0)
    dA = synth.u16_to_m(synth.render_rig_depth_u16(poseA, cams))
    dB = synth.u16_to_m(synth.render_rig_depth_u16(poseB, cams))
    gA, gB = ctx.new_cloud(dA.size), ctx.new_cloud(dB.size)
    ctx.multi_depth_to_cloud(dA, gm, sp, cloud=gA)
    ctx.multi_depth_to_cloud(dB, gm, sp, cloud=gB)
    out["config5_frame_prep_us"] = timeit(lambda: ctx.multi_depth_to_cloud(dB, gm, sp, cloud=gB), 20) * 1e6
    out["config5_align_us"] = timeit(lambda: ctx.multi_align(gA, gB, gm, ap), 20) * 1e6
    r = ctx.multi_align(gA, gB, gm, ap)
    Tm = capi.result_T(r)
    gt5 = np.linalg.inv(poseA) @ poseB
    out["config5_translation_error_m"] = float(np.abs(Tm[:3, 3] - gt5[:3, 3]).max())
    out["config5_inliers"] = int(r.inliers)
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
