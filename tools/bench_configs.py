"""BASELINE configs 1, 2, 3 and 5 (config 4 is bench.py's main line): wall clock through the C-ABI (host buffers in,
result record out) after warm-up, each beside a bounded sample of the same work on the CPU oracle (performance build,
all host threads).  bench.py embeds `run_all()` as the `configs` block of its JSON line at N = 1; run stand-alone it
prints the block.  Measurement infrastructure, not part of the product."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from g2o_frontend_b200 import capi, synth  # noqa: E402

C = bench.CONF
LEVELS = [(4, 3, 6, 10, 0.5), (2, 5, 15, 25, 0.5), (1, 10, 30, 50, 1.0)]  # step, minR, maxR, minPoints, inlier distance


def params(step, minr, maxr, minp, dist, outer=10):
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


def _render(args):
    pose, seed = args
    return synth.render_depth_u16(pose, seed=seed)


def render_sequence(n_frames):
    """config 3's sequence: n smooth poses, one noisy 640x480 frame each; rendered by a process pool
    (0.1 s per frame in numpy)"""
    poses = synth.trajectory(n_frames, seed=0)
    jobs = [(p, 100 + i) for i, p in enumerate(poses)]
    procs = min(os.cpu_count() or 1, 32)
    if procs > 1 and n_frames >= 64:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(procs) as pool:
            raws = pool.map(_render, jobs, chunksize=8)
    else:
        raws = [_render(j) for j in jobs]
    return poses, raws


# The CPU legs (`cpu`, an instance of bench.CpuLeg or None) live in bench.py: it is the one measurement script allowed to
# execute the oracle.


def config1(ctx, cpu):
    """one 640x480 pair, 10 iterations, identity guess (pwn_simple_aligner's inner step)"""
    rawA = synth.render_depth_u16(synth.POSE_A, seed=1)
    rawB = synth.render_depth_u16(synth.POSE_B, seed=2)
    proj, sp, ap = params(1, 10, 30, 50, 1.0)
    cA, cB = ctx.new_cloud(480 * 640), ctx.new_cloud(480 * 640)
    ctx.raw_depth_to_cloud(rawA, proj, sp, cloud=cA)
    ctx.raw_depth_to_cloud(rawB, proj, sp, cloud=cB)
    out = {"workload": "one 640x480 pair, 10 iterations, pwn_aligner_1_1.conf"}
    out["align_us"] = timeit(lambda: ctx.align(cA, cB, proj, ap), 100) * 1e6
    out["frame_prep_us"] = timeit(lambda: (ctx.raw_depth_to_cloud(rawB, proj, sp, cloud=cB), ctx.synchronize()), 100) * 1e6
    out["prep2_plus_align_us"] = timeit(lambda: (ctx.raw_depth_to_cloud(rawA, proj, sp, cloud=cA),
                                                 ctx.raw_depth_to_cloud(rawB, proj, sp, cloud=cB),
                                                 ctx.align(cA, cB, proj, ap)), 100) * 1e6
    out["value"], out["unit"] = 1e6 / out["prep2_plus_align_us"], "pairs/s (2 frame preps + align, host buffers in)"
    if cpu:
        def one():
            a, b = cpu.cloud(rawA, 1, 10, 30, 50), cpu.cloud(rawB, 1, 10, 30, 50)
            cpu.align(a, b, 1, 1.0)
        t = timeit(one, 2, warm=1)
        out["cpu_baseline"] = {"value": 1.0 / t, "unit": "pairs/s", "cores": cpu.cores, "kind": "port",
                               "sample": "2 x (2 cloud builds + 1 alignment), 1 warm-up"}
    return out


def config2(ctx, cpu):
    """3-level pyramid 160x120 / 320x240 / 640x480 (DepthImage_scale steps 4, 2, 1), T carried down"""
    rawA = synth.render_depth_u16(synth.POSE_A, seed=1)
    rawB = synth.render_depth_u16(synth.POSE_B, seed=2)
    lv = [(s,) + params(s, a, b, c, d) for s, a, b, c, d in LEVELS]
    clouds = [(ctx.new_cloud(480 * 640 // (s * s)), ctx.new_cloud(480 * 640 // (s * s))) for s, *_ in LEVELS]

    def pyramid():
        T = np.eye(4, dtype=np.float32)
        for (s, pj, st, al), (ca, cb) in zip(lv, clouds):
            ctx.raw_depth_to_cloud(rawA, pj, st, step=s, cloud=ca)
            ctx.raw_depth_to_cloud(rawB, pj, st, step=s, cloud=cb)
            T = capi.result_T(ctx.align(ca, cb, pj, al, guess=T))
        return T

    out = {"workload": "3-level pyramid (160x120, 320x240, 640x480), per level 2 cloud builds from the raw frames + align"}
    out["pyramid_us"] = timeit(pyramid, 50) * 1e6
    T = pyramid()
    out["translation_error_m"] = float(np.abs(T[:3, 3] - synth.POSE_B[:3, 3]).max())
    out["value"], out["unit"] = 1e6 / out["pyramid_us"], "pyramids/s"
    if cpu:
        def one():
            T = None
            for s, a, b, c, d in LEVELS:
                ca, cb = cpu.cloud(rawA, s, a, b, c), cpu.cloud(rawB, s, a, b, c)
                T = cpu.align(ca, cb, s, d, guess=T).T
        t = timeit(one, 1, warm=1)
        out["cpu_baseline"] = {"value": 1.0 / t, "unit": "pyramids/s", "cores": cpu.cores, "kind": "port",
                               "sample": "1 pyramid, 1 warm-up"}
    return out


def config3(ctx, cpu, n_frames=2000):
    """keyframe tracking over a synthetic 640x480 sequence (PwnTracker::processFrame logic, pwn_tracker.cpp:106-282):
    per frame one cloud build from the raw image + one alignment against the key frame"""
    proj, sp, ap = params(1, 10, 30, 50, 1.0)
    poses, raws = render_sequence(n_frames)
    try:
        # the frames a camera driver hands over sit in pinned memory (like bench.py's main workload): the upload of a frame
        # is then asynchronous instead of a blocking staged copy
        import torch
        pinned = torch.empty((n_frames, 480, 640), dtype=torch.int16).pin_memory()
        host = pinned.numpy().view(np.uint16)
        for i, r in enumerate(raws):
            host[i] = r
        raws = [host[i] for i in range(n_frames)]
    except Exception:
        pinned = None
    key, cur = ctx.new_cloud(480 * 640), ctx.new_cloud(480 * 640)

    def track(n):
        nonlocal key, cur
        globalT = np.eye(4)
        keyT = np.eye(4)
        ctx.raw_depth_to_cloud(raws[0], proj, sp, cloud=key)
        nkey = 1
        for i in range(1, n):
            ctx.raw_depth_to_cloud(raws[i], proj, sp, cloud=cur)
            guess = (np.linalg.inv(keyT) @ globalT).astype(np.float32)
            r = ctx.align(key, cur, proj, ap, guess=guess)
            if r.inliers > 0:
                globalT = keyT @ capi.result_T(r).astype(np.float64)
            if r.inliers / float(480 * 640) < 0.4:  # pwn_tracker.cpp:164-167: new key frame
                key, cur = cur, key
                keyT = globalT.copy()
                nkey += 1
        return globalT, nkey

    # The same tracker with the cloud of frame i + 1 built (second context = second stream of the same GPU) while frame i
    # is aligned: the frames are still processed in order with the same guesses, so the poses are the sequential ones bit
    # for bit; three cloud buffers (key, current, next).
    prep = capi.Context(ctx.device)
    bufs = [prep.new_cloud(480 * 640) for _ in range(3)]

    def track_pipelined(n):
        globalT = np.eye(4)
        keyT = np.eye(4)
        key, cur, free = bufs[0], bufs[1], bufs[2]
        prep.raw_depth_to_cloud(raws[0], proj, sp, cloud=key)
        if n > 1:
            prep.raw_depth_to_cloud(raws[1], proj, sp, cloud=cur)
        nkey = 1
        for i in range(1, n):
            prep.synchronize()  # the cloud of frame i is complete
            if i + 1 < n:
                prep.raw_depth_to_cloud(raws[i + 1], proj, sp, cloud=free)
            guess = (np.linalg.inv(keyT) @ globalT).astype(np.float32)
            r = ctx.align(key, cur, proj, ap, guess=guess)
            if r.inliers > 0:
                globalT = keyT @ capi.result_T(r).astype(np.float64)
            if r.inliers / float(480 * 640) < 0.4:
                key, cur, free = cur, free, key
                keyT = globalT.copy()
                nkey += 1
            else:
                cur, free = free, cur
        prep.synchronize()
        return globalT, nkey

    track(min(100, n_frames))
    t0 = time.perf_counter()
    G, nkey = track(n_frames)
    dt = time.perf_counter() - t0
    track_pipelined(min(100, n_frames))
    t0 = time.perf_counter()
    G2, nkey2 = track_pipelined(n_frames)
    dt2 = time.perf_counter() - t0
    for b in bufs:
        b.close()
    prep.close()
    gt = np.linalg.inv(poses[0]) @ poses[-1]
    out = {"workload": "sequential keyframe tracking, %d-frame synthetic 640x480 sequence" % n_frames,
           "frames": n_frames, "value": (n_frames - 1) / dt2, "unit": "frames/s", "sequence_s": dt2, "keyframes": nkey2,
           "schedule": "the cloud of frame i + 1 is built on a second stream while frame i is aligned (same poses as the "
                       "strictly sequential schedule: poses_identical below)",
           "strictly_sequential": {"value": (n_frames - 1) / dt, "unit": "frames/s", "sequence_s": dt, "keyframes": nkey},
           "poses_identical": bool(np.array_equal(G, G2) and nkey == nkey2),
           "final_translation_error_m": float(np.abs(G2[:3, 3] - gt[:3, 3]).max())}
    if cpu:
        m = min(6, n_frames)
        t0 = time.perf_counter()
        k = cpu.cloud(raws[0], 1, 10, 30, 50)
        for i in range(1, m):
            c = cpu.cloud(raws[i], 1, 10, 30, 50)
            cpu.align(k, c, 1, 1.0)
        t = (time.perf_counter() - t0) / (m - 1)
        out["cpu_baseline"] = {"value": 1.0 / t, "unit": "frames/s", "cores": cpu.cores, "kind": "port",
                               "sample": "the first %d frames of the sequence (cloud build + alignment per frame)" % m}
    return out


def config5(ctx, cpu):
    """MultiPointProjector rig, 4 x 1280x960 (composite 1280 x 3840)"""
    proj, sp, ap = params(1, 10, 30, 50, 1.0)
    cams = synth.make_rig(4, 1280, 960)
    gm = capi.make_multi_projector(cams)
    poseA = synth.make_pose((0.1, -0.05, 0.2), (0, 1, 0), 10.0)
    poseB = poseA @ synth.make_pose((0.03, -0.01, 0.04), (0.2, 1.0, 0.1), 2.0)
    dA = synth.u16_to_m(synth.render_rig_depth_u16(poseA, cams))
    dB = synth.u16_to_m(synth.render_rig_depth_u16(poseB, cams))
    gA, gB = ctx.new_cloud(dA.size), ctx.new_cloud(dB.size)
    ctx.multi_depth_to_cloud(dA, gm, sp, cloud=gA)
    ctx.multi_depth_to_cloud(dB, gm, sp, cloud=gB)
    out = {"workload": "MultiPointProjector rig 4 x 1280x960 (composite 1280x3840), 10 iterations"}
    out["frame_prep_us"] = timeit(lambda: ctx.multi_depth_to_cloud(dB, gm, sp, cloud=gB), 20) * 1e6
    out["align_us"] = timeit(lambda: ctx.multi_align(gA, gB, gm, ap), 20) * 1e6
    r = ctx.multi_align(gA, gB, gm, ap)
    Tm = capi.result_T(r)
    gt5 = np.linalg.inv(poseA) @ poseB
    out["translation_error_m"] = float(np.abs(Tm[:3, 3] - gt5[:3, 3]).max())
    out["inliers"] = int(r.inliers)
    out["value"], out["unit"] = 1e6 / out["align_us"], "alignments/s"
    if cpu:
        t = cpu.multi_pair_align_seconds(cams, dA, dB)
        out["cpu_baseline"] = {"value": 1.0 / t, "unit": "alignments/s", "cores": cpu.cores, "kind": "port",
                               "sample": "1 alignment of the rig pair (clouds prebuilt)"}
    return out


def run_all(ctx=None, cpu=None, tracking_frames=2000):
    own = ctx is None
    if own:
        ctx = capi.Context(0)
    out = {}
    for name, fn in (("1_single_pair", lambda: config1(ctx, cpu)), ("2_pyramid", lambda: config2(ctx, cpu)),
                     ("3_tracking", lambda: config3(ctx, cpu, tracking_frames)), ("5_multi_projector", lambda: config5(ctx, cpu))):
        try:
            out[name] = fn()
        except Exception as e:  # a secondary measurement never takes the main line down
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}
    if own:
        ctx.close()
    return out


if __name__ == "__main__":
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    print(json.dumps(run_all(cpu=bench.CpuLeg(), tracking_frames=frames)))
