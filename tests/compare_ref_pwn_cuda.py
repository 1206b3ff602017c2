"""Harness (test infrastructure, lives under tests/ because it loads oracle/_ref): the REFERENCE'S OWN GPU iteration
(g2o_frontend/pwn_cuda/cudaaligner_rk.cu, compiled unmodified into oracle/_ref/libpwn_cuda_ref.so) beside this
library's, on a B200, same clouds and the same transform.

  python tests/compare_ref_pwn_cuda.py [--step 1] [--iters 20]

Prints one JSON object: per-iteration time of the reference's simpleIteration (z-buffer projection with its two-pass
atomicMin(int mm) scheme, 64-thread fused gate + linearise kernel with a 14 KB shared-memory stash, relaunched block
sums, whole-context cudaMemcpy both ways) and of this library's alignment loop, plus how far the two H / b / inlier
counts are apart.  They are NOT expected to agree to rounding: pwn_cuda quantises depth to millimetres, truncates the
pixel coordinates instead of rounding them and has no zero-normal test, so its index images differ from
pwn_core's on a few percent of the pixels; the per-correspondence arithmetic itself is checked exactly on the CPU by
tests/test_reference_pwn_cuda.py.  NOT run by pytest (first GPU use of the reference code is for the next round to
look at; a stale Fermi-era kernel that misbehaves on sm_100a must not turn the test suite red)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import ROOT, CONF_1_1, CONF_1_4  # noqa: E402
from g2o_frontend_b200 import capi, synth  # noqa: E402


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def full16(sym6):
    """n x 6 upper triangles -> n x 16 column-major 4x4 (last row / column zero), the layout pwn_cuda expects"""
    n = sym6.shape[0]
    M = np.zeros((n, 4, 4), np.float32)
    idx = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    for k, (r, c) in enumerate(idx):
        M[:, r, c] = sym6[:, k]
        M[:, c, r] = sym6[:, k]
    return np.ascontiguousarray(M.transpose(0, 2, 1).reshape(n, 16))


def main():
    ap_ = argparse.ArgumentParser()
    ap_.add_argument("--step", type=int, default=1)
    ap_.add_argument("--iters", type=int, default=20)
    args = ap_.parse_args()
    so = os.path.join(ROOT, "oracle", "_ref", "libpwn_cuda_ref.so")
    if not os.path.exists(so):
        print(json.dumps({"unavailable": "oracle/_ref/libpwn_cuda_ref.so not built"}))
        return
    R = C.CDLL(so)
    conf = CONF_1_1 if args.step == 1 else CONF_1_4
    step = args.step
    rows, cols = 480 // step, 640 // step
    K = synth.scaled_K(synth.K_KINECT, 1.0 / step)
    ctx = capi.Context(0)
    proj = capi.make_projector(K, rows, cols, conf["minD"], conf["maxD"])
    sp = capi.make_stats_params(conf["worldRadius"], conf["minImageRadius"], conf["maxImageRadius"], conf["minPoints"],
                                conf["curvatureThreshold"], conf["omegaCurvatureThreshold"])
    ap = capi.make_align_params(conf["inlierDistanceThreshold"], conf["inlierNormalAngularThreshold"],
                                conf["flatCurvatureThreshold"], conf["inlierCurvatureRatioThreshold"], conf["inlierMaxChi2"],
                                True, 10, 1)
    cA, _ = ctx.raw_depth_to_cloud(synth.render_depth_u16(synth.POSE_A, seed=1), proj, sp, step=step)
    cB, _ = ctx.raw_depth_to_cloud(synth.render_depth_u16(synth.POSE_B, seed=2), proj, sp, step=step)
    a, b = cA.download(), cB.download()
    nA, nB = a["points"].shape[0], b["points"].shape[0]
    T = np.eye(4, dtype=np.float32)  # the transform applied to the reference points (Linearizer::T), identity guess

    # ---- ours: stage level (same T) for the comparison, the alignment loop for the timing
    KRt, _ = capi.update_matrices(K, np.eye(4, dtype=np.float32))
    ref_index, _ = ctx.project(cA, KRt, rows, cols, conf["minD"], conf["maxD"])
    cur_index, _ = ctx.project(cB, KRt, rows, cols, conf["minD"], conf["maxD"])
    H, bb, err, inl, nc, _ = ctx.correspond_linearize(cA, cB, ref_index, cur_index, T, ap)
    for _ in range(3):
        ctx.align(cA, cB, proj, ap)
    t0 = time.perf_counter()
    for _ in range(args.iters):
        ctx.align(cA, cB, proj, ap)
    ours_ms_per_iteration = (time.perf_counter() - t0) / args.iters / 10 * 1e3

    # ---- the reference's GPU implementation
    h = C.c_void_p()
    out = {"rows": rows, "cols": cols, "reference_points": int(nA), "current_points": int(nB)}
    rc = R.refcuda_create(C.byref(h), nA, nB, rows, cols)
    if rc:
        out["reference_error"] = "createContext failed with operation code %d" % rc
        print(json.dumps(out))
        return
    params = np.array([conf["inlierDistanceThreshold"] ** 2, conf["inlierNormalAngularThreshold"], conf["flatCurvatureThreshold"],
                       1.0 / conf["inlierCurvatureRatioThreshold"], conf["inlierCurvatureRatioThreshold"],
                       conf["inlierMaxChi2"]], np.float32)
    R.refcuda_set_params(h, fp(params), 1)
    K9 = np.ascontiguousarray(np.asarray(K, np.float32).T.reshape(-1))
    eye = np.ascontiguousarray(np.eye(4, dtype=np.float32).reshape(-1))
    arrs = [np.ascontiguousarray(x, np.float32) for x in
            (a["points"], a["normals"], a["curvature"], b["points"], b["normals"], b["curvature"],
             full16(b["omega_p"]), full16(b["omega_n"]))]
    rc = R.refcuda_init_computation(h, fp(K9), fp(eye), fp(arrs[0]), fp(arrs[1]), fp(arrs[2]), nA, fp(arrs[3]), fp(arrs[4]),
                                    fp(arrs[5]), fp(arrs[6]), fp(arrs[7]), nB)
    if rc:
        out["reference_error"] = "initComputation failed with operation code %d" % rc
        print(json.dumps(out))
        return
    Tc = np.ascontiguousarray(T.T.reshape(-1))
    Hb = np.zeros(56, np.float32)
    rinl, rerr = C.c_int(0), C.c_float(0)
    for _ in range(3):
        rc = R.refcuda_iteration(h, fp(Tc), fp(Hb), C.byref(rinl), C.byref(rerr))
    t0 = time.perf_counter()
    for _ in range(args.iters):
        rc = R.refcuda_iteration(h, fp(Tc), fp(Hb), C.byref(rinl), C.byref(rerr))
    ref_ms = (time.perf_counter() - t0) / args.iters * 1e3
    ri = np.zeros((rows, cols), np.int32)
    ci = np.zeros((rows, cols), np.int32)
    R.refcuda_get_indices(h, ri.ctypes.data_as(C.POINTER(C.c_int)), ci.ctypes.data_as(C.POINTER(C.c_int)))
    R.refcuda_destroy(h)
    Htt, Htr, Hrr = (Hb[16 * j:16 * j + 16].reshape(4, 4).T[:3, :3] for j in range(3))
    Href = np.block([[Htt, Htr], [Htr.T, Hrr]])
    bref = np.concatenate([Hb[48:51], Hb[52:55]])
    out.update({
        "reference_status": int(rc),
        "reference_ms_per_iteration": ref_ms,
        "ours_ms_per_iteration_inside_align": ours_ms_per_iteration,
        "speedup": ref_ms / ours_ms_per_iteration,
        "reference_inliers": int(rinl.value), "ours_inliers": int(inl),
        "reference_index_agreement": float((ri == ref_index).mean()), "current_index_agreement": float((ci == cur_index).mean()),
        "H_rel_diff": float(np.abs(H - Href).max() / np.abs(H).max()),
        "b_rel_diff": float(np.abs(bb - bref).max() / max(np.abs(bb).max(), 1e-9)),
    })
    print(json.dumps(out))
    ctx.close()


if __name__ == "__main__":
    main()
