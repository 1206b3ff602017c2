"""How much does the unpinned Eigen version matter?  Frame prep + one alignment with the oracle's restatement of
SelfAdjointEigenSolver<Matrix3f>::computeDirect as Eigen 3.2.x has it and as Eigen >= 3.3 has it, on the synthetic
640x480 scenes and on the real Kinect frame the reference ships.  CPU only (test infrastructure: it runs the oracle);
prints the table quoted in DESIGN.md section 2."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import CONF_1_1  # noqa: E402
from g2o_frontend_b200 import synth  # noqa: E402


def clouds(O, dA, dB, K, conf, variant):
    O.set_eigen_variant(variant)
    sp = O.default_stats_params(minImageRadius=conf["minImageRadius"], maxImageRadius=conf["maxImageRadius"],
                                minPoints=conf["minPoints"], curvatureThreshold=conf["curvatureThreshold"],
                                worldRadius=conf["worldRadius"], omegaCurvatureThreshold=conf["omegaCurvatureThreshold"])
    a = O.depth_to_cloud(dA, K, conf["minD"], conf["maxD"], sp)[0]
    b = O.depth_to_cloud(dB, K, conf["minD"], conf["maxD"], sp)[0]
    O.set_eigen_variant(0)
    return a, b


def study(O, name, dA, dB, K, conf, guess=None):
    cp = O.default_corr_params(inlierDistanceThreshold=conf["inlierDistanceThreshold"],
                               inlierNormalAngularThreshold=conf["inlierNormalAngularThreshold"],
                               flatCurvatureThreshold=conf["flatCurvatureThreshold"],
                               inlierCurvatureRatioThreshold=conf["inlierCurvatureRatioThreshold"])
    out = {}
    res = {}
    for v in (0, 1):
        a, b = clouds(O, dA, dB, K, conf, v)
        ap = O.make_align_params(K, dA.shape[0], dA.shape[1], conf["minD"], conf["maxD"], cp, guess=guess,
                                 max_chi2=conf["inlierMaxChi2"], num_threads=8)
        res[v] = (a, b, O.align(a, b, ap))
    a0, a1 = res[0][0], res[1][0]
    n0, n1 = a0.normals[:, :3].astype(np.float64), a1.normals[:, :3].astype(np.float64)
    both = (np.abs(n0).sum(1) > 0) & (np.abs(n1).sum(1) > 0)
    ang = np.arccos(np.clip((n0[both] * n1[both]).sum(1), -1, 1))
    out["points"] = int(a0.n)
    out["normals_bit_identical"] = float((a0.normals.view(np.uint32) == a1.normals.view(np.uint32)).all(1).mean())
    out["normal_angle_rad"] = {"median": float(np.median(ang)), "p99": float(np.percentile(ang, 99)), "max": float(ang.max())}
    out["zero_normal_flips"] = int(((np.abs(n0).sum(1) > 0) != (np.abs(n1).sum(1) > 0)).sum())
    out["curvature_max_abs_diff"] = float(np.abs(a0.curvature - a1.curvature).max())
    r0, r1 = res[0][2], res[1][2]
    c0, c1 = set(map(tuple, r0.corr.tolist())), set(map(tuple, r1.corr.tolist()))
    out["correspondences"] = [len(c0), len(c1)]
    out["correspondence_jaccard"] = len(c0 & c1) / max(len(c0 | c1), 1)
    R = r0.T[:3, :3].astype(np.float64).T @ r1.T[:3, :3].astype(np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    out["final_T_rotation_diff_rad"] = float(np.arcsin(min(1.0, np.linalg.norm(w))))
    out["final_T_translation_diff_m"] = float(np.abs(r0.T[:3, 3] - r1.T[:3, 3]).max())
    return name, out


def main():
    from oracle import pwn_oracle as O
    K = synth.K_KINECT
    rows = []
    for seed, dropout in ((None, 0.0), (0, 0.05), (2, 0.05)):
        dA = synth.u16_to_m(synth.render_depth_u16(synth.POSE_A, seed=seed, dropout=dropout))
        dB = synth.u16_to_m(synth.render_depth_u16(synth.POSE_B, seed=None if seed is None else seed + 1, dropout=dropout))
        rows.append(study(O, "synthetic 640x480, seed %s, dropout %.2f" % (seed, dropout), dA, dB, K, CONF_1_1))
    real = os.path.join(ROOT, "tests", "golden", "real_depth_640x480.npz")
    if os.path.exists(real):
        z = np.load(real)
        d = synth.u16_to_m(z[z.files[0]])
        g = synth.make_pose((0.02, -0.01, 0.03), (0.2, 1.0, 0.1), 1.5).astype(np.float32)
        rows.append(study(O, "real Kinect frame (reference's test image), self-alignment from a perturbed guess", d, d, K, CONF_1_1, guess=g))
    print(json.dumps(dict(rows), indent=1))


if __name__ == "__main__":
    main()
