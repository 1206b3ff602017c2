"""Generates tests/golden/*.npz.

The reference ships no golden vectors, so these fixtures are produced by the CPU oracle
(oracle/pwn_oracle.c, verification build) on seeded synthetic inputs.  They give the GPU tests box-independent
expected values and pin the oracle against accidental change.  They are CERTIFIED BY THE REFERENCE'S OWN SOURCES:
tests/test_reference_pwn_core.py::test_committed_golden_fixtures_are_what_the_reference_computes recomputes
small_pair.npz and map_ops_small.npz with oracle/_ref/libpwn_core_ref.so (g2o_frontend/pwn_core/*.cpp compiled against
the Eigen / OpenCV stand-ins of oracle/shim) and finds them bit-identical; what Eigen computes internally
(eigen3.npz: computeDirect) stays unpinned.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from conftest import get_scene  # noqa: E402
from oracle import pwn_oracle as O  # noqa: E402


def main():
    # 80x60 frames (scale 8 would change K; use a crop-free render at step 8 through DepthImage_scale)
    s = get_scene(4, 0, 0.05)
    rows, cols = 30, 40
    dA = np.ascontiguousarray(s.depthA[40:40 + rows, 60:60 + cols])
    dB = np.ascontiguousarray(s.depthB[40:40 + rows, 60:60 + cols])
    K = s.K.copy()
    K[0, 2] -= 60
    K[1, 2] -= 40
    sp = O.default_stats_params(minImageRadius=3, maxImageRadius=6, minPoints=10, curvatureThreshold=0.2)
    cA, iA, itvA, integA = O.depth_to_cloud(dA, K, 0.5, 4.5, sp, want_aux=True)
    cB, iB = O.depth_to_cloud(dB, K, 0.5, 4.5, sp)
    cp = O.default_corr_params(inlierDistanceThreshold=0.5, inlierNormalAngularThreshold=0.95)
    ap = O.make_align_params(K, rows, cols, 0.5, 4.5, cp, num_threads=1)
    out = O.align(cA, cB, ap)
    np.savez_compressed(
        os.path.join(HERE, "small_pair.npz"), depthA=dA, depthB=dB, K=K,
        pointsA=cA.points, normalsA=cA.normals, curvatureA=cA.curvature, omegaPA=cA.omegaP6(), omegaNA=cA.omegaN6(),
        eigvalsA=cA.eigvals, statsNA=cA.statsN, indexA=iA, intervalA=itvA, integralA=integA,
        pointsB=cB.points, normalsB=cB.normals, curvatureB=cB.curvature, omegaPB=cB.omegaP6(), omegaNB=cB.omegaN6(),
        T=out.T, H=out.H, b=out.b, error=out.error, inliers=out.inliers, numCorr=out.numCorrespondences,
        corr=out.corr, refIndex=out.refIndex, refDepth=out.refDepth, curIndex=out.curIndex,
        trace_T=np.stack(out.trace_T), trace_H=np.stack(out.trace_H), trace_b=np.stack(out.trace_b), omega=out.omega)
    # local-map maintenance: gaussians of frame A, the two-frame map A + add(B, T), Merger::merge, VoxelCalculator
    gA, fA, _, _ = O.unproject_gaussians(dA, K, 0.5, 4.5, 0.075, 0.1)
    gB, fB, _, _ = O.unproject_gaussians(dB, K, 0.5, 4.5, 0.075, 0.1)
    import ctypes as C
    fp = lambda x: x.ctypes.data_as(C.POINTER(C.c_float))
    T = out.T.astype(np.float32)
    ptsB, nrmB, stB, opB, onB = (cB.points.copy(), cB.normals.copy(), cB.statsM.copy(), cB.omegaP.copy(), cB.omegaN.copy())
    O.lib().orc_cloud_transform(fp(O.colmajor(T)), cB.n, fp(ptsB), fp(nrmB), fp(stB), fp(opB), fp(onB))
    gB, fB = O.gaussians_transform(T, gB, fB)
    m = O.Cloud(cA.n + cB.n)
    cat = lambda a, b: np.ascontiguousarray(np.concatenate([a, b]))
    m.points, m.normals, m.statsM = cat(cA.points, ptsB), cat(cA.normals, nrmB), cat(cA.statsM, stB)
    m.omegaP, m.omegaN = cat(cA.omegaP, opB), cat(cA.omegaN, onB)
    m.eigvals, m.statsN, m.curvature = cat(cA.eigvals, cB.eigvals), cat(cA.statsN, cB.statsN), cat(cA.curvature, cB.curvature)
    mg, mf = cat(gA, gB), cat(fA, fB)
    res, g2, f2, col = O.merge(m, mg, mf, rows, cols, K, np.eye(4, dtype=np.float32), 0.5, 4.5)
    vox = O.voxelize(m.points, 0.05, strict=True)
    vox_raw = O.voxelize(m.points, 0.05, strict=False)
    np.savez_compressed(os.path.join(HERE, "map_ops_small.npz"), depthA=dA, depthB=dB, K=K, T=T,
                        gaussA=gA, map_points=m.points, map_normals=m.normals, map_gauss=mg, map_flags=mf,
                        merged_points=res.points, merged_gauss=g2, merged_flags=f2, collapsed=col,
                        voxel_rep=vox, voxel_rep_as_written=vox_raw)
    # eigen-solver known answers
    rng = np.random.default_rng(7)
    mats, evs, vecs = [], [], []
    for i in range(64):
        A = rng.standard_normal((3, 3)).astype(np.float32)
        Cm = (A @ A.T).astype(np.float32) * np.float32(10.0 ** rng.integers(-6, 2))
        if i % 4 == 0:  # planar patch: one tiny eigenvalue
            Q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
            Cm = (Q @ np.diag([1e-6, 0.01, 0.02]) @ Q.T).astype(np.float32)
        Cm = ((Cm + Cm.T) / 2).astype(np.float32)
        ev, U = O.eigen3(Cm)
        mats.append(Cm)
        evs.append(ev)
        vecs.append(U)
    np.savez_compressed(os.path.join(HERE, "eigen3.npz"), C=np.stack(mats), evals=np.stack(evs), evecs=np.stack(vecs))
    print("wrote golden fixtures:", os.listdir(HERE))


if __name__ == "__main__":
    main()
