"""The oracle against the REFERENCE'S OWN code.

pwn_core cannot be built here (Eigen / OpenCV absent), but the reference's CUDA implementation of the iteration,
g2o_frontend/pwn_cuda, depends on nothing except the CUDA runtime.  oracle/Makefile compiles cudaaligner_rk.cu unmodified
from /root/reference into oracle/_ref/libpwn_cuda_ref.so behind the extern "C" face of oracle/ref_pwn_cuda.cu; its
small-linear-algebra helpers are __host__ __device__ and its per-correspondence member is made host-callable by the
wrapper, so the comparisons below run on the CPU (here and on the GPU box, where the prebuilt .so travels):

  * bm_se3.h restated by the reference itself in plain float32 (cudasla.cu:137-200: v2t, t2v, isometry inverse, skew)
  * AlignerContext::processCorrespondence (cudaaligner_rk.cu:558-678): the three correspondence gates, the robust
    kernel and the Htt / Htr / Hrr / bt / br terms of one correspondence -- what CorrespondenceFinder::compute +
    Linearizer::update do per pixel.

tools/ref_pwn_cuda_compare.py runs the reference's whole GPU iteration beside ours on a B200."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, get_scene

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpwn_cuda_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libpwn_cuda_ref.so not built "
                                "(needs /root/reference; run __graft_entry__.build() in the container)")


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.fixture(scope="module")
def ref():
    return C.CDLL(REF_SO)


def test_se3_helpers_match_the_reference(ref):
    from oracle import pwn_oracle as O
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(2000):
        v = np.concatenate([rng.uniform(-2, 2, 3), rng.uniform(-0.3, 0.3, 3)]).astype(np.float32)
        m = np.zeros(16, np.float32)
        ref.refcuda_v2t(fp(v), fp(m))
        T = O.v2t(v)
        worst = max(worst, float(np.abs(m.reshape(4, 4).T - T).max()))
        # t2v of a rotation whose quaternion has w > 0 (both sides take the trace branch; the reference's cudasla
        # version does not normalise or flip the sign, pwn_core's mat2quat does: bm_se3.h:24-34)
        v2 = np.zeros(6, np.float32)
        ref.refcuda_t2v(fp(m), fp(v2))
        worst = max(worst, float(np.abs(v2 - O.t2v(T)).max()))
        inv = np.zeros(16, np.float32)
        ref.refcuda_transform_inverse(fp(m), fp(inv))
        worst = max(worst, float(np.abs(inv.reshape(4, 4).T @ T - np.eye(4)).max()))
    assert worst < 2e-6, worst
    # skew(v) = -2 [v]x (bm_se3.h:54-66), the convention the Linearizer's Jacobian is written in
    v = np.array([0.3, -0.7, 1.1, 0.0], np.float32)
    S = np.zeros(16, np.float32)
    ref.refcuda_skew(fp(v), fp(S))
    S = S.reshape(4, 4).T
    x, y, z = 2 * v[:3]
    assert np.array_equal(S[:3, :3], np.array([[0, z, -y], [-z, 0, x], [y, -x, 0]], np.float32))
    assert not S[3].any() and not S[:, 3].any()


@pytest.mark.parametrize("robust", [True, False])
@pytest.mark.parametrize("conf", ["1_4", "tight", "chi2"])
def test_gates_and_linearizer_term_match_the_reference(ref, robust, conf):
    """3000 (reference point, current point) pairs of the synthetic frames, one at a time: the reference's
    processCorrespondence against the oracle's CorrespondenceFinder (1x1 index images) + Linearizer (one correspondence)."""
    from oracle import pwn_oracle as O
    S = get_scene(4)
    A, B = S.cloudA, S.cloudB
    dist, ncos, flat, ratio, chi2 = {"1_4": (0.5, 0.95, 0.02, 1.3, 9e3), "tight": (0.08, 0.995, 0.004, 1.05, 9e3),
                                     "chi2": (0.5, 0.9, 0.02, 1.3, 40.0)}[conf]
    cp = O.default_corr_params(inlierDistanceThreshold=dist, inlierNormalAngularThreshold=ncos, flatCurvatureThreshold=flat,
                               inlierCurvatureRatioThreshold=ratio)
    params = np.array([dist * dist, ncos, flat, 1.0 / ratio, ratio, chi2], np.float32)
    T = O.v2t(np.array([0.01, -0.02, 0.015, 0.004, -0.003, 0.002], np.float32))
    Tc = np.ascontiguousarray(T.T.reshape(-1).astype(np.float32))
    rng = np.random.default_rng(1)
    n_acc = n_rej = n_scaled = 0
    worst = 0.0
    for k in range(3000):
        ri = int(rng.integers(A.n))
        ci = int(rng.integers(B.n)) if k % 3 == 0 else min(max(ri + int(rng.integers(-3, 4)), 0), B.n - 1)
        if not A.normals[ri, :3].any() or not B.normals[ci, :3].any():
            continue  # pwn_core skips zero normals first (correspondencefinder.cpp:69); pwn_cuda has no such test
        out = np.zeros(56, np.float32)
        err = C.c_float(0)
        r = ref.refcuda_process_correspondence(fp(Tc), fp(A.points[ri]), fp(A.normals[ri]), C.c_float(A.curvature[ri]),
                                               fp(B.points[ci]), fp(B.normals[ci]), C.c_float(B.curvature[ci]),
                                               fp(B.omegaP[ci]), fp(B.omegaN[ci]), fp(params), int(robust), fp(out),
                                               C.byref(err))
        corr, _ = O.correspond(np.array([[ri]], np.int32), np.array([[ci]], np.int32), A, B, T, cp, num_threads=1)
        inl = 0
        if len(corr):
            H, b, e, inl = O.linearize(corr, A, B, T, chi2, robust, num_threads=1)
        assert (r != 0) == (inl != 0), (k, ri, ci, r, len(corr), inl)
        if not r:
            n_rej += 1
            continue
        n_acc += 1
        Htt, Htr, Hrr = (out[16 * j:16 * j + 16].reshape(4, 4).T[:3, :3] for j in range(3))
        Href = np.block([[Htt, Htr], [Htr.T, Hrr]])
        bref = np.concatenate([out[48:51], out[52:55]])
        assert out[51] == 1.0  # bt[3] counts the inlier
        chi = float(out[55])    # br[3] = chi2 before the robust scaling
        worst = max(worst, float(np.abs(H - Href).max() / max(np.abs(Href).max(), 1e-6)),
                    float(np.abs(b - bref).max() / max(np.abs(bref).max(), 1e-6)))
        if chi > chi2:
            n_scaled += 1   # robust kernel: Linearizer::update adds k * chi2 with k = sqrt(max / chi2) (linearizer.cpp:56-64)
            assert abs(e - np.sqrt(chi2 / chi) * chi) <= 2e-6 * abs(e)
        else:
            assert abs(e - chi) <= 2e-6 * max(abs(e), 1e-3)
    assert n_acc > 100 and n_rej > 100, (n_acc, n_rej)
    if conf == "chi2" and robust:
        assert n_scaled >= 10, n_scaled
    assert worst < 2e-6, worst
