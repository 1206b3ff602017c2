"""The oracle against the REFERENCE'S OWN pwn_core sources.

oracle/build_ref_pwn_core.sh compiles g2o_frontend/pwn_core/*.cpp from /root/reference against the Eigen / OpenCV
stand-ins of oracle/shim/ into oracle/_ref/libpwn_core_ref.so (extern "C" face: oracle/ref_pwn_core.cpp).  Every loop,
index computation, clamp, gate, branch, accumulation order and OpenMP chunking below is therefore the reference's code
as written; what Eigen would compute inside (fixed-size products in index order, computeDirect, LDLT, ...) is supplied
by the stand-in, the numerical kernels by the oracle's own restatements -- those stay "parity unpinned" (DESIGN.md 2).

The bar is BIT-EXACT for everything up to and including H, b and the alignment result (same arithmetic, same order),
which is what makes the oracle a restatement rather than an approximation of the reference.  CPU only; runs here and
on the GPU box (the prebuilt .so travels)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, get_scene

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libpwn_core_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libpwn_core_ref.so not built "
                                "(needs /root/reference; run __graft_entry__.build() in the container)")


def fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def cm(M):
    return np.ascontiguousarray(np.asarray(M, np.float32).T.reshape(-1))


@pytest.fixture(scope="module")
def R():
    from oracle import pwn_oracle as O
    O.lib()  # liboracle.so (the stand-in's numerical kernels resolve against it) is loaded first
    L = C.CDLL(REF_SO)
    L.refcore_depth_to_cloud.restype = C.c_void_p
    L.refcore_set_threads(1)
    return L


class RefCloud:
    """a pwn::Cloud built by the reference's DepthImageConverterIntegralImage::compute"""

    def __init__(self, R, depth, K, conf, sensor_offset=None):
        rows, cols = depth.shape
        self.R = R
        sp = np.array([conf["worldRadius"], conf["minImageRadius"], conf["maxImageRadius"], conf["minPoints"],
                       conf["curvatureThreshold"], conf["omegaCurvatureThreshold"]], np.float32)
        self.index = np.zeros((rows, cols), np.int32)
        self.interval = np.zeros((rows, cols), np.int32)
        self.integral = np.zeros((rows, cols, 10), np.float32)
        off = cm(np.eye(4) if sensor_offset is None else sensor_offset)
        self.h = C.c_void_p(R.refcore_depth_to_cloud(fp(np.ascontiguousarray(depth, np.float32)), rows, cols, fp(cm(K)),
                                                     C.c_float(conf["minD"]), C.c_float(conf["maxD"]), fp(sp), fp(off),
                                                     ip(self.index), ip(self.interval), fp(self.integral)))
        n = self.n = R.refcore_cloud_size(self.h)
        self.points = np.zeros((n, 4), np.float32)
        self.normals = np.zeros((n, 4), np.float32)
        self.statsM = np.zeros((n, 16), np.float32)
        self.eigvals = np.zeros((n, 3), np.float32)
        self.statsN = np.zeros(n, np.int32)
        self.curvature = np.zeros(n, np.float32)
        self.omegaP = np.zeros((n, 16), np.float32)
        self.omegaN = np.zeros((n, 16), np.float32)
        self.fetch()

    def fetch(self):
        self.R.refcore_cloud_get(self.h, fp(self.points), fp(self.normals), fp(self.statsM), fp(self.eigvals), ip(self.statsN),
                                 fp(self.curvature), fp(self.omegaP), fp(self.omegaN))

    def __del__(self):
        try:
            self.R.refcore_cloud_free(self.h)
        except Exception:
            pass


def same_cloud(ref, orc):
    assert ref.n == orc.n
    for k in ("points", "normals", "statsM", "eigvals", "statsN", "curvature", "omegaP", "omegaN"):
        a, b = getattr(ref, k), getattr(orc, k)
        assert np.array_equal(a, b, equal_nan=True), (k, int((a != b).sum()))


def finder_params(conf):
    return np.array([conf["inlierDistanceThreshold"], conf["inlierNormalAngularThreshold"], conf["flatCurvatureThreshold"],
                     conf["inlierCurvatureRatioThreshold"]], np.float32)


# ---------------------------------------------------------------------------------------------------------------
def test_depth_conversion_and_scaling(R):
    """pwn_static.cpp:5-68"""
    from oracle import pwn_oracle as O
    S = get_scene(4, seed=3, dropout=0.05)
    raw = S.rawA
    d = np.zeros(raw.shape, np.float32)
    R.refcore_depth_u16_to_f32(raw.ctypes.data_as(C.POINTER(C.c_ushort)), raw.shape[0], raw.shape[1], C.c_float(0.001), fp(d))
    assert np.array_equal(d, O.depth_u16_to_f32(raw))
    for step in (1, 2, 3, 4, 7):
        out = np.zeros((raw.shape[0] // step, raw.shape[1] // step), np.float32)
        R.refcore_depth_scale(fp(d), raw.shape[0], raw.shape[1], step, C.c_float(0.01), fp(out))
        assert np.array_equal(out, O.depth_scale(d, step)), step


@pytest.mark.parametrize("case", ["clean", "noise_dropout", "sensor_offset", "full_resolution"])
def test_depth_to_cloud_is_bit_identical(R, case):
    """DepthImageConverterIntegralImage::compute: unProject, projectIntervals, PointIntegralImage (all 10 channels, every
    pixel), StatsCalculatorIntegralImage (region clamps, minPoints, eigen-decomposition, curvature, normal flip),
    Point/NormalInformationMatrixCalculator, Cloud::transformInPlace"""
    S = {"clean": lambda: get_scene(4), "noise_dropout": lambda: get_scene(4, seed=1, dropout=0.05),
         "sensor_offset": lambda: get_scene(4, offset=True), "full_resolution": lambda: get_scene(1)}[case]()
    ref = RefCloud(R, S.depthA, S.K, S.conf, S.sensor_offset)
    assert np.array_equal(ref.index, S.indexA)
    assert np.array_equal(ref.interval, S.intervalA)
    assert np.array_equal(ref.integral.reshape(-1), np.asarray(S.integralA, np.float32).reshape(-1))
    same_cloud(ref, S.cloudA)
    assert (np.abs(ref.normals[:, :3]).sum(1) > 0).mean() > 0.5  # not vacuous: most points have a normal


def test_projection_is_bit_identical(R):
    """PinholePointProjector::project (z-buffer: nearest wins, first index wins ties, empty = FLT_MAX / -1) and the
    3-argument unProject at a non-identity pose"""
    from oracle import pwn_oracle as O
    S = get_scene(4)
    c = S.conf
    ref = RefCloud(R, S.depthA, S.K, c)
    for v in ([0, 0, 0, 0, 0, 0], [0.03, -0.02, 0.05, 0.01, -0.015, 0.005], [-0.2, 0.1, 0.3, -0.05, 0.08, 0.02],
              [0.5, 0.0, -1.0, 0.0, 0.3, 0.0]):
        T = O.v2t(np.array(v, np.float32))
        ii = np.zeros((S.rows, S.cols), np.int32)
        dd = np.zeros((S.rows, S.cols), np.float32)
        R.refcore_project(ref.h, fp(cm(S.K)), fp(cm(T)), S.rows, S.cols, C.c_float(c["minD"]), C.c_float(c["maxD"]), ip(ii), fp(dd))
        oi, od = O.project(S.cloudA.points, S.rows, S.cols, S.K, T, c["minD"], c["maxD"])
        assert np.array_equal(ii, oi) and np.array_equal(dd, od), v
        pts = np.zeros((S.rows * S.cols, 4), np.float32)
        idx = np.zeros((S.rows, S.cols), np.int32)
        n = R.refcore_unproject(fp(np.ascontiguousarray(S.depthA)), S.rows, S.cols, fp(cm(S.K)), fp(cm(T)), C.c_float(c["minD"]),
                                C.c_float(c["maxD"]), fp(pts), ip(idx))
        opts, oidx = O.unproject(S.depthA, S.K, T, c["minD"], c["maxD"])
        assert n == len(opts) and np.array_equal(pts[:n], opts) and np.array_equal(idx, oidx), v


@pytest.mark.parametrize("threads", [1, 2, 3, 4, 7, 8, 16])
@pytest.mark.parametrize("robust", [True, False])
def test_finder_and_linearizer_are_bit_identical(R, threads, robust):
    """CorrespondenceFinder::compute + Linearizer::update, including what the reference's OpenMP partitioning does for
    every thread count: rows % numThreads rows are never searched (correspondencefinder.cpp:38), numCorrespondences %
    numThreads correspondences are never linearised (linearizer.cpp:32-39), per-thread float32 partial sums added in
    thread order."""
    from oracle import pwn_oracle as O
    S = get_scene(4, seed=2)
    c = S.conf
    refA, refB = RefCloud(R, S.depthA, S.K, c), RefCloud(R, S.depthB, S.K, c)
    T = O.v2t(np.array([0.01, -0.02, 0.015, 0.004, -0.003, 0.002], np.float32))
    ri, _ = O.project(S.cloudA.points, S.rows, S.cols, S.K, np.linalg.inv(T).astype(np.float32), c["minD"], c["maxD"])
    ci, _ = O.project(S.cloudB.points, S.rows, S.cols, S.K, np.eye(4, dtype=np.float32), c["minD"], c["maxD"])
    chi2 = 9e3 if robust else 60.0
    R.refcore_set_threads(threads)
    try:
        corr = np.full((S.rows * S.cols, 2), -1, np.int32)
        H = np.zeros(36, np.float32)
        b = np.zeros(6, np.float32)
        err, inl = C.c_float(0), C.c_int(0)
        n = R.refcore_correspond_linearize(refA.h, refB.h, ip(np.ascontiguousarray(ri)), ip(np.ascontiguousarray(ci)), S.rows,
                                           S.cols, fp(cm(T)), fp(finder_params(c)), C.c_float(chi2), int(robust), ip(corr),
                                           fp(H), fp(b), C.byref(err), C.byref(inl))
    finally:
        R.refcore_set_threads(1)
    ocorr, _ = O.correspond(ri, ci, S.cloudA, S.cloudB, T, S.cp, num_threads=threads)
    oH, ob, oe, oinl = O.linearize(ocorr, S.cloudA, S.cloudB, T, chi2, robust, num_threads=threads)
    assert n == len(ocorr) > 5000
    assert np.array_equal(corr[:n], ocorr)
    assert inl.value == oinl and err.value == oe
    assert np.array_equal(H.reshape(6, 6).T, oH) and np.array_equal(b, ob)


def run_ref_align(R, refA, refB, S, outer=10, inner=1, guess=None, ref_off=None, cur_off=None, priors=(), chi2=None, threads=1):
    c = S.conf
    P = S.rows * S.cols
    out = dict(T=np.zeros(16, np.float32), omega=np.zeros(36, np.float32), ratios=np.zeros(2, np.float32),
               refIndex=np.zeros((S.rows, S.cols), np.int32), refDepth=np.zeros((S.rows, S.cols), np.float32),
               curIndex=np.zeros((S.rows, S.cols), np.int32), curDepth=np.zeros((S.rows, S.cols), np.float32),
               corr=np.full((P, 2), -1, np.int32))
    err, inl = C.c_float(0), C.c_int(0)
    eye = np.eye(4, dtype=np.float32)
    pr = np.zeros(max(len(priors), 1) * 69, np.float32)
    for j, (kind, mean, info, reference) in enumerate(priors):
        pr[69 * j] = kind
        pr[69 * j + 1:69 * j + 17] = cm(mean)
        pr[69 * j + 17:69 * j + 33] = cm(eye if reference is None else reference)
        pr[69 * j + 33:69 * j + 69] = cm(info)
    R.refcore_set_threads(threads)
    try:
        n = R.refcore_align(refA.h, refB.h, fp(cm(S.K)), S.rows, S.cols, C.c_float(c["minD"]), C.c_float(c["maxD"]),
                            fp(finder_params(c)), C.c_float(c["inlierMaxChi2"] if chi2 is None else chi2), 1, outer, inner,
                            fp(cm(eye if guess is None else guess)), fp(cm(eye if ref_off is None else ref_off)),
                            fp(cm(eye if cur_off is None else cur_off)), fp(pr), len(priors), fp(out["T"]), fp(out["omega"]),
                            C.byref(err), C.byref(inl), fp(out["ratios"]), ip(out["refIndex"]), fp(out["refDepth"]),
                            ip(out["curIndex"]), fp(out["curDepth"]), ip(out["corr"]))
    finally:
        R.refcore_set_threads(1)
    out["n"] = n
    out["error"], out["inliers"] = err.value, inl.value
    out["T"] = out["T"].reshape(4, 4).T.copy()
    out["omega"] = out["omega"].reshape(6, 6).T.copy()
    return out


@pytest.mark.parametrize("case", ["identity_guess", "perturbed_guess_4_threads", "inner_iterations", "sensor_offset",
                                  "one_iteration", "thirteen_iterations"])
def test_align_is_bit_identical(R, case):
    """Aligner::align free-running: projection of the current cloud, then per outer iteration projection of the reference
    cloud, finder, lineariser, damping, LDLT step, v2t / t2v renormalisation -- T, error, inliers, the finder's index and
    depth images and the correspondence list of the last iteration all bit-identical; omega and the eigen-ratios of
    _computeStatistics to rounding."""
    from g2o_frontend_b200 import synth
    S = get_scene(4, offset=(case == "sensor_offset"))
    refA = RefCloud(R, S.depthA, S.K, S.conf, S.sensor_offset)
    refB = RefCloud(R, S.depthB, S.K, S.conf, S.sensor_offset)
    kw = {"identity_guess": dict(), "perturbed_guess_4_threads": dict(guess=synth.make_pose((0.02, 0.01, -0.03), (0.1, 1.0, 0.3), 1.5), threads=4),
          "inner_iterations": dict(inner=3, outer=4), "sensor_offset": dict(ref_off=S.sensor_offset, cur_off=S.sensor_offset),
          "one_iteration": dict(outer=1), "thirteen_iterations": dict(outer=13)}[case]
    ref = run_ref_align(R, refA, refB, S, **kw)
    from oracle import pwn_oracle as O
    ap = S.oracle_align_params(outer=kw.get("outer"), inner=kw.get("inner"), guess=kw.get("guess"), num_threads=kw.get("threads", 1))
    orc = O.align(S.cloudA, S.cloudB, ap)
    assert np.array_equal(ref["T"], orc.T), np.abs(ref["T"] - orc.T).max()
    assert ref["error"] == orc.error and ref["inliers"] == orc.inliers and ref["n"] == orc.numCorrespondences
    assert np.array_equal(ref["refIndex"], orc.refIndex) and np.array_equal(ref["curIndex"], orc.curIndex)
    assert np.array_equal(ref["refDepth"], orc.refDepth) and np.array_equal(ref["curDepth"], orc.curDepth)
    assert np.array_equal(ref["corr"][:ref["n"]], orc.corr)
    assert np.abs(ref["omega"] - orc.omega).max() <= 1e-5 * np.abs(orc.omega).max()
    assert abs(ref["ratios"][0] - orc.translationalRatio) <= 1e-4 * orc.translationalRatio
    assert abs(ref["ratios"][1] - orc.rotationalRatio) <= 1e-4 * orc.rotationalRatio
    # and the alignment is a real one: it recovers the ground-truth motion of the synthetic pair
    if kw.get("outer", 10) >= 10 and case != "sensor_offset":
        assert np.abs(ref["T"] - S.gt).max() < 5e-3


def test_align_with_priors_is_bit_identical(R):
    """SE3RelativePrior / SE3AbsolutePrior (se3_prior.cpp:8-71) inside the Gauss-Newton step (aligner.cpp:97-108)"""
    from oracle import pwn_oracle as O
    from g2o_frontend_b200 import synth
    S = get_scene(4)
    refA, refB = RefCloud(R, S.depthA, S.K, S.conf), RefCloud(R, S.depthB, S.K, S.conf)
    mean = synth.make_pose((0.03, -0.02, 0.05), (0.2, 1.0, 0.1), 2.0)
    info = np.diag([2000, 2000, 2000, 5000, 5000, 5000]).astype(np.float32)
    reference = synth.make_pose((0.5, 0.1, -0.2), (0.0, 1.0, 0.0), 10.0)
    absmean = (reference @ mean).astype(np.float32)
    for name, priors, opriors in (
            ("relative", [(0, mean, info, None)], [O.make_prior(0, mean, info)]),
            ("absolute", [(1, absmean, info, reference)], [O.make_prior(1, absmean, info, reference)]),
            ("both", [(0, mean, info, None), (1, absmean, info, reference)],
             [O.make_prior(0, mean, info), O.make_prior(1, absmean, info, reference)])):
        ref = run_ref_align(R, refA, refB, S, priors=priors)
        c = S.conf
        ap = O.make_align_params(S.K, S.rows, S.cols, c["minD"], c["maxD"], S.cp, max_chi2=c["inlierMaxChi2"], num_threads=1,
                                 priors=opriors)
        orc = O.align(S.cloudA, S.cloudB, ap)
        assert np.array_equal(ref["T"], orc.T), (name, np.abs(ref["T"] - orc.T).max())
        assert ref["inliers"] == orc.inliers and ref["error"] == orc.error, name
