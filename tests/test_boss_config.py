"""BOSS configuration files: the record format of g2o_frontend/pwn_boss, parsed into
the C-ABI parameter structs.  CPU only."""
import glob
import os

import numpy as np
import pytest

from conftest import ROOT
from g2o_frontend_b200 import boss_config as B

FIXTURE = os.path.join(ROOT, "tests", "golden", "boss_pipeline.conf")
REF_CONF = "/root/reference/g2o_frontend/pwn_tracker2/conf"


def test_parse_fixture_and_resolve_pointers():
    objs = B.load(FIXTURE)
    assert [o.cls for o in objs][:2] == ["PinholePointProjector", "StatsCalculatorIntegralImage"]
    p = B.pipeline(objs)
    a = p["align"]
    assert a["outer_iterations"] == 10 and a["inner_iterations"] == 1 and a["robust_kernel"] == 1
    assert a["inlier_max_chi2"] == 9000 and abs(a["inlier_normal_angular_threshold"] - 0.95) < 1e-7
    K = a["projector"]["K"]
    assert K.shape == (3, 3) and K[0, 2] == np.float32(79.875) and K[1, 1] == np.float32(131.25)  # row-major values
    assert (a["projector"]["rows"], a["projector"]["cols"]) == (120, 160)
    s = p["stats"]
    assert (s["min_image_radius"], s["max_image_radius"], s["min_points"]) == (3, 6, 10)
    assert s["flat_omega_p"] == [1000.0, 1.0, 1.0] and s["flat_omega_n"] == [100.0, 100.0, 100.0]
    assert p["merger"]["max_point_depth"] == 10 and p["voxel_resolution"] == 0.02
    # poses are t2v vectors: translation + vector part of a unit quaternion
    T = a["reference_sensor_offset"]
    assert np.allclose(T[:3, 3], [0.05, -0.02, 0.1]) and np.allclose(T[:3, :3] @ T[:3, :3].T, np.eye(3), atol=1e-6)
    Tp = a["projector"]["transform"]
    assert abs(np.degrees(np.arctan2(Tp[1, 0], Tp[0, 0])) - 10.0) < 1e-3   # qz = sin(5 deg) -> 10 deg about z


def test_v2t_matches_the_oracle():
    from oracle import pwn_oracle as O
    rng = np.random.default_rng(0)
    for _ in range(20):
        v = np.concatenate([rng.uniform(-1, 1, 3), rng.uniform(-0.4, 0.4, 3)]).astype(np.float32)
        assert np.allclose(B.v2t(v), O.v2t(v), atol=2e-6)


def test_round_trip():
    objs = B.load(FIXTURE)
    again = B.loads(B.dumps(objs))
    assert [(o.cls, o.fields) for o in objs] == [(o.cls, o.fields) for o in again]
    with pytest.raises(ValueError):
        B.loads('Aligner { "#id": 1 }')          # class name must be quoted
    with pytest.raises(ValueError):
        B.loads('"Aligner" [1, 2]')              # record body must be an object


def test_to_capi_structs():
    capi = pytest.importorskip("g2o_frontend_b200.capi")
    proj, sp, ap, mp = B.to_capi(B.pipeline(B.load(FIXTURE)))
    assert (proj.rows, proj.cols) == (120, 160) and abs(proj.max_distance - 4.5) < 1e-6
    assert abs(proj.K[6] - 79.875) < 1e-6      # column-major in the ABI: K(0,2) is element 6
    assert sp.min_image_radius == 3 and sp.max_image_radius == 6 and list(sp.flat_omega_p) == [1000.0, 1.0, 1.0]
    assert ap.outer_iterations == 10 and ap.robust_kernel == 1 and abs(ap.inlier_max_chi2 - 9000) < 1e-3
    assert abs(mp.normal_threshold - 0.984808) < 1e-6


@pytest.mark.skipif(not os.path.isdir(REF_CONF), reason="reference tree not present (GPU box)")
def test_reference_tracker_configurations_parse():
    files = sorted(glob.glob(os.path.join(REF_CONF, "*.conf")))
    assert files
    seen_aligner = 0
    for f in files:
        objs = B.load(f)
        assert objs, f
        p = B.pipeline(objs)
        if p["align"] is not None:
            seen_aligner += 1
            assert p["align"]["outer_iterations"] >= 1
            assert p["stats"] is None or p["stats"]["min_points"] > 0
    assert seen_aligner >= 3


BIN = os.path.join(ROOT, "g2o_frontend_b200", "lib", "pwn_simple_aligner")


def _cpp_dump(path):
    import json
    import subprocess
    return json.loads(subprocess.check_output([BIN, "--dump-config", path]).decode())


@pytest.mark.skipif(not os.path.exists(BIN), reason="host driver not built")
def test_cpp_reader_agrees_with_python_reader():
    """include/pwn/boss_config.h (what a C++ caller uses) and boss_config.py resolve a file to the same parameters"""
    files = [FIXTURE]
    if os.path.isdir(REF_CONF):
        files += sorted(glob.glob(os.path.join(REF_CONF, "*.conf")))
    checked = 0
    for f in files:
        p = B.pipeline(B.load(f))
        if p["align"] is None or "projector" not in p["align"]:
            continue
        d = _cpp_dump(f)
        a = p["align"]
        assert d["records"] == len(B.load(f))
        assert d["outer_iterations"] == a["outer_iterations"] and d["inner_iterations"] == a["inner_iterations"], f
        assert d["robust_kernel"] == a["robust_kernel"]
        for k in ("inlier_max_chi2", "inlier_distance_threshold", "inlier_normal_angular_threshold",
                  "flat_curvature_threshold", "inlier_curvature_ratio_threshold"):
            assert abs(d[k] - a[k]) <= 1e-6 * max(1.0, abs(a[k])), (f, k)
        q = a["projector"]
        assert np.allclose(np.array(d["K"]).reshape(3, 3), q["K"]) and (d["rows"], d["cols"]) == (q["rows"], q["cols"])
        assert abs(d["max_distance"] - q["max_distance"]) < 1e-6
        assert np.allclose(np.array(d["reference_sensor_offset"]).reshape(4, 4), a["reference_sensor_offset"], atol=2e-6)
        if p["stats"] is not None and "IntegralImage" in "".join(o.cls for o in B.load(f)):
            s = p["stats"]
            assert (d["min_image_radius"], d["max_image_radius"], d["min_points"]) == \
                   (s["min_image_radius"], s["max_image_radius"], s["min_points"]), f
            assert np.allclose(d["flat_omega_p"], s["flat_omega_p"]) and np.allclose(d["flat_omega_n"], s["flat_omega_n"])
        if p["merger"] is not None:
            assert np.allclose(d["merger"], [p["merger"][k] for k in ("distance_threshold", "normal_threshold", "max_point_depth")])
        assert d["has_matcher"] == int(p["matcher"] is not None) and d["has_tracker"] == int(p["tracker"] is not None)
        if p["matcher"] is not None:
            assert d["matcher_scale"] == p["matcher"]["scale"]
            assert abs(d["frame_inlier_depth_threshold"] - p["matcher"]["frame_inlier_depth_threshold"]) < 1e-6
        if p["tracker"] is not None:
            assert abs(d["new_frame_cloud_inliers_fraction"] - p["tracker"]["new_frame_cloud_inliers_fraction"]) < 1e-6
        checked += 1
    assert checked >= 1
    # a malformed file is an error, not a silent default
    import subprocess
    bad = os.path.join(ROOT, "tests", "golden", "boss_pipeline.conf")
    r = subprocess.run([BIN, "--dump-config", bad + ".does_not_exist"], capture_output=True)
    assert r.returncode != 0


def test_acceptance_masks_follow_the_trackers():
    """PwnCloser / PwnTracker acceptance (pwn_closer.cpp:164-171, pwn_tracker.cpp:187-191) over a batch of records"""
    from g2o_frontend_b200 import matcher
    capi = pytest.importorskip("g2o_frontend_b200.capi")
    r = np.zeros(6, capi.RESULT_DTYPE)
    r["image_non_zeros"] = [5000, 2999, 5000, 5000, 5000, 5000]
    r["image_outliers"] = [100, 100, 2001, 100, 100, 100]
    r["image_inliers"] = [4900, 2899, 2999, 999, 4900, 4900]
    r["inliers"] = [20000, 20000, 20000, 20000, 999, 20000]
    r["status"] = [0, 0, 0, 0, 0, 2]
    thr = dict(frame_min_non_zero_threshold=3000, frame_max_outliers_threshold=2000, frame_min_inliers_threshold=1000)
    assert matcher.closer_accept(r, **thr).tolist() == [True, False, False, False, True, False]
    assert matcher.tracker_accept(r, 1000, **thr).tolist() == [True, False, False, False, False, False]
    # boundary values are accepted exactly like the reference's strict comparisons
    e = np.zeros(1, capi.RESULT_DTYPE)
    e["image_non_zeros"], e["image_outliers"], e["image_inliers"], e["inliers"] = 3000, 2000, 1000, 1000
    assert matcher.tracker_accept(e, 1000, **thr).all()
    assert np.array_equal(np.diag(matcher.relation_information("closer")), [100, 100, 100, 1000, 1000, 1000])
    p = B.pipeline(B.loads('"PwnTracker" {"#id": 1, "minCloudInliers": 1000, "newFrameCloudInliersFraction": 0.5, '
                           '"frameMinNonZeroThreshold": 3000, "frameMaxOutliersThreshold": 2000, "frameMinInliersThreshold": 1000}'))
    assert matcher.accept_from_boss(r, p, "tracker").tolist() == [True, False, False, False, False, False]
