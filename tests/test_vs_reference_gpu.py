"""GPU tests whose checker is the REFERENCE'S OWN code rather than the oracle: oracle/_ref/libpwn_core_ref.so =
g2o_frontend/pwn_core/*.cpp compiled (oracle/build_ref_pwn_core.sh) against the Eigen / OpenCV stand-ins of oracle/shim.
tests/test_reference_pwn_core.py shows on the CPU that the oracle is bit-identical to it, so these restate a few of the
parity tests of tests/test_gpu_parity.py with the middle man removed: the CUDA path (through the C-ABI) on one side, the
reference's sources on the other.  Same protocol and tolerances (BASELINE.json north_star): stage-isolated /
teacher-forced, index and correspondence images bit-exact, H and b within 1e-4, T within 1e-4 rad / 1e-4 m."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import get_scene
from test_reference_pwn_core import REF_SO, RefCloud, cm, fp, ip, run_ref_align
from test_gpu_parity import frob_rel, rot_angle, H_RTOL, T_ROT_TOL, T_TRA_TOL

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/libpwn_core_ref.so not built")]

SYM = [0, 4, 8, 5, 9, 10]  # xx xy xz yy yz zz of a column-major 4x4


@pytest.fixture(scope="module")
def R():
    from oracle import pwn_oracle as O
    O.lib()
    L = C.CDLL(REF_SO)
    L.refcore_depth_to_cloud.restype = C.c_void_p
    L.refcore_set_threads(1)
    return L


@pytest.fixture(scope="module", params=["verify", "default"])
def ctx(request):
    from g2o_frontend_b200 import capi
    c = capi.Context(0, verify=(request.param == "verify"))
    yield c
    c.close()


def upload(ctx, rc):
    """cloud of the reference -> device cloud"""
    cl = ctx.new_cloud(max(rc.n, 1))
    cl.upload(rc.points, rc.normals, rc.curvature, np.ascontiguousarray(rc.omegaP[:, SYM]), np.ascontiguousarray(rc.omegaN[:, SYM]))
    return cl


@pytest.mark.parametrize("step,seed,dropout,offset", [(4, None, 0.0, False), (4, 0, 0.05, True), (1, None, 0.0, False)])
def test_frame_prep_against_the_reference_sources(ctx, R, step, seed, dropout, offset):
    """depth image -> cloud: points, index image, interval image, the 10-channel integral image and Stats::n bit-exact;
    normals / curvature within the north_star tolerances"""
    s = get_scene(step, seed, dropout, offset)
    rc = RefCloud(R, s.depthA, s.K, s.conf, s.sensor_offset)
    cl, idx = ctx.depth_to_cloud(s.depthA, s.projector(), s.stats_params(), s.sensor_offset, keep_stats=True)
    assert cl.size() == rc.n
    assert np.array_equal(idx, rc.index)
    assert np.array_equal(ctx.last_interval_image(s.rows, s.cols), rc.interval)
    I = ctx.last_integral_image(s.rows, s.cols)
    assert np.array_equal(np.asarray(I, np.float32).reshape(-1).view(np.uint32), rc.integral.reshape(-1).view(np.uint32))
    d = cl.download()
    assert np.array_equal(d["points"].view(np.uint32), rc.points.view(np.uint32))
    _, _, cnt = cl.download_stats()
    assert np.array_equal(cnt, rc.statsN)
    has_r = np.abs(rc.normals[:, :3]).sum(1) > 0
    has_g = np.abs(d["normals"][:, :3]).sum(1) > 0
    assert (has_r != has_g).mean() < 1e-4
    both = has_r & has_g
    well = both & ((rc.eigvals[:, 1] - rc.eigvals[:, 0]) > 1e-4 * rc.eigvals[:, 2])
    ang = np.arccos(np.clip((d["normals"][well, :3].astype(np.float64) * rc.normals[well, :3]).sum(1), -1, 1))
    assert well.sum() > 0.5 * rc.n and np.quantile(ang, 0.999) <= 1e-3
    cerr = np.abs(d["curvature"][both] - rc.curvature[both]) / np.maximum(np.abs(rc.curvature[both]), 1e-3)
    assert np.quantile(cerr, 0.999) <= 1e-4


def test_projection_against_the_reference_sources(ctx, R):
    """PinholePointProjector::project: index + depth image bit-exact at three projector poses"""
    from oracle import pwn_oracle as O
    s = get_scene(4)
    c = s.conf
    rc = RefCloud(R, s.depthA, s.K, c)
    cl = upload(ctx, rc)
    for v in ([0, 0, 0, 0, 0, 0], [0.03, -0.02, 0.05, 0.01, -0.015, 0.005], [-0.2, 0.1, 0.3, -0.05, 0.08, 0.02]):
        T = O.v2t(np.array(v, np.float32))
        ii = np.zeros((s.rows, s.cols), np.int32)
        dd = np.zeros((s.rows, s.cols), np.float32)
        R.refcore_project(rc.h, fp(cm(s.K)), fp(cm(T)), s.rows, s.cols, C.c_float(c["minD"]), C.c_float(c["maxD"]), ip(ii), fp(dd))
        KRt, _ = O.update_matrices(s.K, T)
        gi, gd = ctx.project(cl, KRt, s.rows, s.cols, c["minD"], c["maxD"])
        assert np.array_equal(gi, ii)
        assert np.array_equal(gd.view(np.uint32), dd.view(np.uint32))


def test_alignment_against_the_reference_sources(ctx, R):
    """Aligner::align: teacher-forced single iterations (index image, correspondences bit-exact, H / b within 1e-4) and
    the free-running 10-iteration result (T within 1e-4 rad / 1e-4 m) against the reference's own Aligner"""
    from g2o_frontend_b200 import capi
    s = get_scene(4, 0, 0.05)
    rA, rB = RefCloud(R, s.depthA, s.K, s.conf), RefCloud(R, s.depthB, s.K, s.conf)
    ref, cur = upload(ctx, rA), upload(ctx, rB)
    full = run_ref_align(R, rA, rB, s)
    res = ctx.align(ref, cur, s.projector(), s.align_params())
    T = capi.result_T(res)
    assert rot_angle(T[:3, :3], full["T"][:3, :3]) <= T_ROT_TOL
    assert np.abs(T[:3, 3] - full["T"][:3, 3]).max() <= T_TRA_TOL
    assert abs(res.inliers - full["inliers"]) <= 1e-3 * full["inliers"] + 8
    # teacher-forced: iteration k restarted from the reference's T after k iterations
    for k in (0, 3, 9):
        Tk = np.eye(4, dtype=np.float32) if k == 0 else run_ref_align(R, rA, rB, s, outer=k)["T"]
        one = run_ref_align(R, rA, rB, s, outer=1, guess=Tk)
        r1 = ctx.align(ref, cur, s.projector(), s.align_params(outer=1), guess=Tk)
        st = ctx.align_state(s.rows, s.cols)
        assert np.array_equal(st["ref_index"], one["refIndex"])
        assert np.array_equal(st["ref_depth"].view(np.uint32), one["refDepth"].view(np.uint32))
        assert np.array_equal(st["cur_index"], one["curIndex"])
        assert r1.num_correspondences == one["n"]
        assert np.array_equal(st["corr"], one["corr"][:one["n"]])
        assert r1.inliers == one["inliers"]
        assert abs(r1.error - one["error"]) <= 1e-3 * abs(one["error"])  # the reference sums chi2 sequentially in float32
        T1 = capi.result_T(r1)
        assert rot_angle(T1[:3, :3], one["T"][:3, :3]) <= T_ROT_TOL
        assert np.abs(T1[:3, 3] - one["T"][:3, 3]).max() <= T_TRA_TOL


def test_drop_in_demo_reference_classes_with_the_b200_backend(tmp_path):
    """The drop-in claim executed: oracle/_ref/drop_in_demo is ONE program written against the reference's own classes
    (built from /root/reference + integration/pwn_b200/b200_pwn.h); it runs converter.compute x2 + aligner.align through
    pwn::DepthImageConverter* / pwn::Aligner* once with the reference's CPU objects and once with the B200 subclasses."""
    from test_reference_pwn_core import DEMO, run_drop_in_demo
    if not os.path.exists(DEMO):
        pytest.skip("oracle/_ref/drop_in_demo not built")
    for step in (4, 1):
        s = get_scene(step)
        rc, out, err = run_drop_in_demo(s, "both", tmp_path)
        assert rc == 0, (out, err)
        cpu, gpu = out["reference_cpu"], out["b200"]
        Tc, Tg = np.array(cpu["T"]).reshape(4, 4), np.array(gpu["T"]).reshape(4, 4)
        assert rot_angle(Tg[:3, :3], Tc[:3, :3]) <= T_ROT_TOL
        assert np.abs(Tg[:3, 3] - Tc[:3, 3]).max() <= T_TRA_TOL
        assert gpu["reference_points"] == cpu["reference_points"] and gpu["current_points"] == cpu["current_points"]
        assert abs(gpu["num_correspondences"] - cpu["num_correspondences"]) <= 1e-3 * cpu["num_correspondences"] + 1
        assert abs(gpu["inliers"] - cpu["inliers"]) <= 1e-3 * cpu["inliers"] + 1
        assert abs(gpu["reference_pixels"] - cpu["reference_pixels"]) <= 1e-2 * cpu["reference_pixels"]
        # the Gaussian3f sensor model the converter leaves in cloud.gaussians() (what a CPU Merger reads)
        assert gpu["gaussians"] == cpu["gaussians"] == cpu["reference_points"]
        assert abs(gpu["gaussian_sum"] - cpu["gaussian_sum"]) <= 1e-6 * cpu["gaussian_sum"]
        print("drop-in demo %dx%d: B200 align %.3f ms (second call), |dT| %.2e" %
              (s.rows, s.cols, out["b200_second_align_ms"], float(np.abs(Tg - Tc).max())))
    # with SE(3) priors added through the reference's own Aligner::addRelativePrior / addAbsolutePrior
    s = get_scene(4)
    rc, out, err = run_drop_in_demo(s, "both", tmp_path, priors=True)
    assert rc == 0, (out, err)
    Tp, Tq = np.array(out["reference_cpu"]["T"]).reshape(4, 4), np.array(out["b200"]["T"]).reshape(4, 4)
    assert rot_angle(Tq[:3, :3], Tp[:3, :3]) <= 2e-4 and np.abs(Tq[:3, 3] - Tp[:3, 3]).max() <= 2e-4


def test_cli_driver_against_the_reference_cli_driver(tmp_path):
    """BASELINE configs[0]: the reference's own pwn_simple_aligner (oracle/_ref/pwn_simple_aligner_ref, compiled unmodified
    from /root/reference) and this repository's driver (pwn:: classes -> C-ABI -> CUDA) on the same 16-bit PGM frames and
    the same configuration: the trajectories agree."""
    import json
    import subprocess
    from conftest import CONF_1_1, CONF_1_4
    from g2o_frontend_b200 import synth
    from test_host_cpp import BIN, write_conf, write_pgm16
    from test_reference_pwn_core import REF_CLI, run_reference_cli
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/pwn_simple_aligner_ref not built")
    poses = [synth.POSE_A, synth.POSE_B, synth.POSE_B @ synth.make_pose((-0.02, 0.01, 0.03), (1.0, 0.3, 0.2), 1.5)]
    raws = [synth.render_depth_u16(p, seed=3 + i) for i, p in enumerate(poses)]
    # the second case starts the trajectory at a non-identity initial pose (tx .. qw of the configuration file)
    start = dict(tx=0.4, ty=-0.1, tz=0.25, qx=0.05, qy=-0.1, qz=0.02, qw=0.9935290634701167)
    for image_scale, conf, initial in ((4, CONF_1_4, None), (4, CONF_1_4, start), (1, CONF_1_1, None)):
        ref = run_reference_cli(tmp_path, raws, conf, image_scale, initial=initial)
        files = []
        for i, r in enumerate(raws):
            p = str(tmp_path / ("depth%d.pgm" % i))
            write_pgm16(p, r)
            files.append(p)
        cfg = str(tmp_path / "aligner.conf")
        write_conf(cfg, conf, image_scale, [0, 0, 0, 0, 0, 0], extra=initial)
        out = str(tmp_path / "out.jsonl")
        subprocess.check_call([BIN, cfg, out] + files)
        lines = [json.loads(l) for l in open(out)]
        G = ref[0].copy()  # both start from the configured initial pose (identity when none is given)
        if initial is None:
            assert np.abs(G - np.eye(4)).max() < 1e-6
        else:
            assert np.abs(G[:3, 3] - (0.4, -0.1, 0.25)).max() < 1e-6
            from oracle import pwn_oracle as O
            g = np.array(lines[2]["global"], np.float64)  # the driver's own global pose (t, qx, qy, qz) after the last frame
            assert np.abs(O.v2t(g.astype(np.float32)) - ref[2]).max() <= 5e-4
        for i in (1, 2):
            G = G @ np.array(lines[i]["T"], np.float64).reshape(4, 4).T
            assert rot_angle(G[:3, :3], ref[i][:3, :3]) <= 3e-4, (image_scale, i)
            assert np.abs(G[:3, 3] - ref[i][:3, 3]).max() <= 3e-4, (image_scale, i)


def test_scene_odometry_driver_against_the_reference_driver(tmp_path):
    """pwn_core/pwn_aligner.cpp (scene-based odometry: align against the re-rendered local map, Cloud::add, Merger::merge,
    a new map after the first alignment and every chunkStep frames), the reference's own driver compiled unmodified
    (oracle/_ref/pwn_aligner_ref) against this repository's driver in its `localmap 1` mode, same PGM frames, same
    configuration, the reference's default chunkStep."""
    import json
    import subprocess
    from conftest import CONF_1_4
    from test_host_cpp import BIN, write_conf, write_pgm16
    from test_reference_pwn_core import REF_MAP_CLI, map_sequence, run_reference_map_cli
    if not os.path.exists(REF_MAP_CLI):
        pytest.skip("oracle/_ref/pwn_aligner_ref not built")
    gt, raws = map_sequence(6)
    files = []
    for i, r in enumerate(raws):
        p = str(tmp_path / ("m%d.pgm" % i))
        write_pgm16(p, r)
        files.append(p)
    # identity start, and a trajectory (and first local map) that starts at a configured initial pose
    start = dict(tx=0.4, ty=-0.1, tz=0.25, qx=0.05, qy=-0.1, qz=0.02, qw=0.9935290634701167)
    for initial in (None, start):
        ref = run_reference_map_cli(tmp_path, raws, CONF_1_4, 4, None, initial=initial)
        cfg = str(tmp_path / "map.conf")
        write_conf(cfg, CONF_1_4, 4, [0, 0, 0, 0, 0, 0], extra=dict(initial or {}, localmap=1))
        out = str(tmp_path / "map.jsonl")
        subprocess.check_call([BIN, cfg, out] + files)
        lines = [json.loads(l) for l in open(out)]
        assert [l["new_map"] for l in lines[:6]] == [0, 1, 0, 0, 0, 0]
        for i in range(6):
            G = np.array(lines[i]["globalT"], np.float64).reshape(4, 4).T
            # free-running through Merger::merge: a pose difference of 1e-6 moves points across pixel borders of the
            # rendered map, so the trajectories drift apart by a few 1e-4 over the six frames
            assert np.abs(G - ref[i]).max() <= 3e-3, (initial is not None, i, G, ref[i])


def test_same_command_line_as_the_reference_driver(tmp_path):
    """Drop-in at the command line: `pwn_simple_aligner config list odometry`, the reference's own three arguments, with
    the reference's binary (oracle/_ref/pwn_simple_aligner_ref) and with this repository's driver on the GPU: same odometry
    file (numbers within the alignment tolerance; byte-identical through the test double on the CPU), same clouds in the
    .pwn files left next to the frames (points and Stats::n bit-identical)."""
    from test_reference_pwn_core import REF_CLI, pwn_payload, run_both_drivers_with_the_reference_command_line
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/pwn_simple_aligner_ref not built")
    out = run_both_drivers_with_the_reference_command_line(tmp_path, 4, None)
    ours, ref = out["ours"][0].splitlines(), out["reference"][0].splitlines()
    assert len(ours) == len(ref) == 3
    for a, b in zip(ours, ref):
        fa, fb = a.split(), b.split()
        assert fa[0] == fb[0]  # timestamp
        va, vb = np.array(fa[1:], np.float64), np.array(fb[1:], np.float64)
        assert np.abs(va - vb).max() <= 3e-4, (a, b)
    for a, b in zip(out["ours"][1], out["reference"][1]):
        pa, pb = pwn_payload(a), pwn_payload(b)
        assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[3], pb[3])
        has_a, has_b = np.abs(pa[1][:, :3]).sum(1) > 0, np.abs(pb[1][:, :3]).sum(1) > 0
        assert (has_a != has_b).mean() < 1e-4
