import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


# ---- shared synthetic inputs, parameters of pwn_core/conf/pwn_aligner_1_1.conf ----
CONF_1_1 = dict(minD=0.5, maxD=4.5, minImageRadius=10, maxImageRadius=30, minPoints=50, curvatureThreshold=0.2,
                worldRadius=0.1, omegaCurvatureThreshold=0.02, inlierDistanceThreshold=1.0,
                inlierNormalAngularThreshold=0.95, inlierCurvatureRatioThreshold=1.3, flatCurvatureThreshold=0.02,
                inlierMaxChi2=9000.0, robustKernel=1, outerIterations=10, innerIterations=1)
# pwn_aligner_1_4.conf radii for the down-scaled images
CONF_1_4 = dict(CONF_1_1, minImageRadius=3, maxImageRadius=6, minPoints=10, inlierDistanceThreshold=0.5)


class Scene:
    """Two synthetic frames (A, B) at a given scale plus everything the oracle derives from them."""

    def __init__(self, step, conf, seed=None, dropout=0.0, sensor_offset=None):
        from g2o_frontend_b200 import synth
        from oracle import pwn_oracle as O
        self.conf = conf
        self.step = step
        self.rows, self.cols = 480 // step, 640 // step
        self.K = synth.scaled_K(synth.K_KINECT, 1.0 / step)
        rawA = synth.render_depth_u16(synth.POSE_A, seed=seed, dropout=dropout)
        rawB = synth.render_depth_u16(synth.POSE_B, seed=None if seed is None else seed + 1, dropout=dropout)
        self.rawA, self.rawB = rawA, rawB
        dA, dB = synth.u16_to_m(rawA), synth.u16_to_m(rawB)
        if step > 1:
            dA, dB = O.depth_scale(dA, step), O.depth_scale(dB, step)
        self.depthA, self.depthB = dA, dB
        self.sp = O.default_stats_params(minImageRadius=conf["minImageRadius"], maxImageRadius=conf["maxImageRadius"],
                                         minPoints=conf["minPoints"], curvatureThreshold=conf["curvatureThreshold"],
                                         worldRadius=conf["worldRadius"],
                                         omegaCurvatureThreshold=conf["omegaCurvatureThreshold"])
        self.cp = O.default_corr_params(inlierDistanceThreshold=conf["inlierDistanceThreshold"],
                                        inlierNormalAngularThreshold=conf["inlierNormalAngularThreshold"],
                                        flatCurvatureThreshold=conf["flatCurvatureThreshold"],
                                        inlierCurvatureRatioThreshold=conf["inlierCurvatureRatioThreshold"])
        self.sensor_offset = np.eye(4, dtype=np.float32) if sensor_offset is None else np.asarray(sensor_offset, np.float32)
        self.cloudA, self.indexA, self.intervalA, self.integralA = O.depth_to_cloud(
            dA, self.K, conf["minD"], conf["maxD"], self.sp, self.sensor_offset, want_aux=True)
        self.cloudB, self.indexB = O.depth_to_cloud(dB, self.K, conf["minD"], conf["maxD"], self.sp, self.sensor_offset)
        self.gt = synth.POSE_B.astype(np.float32)

    # ---- GPU-side parameter structs
    def projector(self):
        from g2o_frontend_b200 import capi
        return capi.make_projector(self.K, self.rows, self.cols, self.conf["minD"], self.conf["maxD"])

    def stats_params(self):
        from g2o_frontend_b200 import capi
        c = self.conf
        return capi.make_stats_params(c["worldRadius"], c["minImageRadius"], c["maxImageRadius"], c["minPoints"],
                                      c["curvatureThreshold"], c["omegaCurvatureThreshold"])

    def align_params(self, outer=None, inner=None):
        from g2o_frontend_b200 import capi
        c = self.conf
        return capi.make_align_params(c["inlierDistanceThreshold"], c["inlierNormalAngularThreshold"],
                                      c["flatCurvatureThreshold"], c["inlierCurvatureRatioThreshold"], c["inlierMaxChi2"],
                                      bool(c["robustKernel"]), c["outerIterations"] if outer is None else outer,
                                      c["innerIterations"] if inner is None else inner)

    def oracle_align_params(self, outer=None, inner=None, guess=None, num_threads=8, ref_offset=None, cur_offset=None):
        from oracle import pwn_oracle as O
        c = self.conf
        return O.make_align_params(self.K, self.rows, self.cols, c["minD"], c["maxD"], self.cp,
                                   outer=c["outerIterations"] if outer is None else outer,
                                   inner=c["innerIterations"] if inner is None else inner, guess=guess,
                                   max_chi2=c["inlierMaxChi2"], robust=bool(c["robustKernel"]), num_threads=num_threads,
                                   ref_offset=self.sensor_offset if ref_offset is None else ref_offset,
                                   cur_offset=self.sensor_offset if cur_offset is None else cur_offset)


_SCENES = {}


def get_scene(step=4, seed=None, dropout=0.0, offset=False):
    key = (step, seed, dropout, offset)
    if key not in _SCENES:
        conf = CONF_1_1 if step == 1 else CONF_1_4
        so = None
        if offset:
            from g2o_frontend_b200 import synth
            so = synth.make_pose((0.1, -0.05, 0.3), (0.3, 1.0, 0.2), 7.0)
        _SCENES[key] = Scene(step, conf, seed, dropout, so)
    return _SCENES[key]


@pytest.fixture(scope="session")
def scene_small():
    return get_scene(4)


@pytest.fixture(scope="session")
def scene_full():
    return get_scene(1)
