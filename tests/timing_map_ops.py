"""Timing harness (test infrastructure, not part of the product; lives under tests/ because it runs the oracle): local-map maintenance at 640x480 -- gaussians, Cloud::add, Merger::merge,
VoxelCalculator -- on a two-frame map, GPU (C-ABI, host-synchronised calls) vs the CPU oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import get_scene  # noqa: E402
from g2o_frontend_b200 import capi  # noqa: E402
from oracle import pwn_oracle as O  # noqa: E402


def timed(fn, reps=5):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    return (time.perf_counter() - t0) / reps * 1e3, r


def main():
    s = get_scene(1, None, 0.0, False)
    c = s.conf
    ctx = capi.Context(0)
    proj, sp = s.projector(), s.stats_params()
    a, _ = ctx.depth_to_cloud(s.depthA, proj, sp, s.sensor_offset)
    b, _ = ctx.depth_to_cloud(s.depthB, proj, sp, s.sensor_offset)
    t_g, _ = timed(lambda: a.compute_gaussians(s.depthA, proj, 0.075, 0.1, s.sensor_offset))
    b.compute_gaussians(s.depthB, proj, 0.075, 0.1, s.sensor_offset)
    n = a.size() + b.size()

    def build():
        m = ctx.new_cloud(n)
        m.append(a)
        m.append(b, s.gt)
        ctx.synchronize()
        return m
    t_add, _ = timed(build)
    pre = ctx.new_cloud(n)
    pre.append(a)
    pre.append(b, s.gt)
    ctx.synchronize()
    # append into an existing allocation: time the kernels only (destination reset by re-creating the count is not
    # exposed, so time a transform of the whole map instead, which touches the same arrays)
    t_tr, _ = timed(lambda: (pre.transform(s.gt), ctx.synchronize()))

    def merge():
        m = build()
        t0 = time.perf_counter()
        k, _ = m.merge(proj)
        dt = time.perf_counter() - t0
        m.close()
        return dt, k
    merge()
    res = [merge() for _ in range(5)]
    t_merge = np.mean([r[0] for r in res]) * 1e3

    def voxel():
        m = build()
        t0 = time.perf_counter()
        k, _ = m.voxelize(0.01)
        dt = time.perf_counter() - t0
        m.close()
        return dt, k
    voxel()
    resv = [voxel() for _ in range(5)]
    t_vox = np.mean([r[0] for r in resv]) * 1e3
    print("GPU  640x480 two-frame map (%d points): gaussians %.2f ms | allocate + Cloud::add x2 %.2f ms | transformInPlace of the map %.2f ms | Merger::merge %.2f ms "
          "(-> %d points) | VoxelCalculator 1 cm %.2f ms (-> %d points)" % (n, t_g, t_add, t_tr, t_merge, res[0][1], t_vox, resv[0][1]))

    # CPU oracle on the same map
    gA, fA, _, _ = O.unproject_gaussians(s.depthA, s.K, c["minD"], c["maxD"])
    t0 = time.perf_counter()
    O.unproject_gaussians(s.depthA, s.K, c["minD"], c["maxD"])
    t_cg = (time.perf_counter() - t0) * 1e3
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_map_ops import two_frame_map
    m, g, f = two_frame_map(s)
    t0 = time.perf_counter()
    r = O.merge(m, g, f, s.rows, s.cols, s.K, np.eye(4, dtype=np.float32), c["minD"], c["maxD"])
    t_cm = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter()
    v = O.voxelize(m.points, 0.01, True)
    t_cv = (time.perf_counter() - t0) * 1e3
    print("CPU oracle (1 thread, incl. numpy copies): gaussians %.1f ms | Merger::merge %.1f ms (-> %d) | VoxelCalculator %.1f ms (-> %d)"
          % (t_cg, t_cm, r[0].n, t_cv, len(v)))
    ctx.close()


if __name__ == "__main__":
    main()
