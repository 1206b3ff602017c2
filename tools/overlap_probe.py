"""Probe (not part of the product): does running two chunks of a batch concurrently on two streams of one GPU pay?
k_project is bound by HBM / L2 atomics, the fused kernel by issue slots, so their bottlenecks are complementary.
Two contexts (each its own stream and scratch) align half of the pairs each from two host threads."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from g2o_frontend_b200 import capi, synth  # noqa: E402


def main():
    n_cur, n_cand = int(os.environ.get("TUNE_CUR", 8)), int(os.environ.get("TUNE_CAND", 64))
    reps = int(os.environ.get("TUNE_REPS", 3))
    nctx = int(os.environ.get("PROBE_CONTEXTS", 2))
    raws_cur, raws_cand, pairs, guesses = bench.make_workload(n_cur, n_cand, 0)
    C = bench.CONF
    proj = capi.make_projector(synth.K_KINECT, bench.ROWS, bench.COLS, C["minD"], C["maxD"])
    sp = capi.make_stats_params(C["worldRadius"], C["minImageRadius"], C["maxImageRadius"], C["minPoints"],
                                C["curvatureThreshold"], C["omegaCurvatureThreshold"])
    ap = capi.make_align_params(C["inlierDistanceThreshold"], C["inlierNormalAngularThreshold"], C["flatCurvatureThreshold"],
                                C["inlierCurvatureRatioThreshold"], C["inlierMaxChi2"], True, 10, 1)
    ctxs = [capi.Context(0) for _ in range(nctx)]
    work = []
    n = len(pairs)
    for k, ctx in enumerate(ctxs):
        clouds = [ctx.raw_depth_to_cloud(r, proj, sp)[0] for r in raws_cur + raws_cand]
        lo, hi = k * n // nctx, (k + 1) * n // nctx
        refs = [clouds[n_cur + ri] for ri, ci in pairs[lo:hi]]
        curs = [clouds[ci] for ri, ci in pairs[lo:hi]]
        work.append((ctx, refs, curs, guesses[lo:hi]))
    out = [None] * nctx

    def run(k):
        ctx, refs, curs, g = work[k]
        out[k] = ctx.align_batch(refs, curs, proj, ap, g)

    def step():
        th = [threading.Thread(target=run, args=(k,)) for k in range(nctx)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    step()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    dt = (time.perf_counter() - t0) / reps
    inl = sum(int(o["inliers"].sum()) for o in out)
    print("contexts=%d slots=%s pairs=%d: step %.2f ms (%.0f align/s) checksum inliers=%d" %
          (nctx, os.environ.get("NICP_BATCH_SLOTS", "64"), n, dt * 1e3, n / dt, inl))


if __name__ == "__main__":
    main()
