"""Acceptance of alignment results, as the reference's trackers apply it to `PwnMatcherBase::MatcherResult`
(the fields are part of the 256-byte `nicp_align_result` record, so a whole batch is decided with a few numpy masks).

  PwnCloser::registerNodes   pwn_tracker2/pwn_closer.cpp:164-171   loop-closure candidate: image statistics only
  PwnTracker::registerNodes  pwn_tracker2/pwn_tracker.cpp:187-191  odometry step: also the inlier count of the aligner

Both then attach constant information matrices to the accepted relation (pwn_closer.cpp:177-180:
diag(100 x3, 1000 x3); pwn_tracker.cpp:197-200: diag(10 x3, 100 x3)) -- `Aligner::omega()` is not used there.
"""
import numpy as np


def closer_accept(records, frame_min_non_zero_threshold, frame_max_outliers_threshold, frame_min_inliers_threshold):
    """boolean mask over RESULT_DTYPE records: the candidate becomes a PwnCloserRelation"""
    r = records
    return ~((r["image_non_zeros"] < frame_min_non_zero_threshold) |
             (r["image_outliers"] > frame_max_outliers_threshold) |
             (r["image_inliers"] < frame_min_inliers_threshold)) & (r["status"] == 0)


def tracker_accept(records, min_cloud_inliers, frame_min_non_zero_threshold, frame_max_outliers_threshold,
                   frame_min_inliers_threshold):
    """boolean mask: the step becomes a PwnTrackerRelation (cloud_inliers = Aligner::inliers())"""
    return closer_accept(records, frame_min_non_zero_threshold, frame_max_outliers_threshold,
                         frame_min_inliers_threshold) & ~(records["inliers"] < min_cloud_inliers)


def relation_information(kind):
    """the constant 6x6 information matrix the reference attaches to an accepted relation"""
    t, r = (100.0, 1000.0) if kind == "closer" else (10.0, 100.0)
    return np.diag([t, t, t, r, r, r])


def accept_from_boss(records, pipeline, kind="closer"):
    """thresholds taken from the PwnTracker record of a BOSS file (boss_config.pipeline)"""
    t = pipeline.get("tracker")
    if t is None:
        raise ValueError("the configuration holds no PwnTracker record")
    if kind == "closer":
        return closer_accept(records, t["frame_min_non_zero_threshold"], t["frame_max_outliers_threshold"],
                             t["frame_min_inliers_threshold"])
    return tracker_accept(records, t["min_cloud_inliers"], t["frame_min_non_zero_threshold"],
                          t["frame_max_outliers_threshold"], t["frame_min_inliers_threshold"])
