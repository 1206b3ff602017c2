"""Reader / writer for the BOSS configuration files of the reference's trackers.

`pwn_boss` serialises every `pwn::` object as one record `"ClassName" { ...json... }` with an integer `"#id"`; pointers
between objects are `{ "#pointer" : id }` (g2o_frontend/pwn_boss/*.cpp, e.g. pinholepointprojector.cpp:10-24,
aligner.cpp:12-47; the files live in g2o_frontend/pwn_tracker2/conf/).  Eigen matrices are `{ "values" : [...] }` in
ROW-major order (boss_map/eigen_boss_plugin.hpp:1-33), poses are the 6-vectors of `t2v` (translation + the vector part
of the unit quaternion).  This module parses such a file and turns the objects the NICP path needs into the C-ABI
parameter structs of `g2o_frontend_b200.capi`, so a deployment can keep its existing configuration files.
"""
import json
import re

import numpy as np


class BossObject:
    def __init__(self, cls, fields):
        self.cls = cls
        self.fields = fields
        self.id = fields.get("#id")

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default)

    def pointer(self, k):
        """id a pointer field refers to, or None"""
        v = self.fields.get(k)
        if isinstance(v, dict) and "#pointer" in v:
            return v["#pointer"]
        return None

    def matrix(self, k, rows, cols):
        v = np.asarray(self.fields[k]["values"], np.float32)
        if v.size != rows * cols:
            raise ValueError("%s.%s: expected %d values, found %d" % (self.cls, k, rows * cols, v.size))
        return v.reshape(rows, cols)

    def __repr__(self):
        return "BossObject(%s #%s)" % (self.cls, self.id)


def loads(text):
    """parse the records of a BOSS file -> list of BossObject in file order"""
    # BOSS's own number parser accepts hand-edited integers such as `000` (pwn_slam_gui_short.conf); JSON does not
    text = re.sub(r"(?<![\w.])0+(?=\d)", "", text)
    dec = json.JSONDecoder()
    out, i, n = [], 0, len(text)
    while True:
        while i < n and text[i] in " \t\r\n":
            i += 1
        if i >= n:
            break
        if text[i] != '"':
            raise ValueError("BOSS record must start with a quoted class name (offset %d)" % i)
        cls, i = dec.raw_decode(text, i)
        while i < n and text[i] in " \t\r\n":
            i += 1
        fields, i = dec.raw_decode(text, i)
        if not isinstance(cls, str) or not isinstance(fields, dict):
            raise ValueError("malformed BOSS record near offset %d" % i)
        out.append(BossObject(cls, fields))
    return out


def load(path):
    with open(path) as f:
        return loads(f.read())


def dumps(objects):
    """inverse of loads (one record per line, like the reference's serializer)"""
    return "".join('"%s" %s\n' % (o.cls, json.dumps(o.fields)) for o in objects)


def by_id(objects):
    return {o.id: o for o in objects if o.id is not None}


def v2t(v):
    """6-vector (t, qx, qy, qz) -> 4x4 (bm_se3.h:9-23), float64 on the host (configuration time)"""
    v = np.asarray(v, np.float64)
    q = v[3:6]
    n2 = float(q @ q)
    w = np.sqrt(max(0.0, 1.0 - n2)) if n2 <= 1.0 else 0.0
    if n2 > 1.0:
        q = q / np.sqrt(n2)
    x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = v[:3]
    return T.astype(np.float32)


# ---- objects -> parameter dictionaries (plain numbers / numpy, no GPU needed) -------------------------------------
def projector_params(o):
    """PinholePointProjector record -> dict (pwn_boss/pointprojector.cpp:13-35, pinholepointprojector.cpp:10-24)"""
    if o.cls != "PinholePointProjector":
        raise ValueError("not a PinholePointProjector: %r" % o)
    return {"K": o.matrix("cameraMatrix", 3, 3), "rows": int(o["imageRows"]), "cols": int(o["imageCols"]),
            "min_distance": float(o["minDistance"]), "max_distance": float(o["maxDistance"]),
            "baseline": float(o.get("baseline", 0.075)), "alpha": float(o.get("alpha", 0.1)),
            "transform": v2t(o["transform"]["values"])}


def multi_projector_params(o, table):
    """MultiPointProjector record -> list of (projector dict, sensor offset 4x4) (pwn_boss/multipointprojector.cpp)"""
    out = []
    for child in o.get("childProjectors", []):
        pid = child["projector"]["#pointer"]
        out.append((projector_params(table[pid]), v2t(child["sensorOffset"]["values"])))
    return out


def stats_params(stats, point_info=None, normal_info=None):
    """StatsCalculatorIntegralImage (+ the two information-matrix calculators) -> dict with the C-ABI field names.
    NB the reference's files store imageMinRadius / imageMaxRadius exactly as the setters received them."""
    d = {"world_radius": float(stats["worldRadius"]), "min_image_radius": int(stats["imageMinRadius"]),
         "max_image_radius": int(stats["imageMaxRadius"]), "min_points": int(stats["minPoints"]),
         "curvature_threshold": float(stats["curvatureThreshold"]), "omega_curvature_threshold": 0.02,
         "flat_omega_p": [1000.0, 1.0, 1.0], "flat_omega_n": [100.0, 100.0, 100.0], "nonflat_omega_n": [1.0, 1.0, 1.0]}
    if point_info is not None:
        d["flat_omega_p"] = np.diag(point_info.matrix("flatInformationMatrix", 4, 4))[:3].tolist()
    if normal_info is not None:
        d["flat_omega_n"] = np.diag(normal_info.matrix("flatInformationMatrix", 4, 4))[:3].tolist()
        d["nonflat_omega_n"] = np.diag(normal_info.matrix("nonflatInformationMatrix", 4, 4))[:3].tolist()
    return d


def align_params(aligner, table):
    """Aligner record (+ its Linearizer / CorrespondenceFinder / projector) -> dict"""
    lin = table[aligner.pointer("linearizer")]
    cf = table[aligner.pointer("correspondenceFinder")]
    d = {"inlier_distance_threshold": float(cf["inlierDistanceThreshold"]),
         "inlier_normal_angular_threshold": float(cf["inlierNormalAngularThreshold"]),
         "flat_curvature_threshold": float(cf["flatCurvatureThreshold"]),
         "inlier_curvature_ratio_threshold": float(cf["inlierCurvatureRatioThreshold"]),
         "inlier_max_chi2": float(lin["inlierMaxChi2"]), "robust_kernel": int(bool(lin["robustKernel"])),
         "outer_iterations": int(aligner["outerIterations"]), "inner_iterations": int(aligner["innerIterations"]),
         "reference_sensor_offset": v2t(aligner["referenceSensorOffset"]["values"]),
         "current_sensor_offset": v2t(aligner["currentSensorOffset"]["values"])}
    pid = aligner.pointer("projector")
    if pid is not None and table[pid].cls == "PinholePointProjector":
        d["projector"] = projector_params(table[pid])
    return d


def merger_params(o):
    return {"distance_threshold": float(o["distanceThreshold"]), "normal_threshold": float(o["normalThreshold"]),
            "max_point_depth": float(o["maxPointDepth"])}


def pipeline(objects):
    """the first Aligner and the first DepthImageConverterIntegralImage of a file, resolved:
    -> {"align", "converter_projector", "stats", "merger", "voxel_resolution", "matcher", "tracker"} (None when absent)"""
    table = by_id(objects)
    first = lambda cls: next((o for o in objects if o.cls == cls), None)
    out = {"align": None, "converter_projector": None, "stats": None, "merger": None, "voxel_resolution": None,
           "matcher": None, "tracker": None}
    al = first("Aligner")
    if al is not None:
        out["align"] = align_params(al, table)
    conv = first("DepthImageConverterIntegralImage") or first("DepthImageConverter")
    if conv is not None:
        pp = table[conv.pointer("pointProjector")]
        if pp.cls == "PinholePointProjector":
            out["converter_projector"] = projector_params(pp)
        elif pp.cls == "MultiPointProjector":
            out["converter_projector"] = multi_projector_params(pp, table)
        out["stats"] = stats_params(table[conv.pointer("statsCalculator")], table.get(conv.pointer("pointInfoCalculator")),
                                    table.get(conv.pointer("normalInfoCalculator")))
    mg = first("Merger")
    if mg is not None:
        out["merger"] = merger_params(mg)
    vx = first("VoxelCalculator")
    if vx is not None:
        out["voxel_resolution"] = float(vx["resolution"])
    mt = first("PwnMatcherBase")  # pwn_tracker2/pwn_matcher_base.cpp:20-36
    if mt is not None:
        out["matcher"] = {"scale": int(mt["scale"]), "frame_inlier_depth_threshold": float(mt["frameInlierDepthThreshold"])}
    tr = first("PwnTracker")      # pwn_tracker2/pwn_tracker.cpp serialize
    if tr is not None:
        out["tracker"] = {"new_frame_cloud_inliers_fraction": float(tr.get("newFrameCloudInliersFraction", 0.4)),
                          "min_cloud_inliers": int(tr.get("minCloudInliers", 0)),
                          "frame_min_non_zero_threshold": int(tr.get("frameMinNonZeroThreshold", 0)),
                          "frame_max_outliers_threshold": int(tr.get("frameMaxOutliersThreshold", 0)),
                          "frame_min_inliers_threshold": int(tr.get("frameMinInliersThreshold", 0))}
    return out


# ---- dictionaries -> C-ABI structs (imports capi lazily: needs the built library only when used) ----------------------
def to_capi(p):
    """pipeline() dictionary -> (capi.Projector of the aligner, capi.StatsParams, capi.AlignParams, capi.MergeParams|None)"""
    from . import capi
    proj = sp = ap = mp = None
    a = p.get("align")
    if a is not None:
        ap = capi.make_align_params(a["inlier_distance_threshold"], a["inlier_normal_angular_threshold"],
                                    a["flat_curvature_threshold"], a["inlier_curvature_ratio_threshold"],
                                    a["inlier_max_chi2"], bool(a["robust_kernel"]), a["outer_iterations"],
                                    a["inner_iterations"])
        if "projector" in a:
            q = a["projector"]
            proj = capi.make_projector(q["K"], q["rows"], q["cols"], q["min_distance"], q["max_distance"])
    s = p.get("stats")
    if s is not None:
        sp = capi.make_stats_params(s["world_radius"], s["min_image_radius"], s["max_image_radius"], s["min_points"],
                                    s["curvature_threshold"], s["omega_curvature_threshold"])
        sp.flat_omega_p[:] = s["flat_omega_p"]
        sp.flat_omega_n[:] = s["flat_omega_n"]
        sp.nonflat_omega_n[:] = s["nonflat_omega_n"]
    m = p.get("merger")
    if m is not None:
        mp = capi.make_merge_params(m["distance_threshold"], m["normal_threshold"], m["max_point_depth"])
    return proj, sp, ap, mp
