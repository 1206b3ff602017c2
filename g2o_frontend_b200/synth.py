"""Deterministic synthetic depth frames.

Ray-cast scene rendered in float64 and quantised like a Kinect: depth = camera-z in metres,
stored as uint16 millimetres (0 = no return).  Camera convention: x right, y down, z forward;
pixel (r, c) looks along K^-1 (c, r, 1).  Scene: closed room (walls x=-2.5, x=2.8, floor y=1.2,
ceiling y=-1.6, back wall z=3.5, front wall z=-1.0) + sphere + axis-aligned box.

The reference ships no depth sequences (datasets/2D only), so every input of the parity tests
and the bench is generated here.
"""
import numpy as np

K_KINECT = np.array([[525.0, 0.0, 319.5], [0.0, 525.0, 239.5], [0.0, 0.0, 1.0]], dtype=np.float32)

_PLANES = [  # (normal, offset): n.x = c
    ((1.0, 0.0, 0.0), -2.5), ((1.0, 0.0, 0.0), 2.8),
    ((0.0, 1.0, 0.0), 1.2), ((0.0, 1.0, 0.0), -1.6),
    ((0.0, 0.0, 1.0), 3.5), ((0.0, 0.0, 1.0), -1.0),
]
_SPHERE = ((0.4, 0.2, 2.2), 0.5)
_BOX = ((-1.3, 0.3, 1.9), (-0.6, 1.2, 2.7))


def scaled_K(K, f):
    """PinholePointProjector::scale (pinholepointprojector.cpp:149-154): first two rows times f."""
    K = np.array(K, dtype=np.float32).copy()
    K[:2, :] *= np.float32(f)
    return K


def axis_angle(axis, angle_rad):
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    Kx = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + np.sin(angle_rad) * Kx + (1 - np.cos(angle_rad)) * (Kx @ Kx)


def make_pose(t=(0, 0, 0), axis=(0, 1, 0), angle_deg=0.0):
    T = np.eye(4)
    T[:3, :3] = axis_angle(axis, np.deg2rad(angle_deg))
    T[:3, 3] = t
    return T


POSE_A = make_pose()
POSE_B = make_pose((0.03, -0.02, 0.05), (0.2, 1.0, 0.1), 2.0)  # B = A o delta


def render_depth_m(pose, rows=480, cols=640, K=K_KINECT, zmin=0.5, zmax=4.5):
    """float64 camera-z per pixel (0 where the nearest hit is outside [zmin, zmax])."""
    K = np.asarray(K, np.float64)
    R, o = pose[:3, :3], pose[:3, 3]
    c, r = np.meshgrid(np.arange(cols, dtype=np.float64), np.arange(rows, dtype=np.float64))
    dc = np.stack([(c - K[0, 2]) / K[0, 0], (r - K[1, 2]) / K[1, 1], np.ones_like(c)], axis=-1)
    d = dc @ R.T  # world directions, ray parameter == camera z
    best = np.full((rows, cols), np.inf)
    for n, off in _PLANES:
        n = np.asarray(n)
        den = d @ n
        with np.errstate(divide="ignore", invalid="ignore"):
            t = (off - o @ n) / den
        t = np.where((t > 1e-6) & np.isfinite(t), t, np.inf)
        best = np.minimum(best, t)
    ctr, rad = np.asarray(_SPHERE[0]), _SPHERE[1]
    oc = o - ctr
    a = np.sum(d * d, -1)
    b = 2 * (d @ oc)
    cc = oc @ oc - rad * rad
    disc = b * b - 4 * a * cc
    with np.errstate(invalid="ignore"):
        t = (-b - np.sqrt(disc)) / (2 * a)
    t = np.where((disc > 0) & (t > 1e-6), t, np.inf)
    best = np.minimum(best, t)
    lo, hi = np.asarray(_BOX[0]), np.asarray(_BOX[1])
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo - o) / d
        t2 = (hi - o) / d
    tn = np.max(np.minimum(t1, t2), -1)
    tf = np.min(np.maximum(t1, t2), -1)
    t = np.where((tn <= tf) & (tn > 1e-6), tn, np.inf)
    best = np.minimum(best, t)
    z = np.where(np.isfinite(best) & (best >= zmin) & (best <= zmax), best, 0.0)
    return z


def render_depth_u16(pose, rows=480, cols=640, K=K_KINECT, seed=None, dropout=0.0, zmin=0.5, zmax=4.5):
    """uint16 millimetre depth.  seed != None adds Kinect-like noise sigma_z = 0.0012+0.0019(z-0.4)^2;
    dropout > 0 zeroes that fraction of pixels (rng seed+100)."""
    z = render_depth_m(pose, rows, cols, K, zmin, zmax)
    if seed is not None:
        rng = np.random.default_rng(seed)
        sig = 0.0012 + 0.0019 * (z - 0.4) ** 2
        z = np.where(z > 0, z + rng.standard_normal(z.shape) * sig, 0.0)
    mm = np.where(z > 0, np.round(z * 1000.0), 0.0)
    mm = np.clip(mm, 0, 65535).astype(np.uint16)
    if dropout > 0:
        rng = np.random.default_rng((0 if seed is None else seed) + 100)
        mm = np.where(rng.random(mm.shape) < dropout, 0, mm).astype(np.uint16)
    return mm


def u16_to_m(raw, scale=0.001):
    """DepthImage_convert_16UC1_to_32FC1 (pwn_static.cpp:54-68): scale*v, 0 stays 0 (float32)."""
    raw = np.asarray(raw)
    return np.where(raw > 0, np.float32(scale) * raw.astype(np.float32), np.float32(0)).astype(np.float32)


def perturbed_pose(rng, base, max_t=0.05, max_deg=3.0):
    """base o U(+-max_t, +-max_deg) -- config 4's loop-closure candidates."""
    ax = rng.standard_normal(3)
    return base @ make_pose(rng.uniform(-max_t, max_t, 3), ax, rng.uniform(-max_deg, max_deg))


def trajectory(n, seed=0, step_t=0.01, step_deg=0.5):
    """n smooth camera poses (per-frame delta <= 2 cm / 1 deg) -- config 3's sequence."""
    rng = np.random.default_rng(seed)
    poses = [np.eye(4)]
    vel_t = np.zeros(3)
    for i in range(1, n):
        vel_t = 0.9 * vel_t + 0.1 * rng.uniform(-step_t, step_t, 3)
        ang = step_deg * np.sin(i * 0.05)
        d = make_pose(np.clip(vel_t, -0.02, 0.02), (0.1, 1.0, 0.05), float(np.clip(ang, -1.0, 1.0)))
        nxt = poses[-1] @ d
        # keep the camera inside the room and looking roughly forward
        nxt[:3, 3] = np.clip(nxt[:3, 3], [-0.8, -0.4, -0.3], [0.8, 0.3, 0.5])
        poses.append(nxt)
    return poses


def make_rig(n_cameras=4, width=1280, height=960, K=None, zmin=0.5, zmax=4.5):
    """BASELINE config 5 rig: pinholes yawed 0/90/180/270 deg about the vertical axis, small lever arms
    (in the spirit of pwn_test/multiprojectortest.cpp:60-67).  width/height are the arguments the
    reference passes to MultiPointProjector::addPointProjector."""
    if K is None:
        K = scaled_K(K_KINECT, width / 640.0)
    cams = []
    for i in range(n_cameras):
        yaw = 360.0 * i / n_cameras
        off = make_pose((0.0, 0.0, 0.0), (0, 1, 0), yaw) @ make_pose((0.0, 0.0, 0.05))
        cams.append(dict(K=np.asarray(K, np.float32), width=width, height=height, minD=zmin, maxD=zmax,
                         offset=off.astype(np.float32)))
    return cams


def render_rig_depth_u16(rig_pose, cams, seed=None):
    """Composite raw depth image of a MultiPointProjector rig in the layout the Aligner projects into:
    rows = max width (pixel u), cols = sum of heights (pixel v + column offset): block i is the transpose
    of camera i's (height x width) image."""
    rows = max(c["width"] for c in cams)
    cols = sum(c["height"] for c in cams)
    out = np.zeros((rows, cols), np.uint16)
    off = 0
    for i, c in enumerate(cams):
        img = render_depth_u16(rig_pose @ c["offset"].astype(np.float64), c["height"], c["width"], c["K"],
                               seed=None if seed is None else seed + 17 * i, zmin=c["minD"], zmax=c["maxD"])
        out[:c["width"], off:off + c["height"]] = img.T
        off += c["height"]
    return out
