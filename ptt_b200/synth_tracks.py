"""Synthetic tracklets for the batched tracking loop: a car-sized box moving over a ground plane
with clutter, seen as a LiDAR-like float32 cloud per frame in WORLD coordinates -- the inputs of the reference's
per-tracklet evaluation loop (tools/eval_utils/eval_tracking_utils.py:77-120: PCs, BBs).  Numpy only; deterministic."""
import numpy as np

CAR_WLH = (1.6, 3.9, 1.56)


class Box:
    """center (3,), R (3,3) rotation matrix, wlh (3,) = (width, length, height): float64 (the reference's Box with its
    quaternion written as a matrix, kitti_tracking_utils.py:67-82)."""

    def __init__(self, center, R, wlh):
        self.center = np.asarray(center, np.float64).copy()
        self.R = np.asarray(R, np.float64).copy()
        self.wlh = np.asarray(wlh, np.float64).copy()

    def copy(self):
        return Box(self.center, self.R, self.wlh)

    def as_row(self):
        """15 doubles: center | R row-major | wlh (the layout of the device-side box state)."""
        return np.concatenate([self.center, self.R.reshape(-1), self.wlh])

    @staticmethod
    def from_row(row):
        row = np.asarray(row, np.float64)
        return Box(row[0:3], row[3:12].reshape(3, 3), row[12:15])


def rot_z(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def jitter(box, seed, sigma_xy=0.15, sigma_deg=3.0):
    """A tracker-like perturbation of a box (what a previous result looks like next to the ground truth)."""
    rs = np.random.RandomState(seed)
    b = box.copy()
    b.center = b.center + np.array([rs.normal(0, sigma_xy), rs.normal(0, sigma_xy), rs.normal(0, 0.03)])
    b.R = b.R @ rot_z(np.deg2rad(rs.normal(0, sigma_deg)))
    return b


def _surface(rs, n, wlh):
    """n points on the surface of an axis-aligned box (length along x, width along y), object frame."""
    w, l, h = wlh
    face = rs.randint(0, 5, size=n)              # +x, -x, +y, -y, top
    u, v = rs.uniform(-0.5, 0.5, size=n), rs.uniform(-0.5, 0.5, size=n)
    p = np.zeros((n, 3))
    for f, (ax, sign) in enumerate(((0, 1), (0, -1), (1, 1), (1, -1), (2, 1))):
        m = face == f
        dims = [l, w, h]
        other = [a for a in range(3) if a != ax]
        p[m, ax] = sign * dims[ax] / 2
        p[m, other[0]] = u[m] * dims[other[0]]
        p[m, other[1]] = v[m] * dims[other[1]]
    return p


def make_tracklets(n_tracklets, n_frames, seed=0, points_per_frame=4000, wlh=CAR_WLH):
    """-> [(clouds, boxes)] with clouds = [float32 (3, n_i)] (n_i varies per frame) and boxes = [Box] (ground truth)."""
    out = []
    for t in range(n_tracklets):
        rs = np.random.RandomState(seed * 7919 + t * 104729 + 17)
        center = np.array([rs.uniform(6, 30), rs.uniform(-6, 6), -1.7 + wlh[2] / 2])
        yaw = rs.uniform(-np.pi, np.pi)
        speed, yaw_rate = rs.uniform(0.2, 1.0), rs.uniform(-0.04, 0.04)
        clutter = [(center[:2] + rs.uniform(-9, 9, size=2), rs.uniform(0.5, 2.5, size=3)) for _ in range(6)]
        clouds, boxes = [], []
        for i in range(n_frames):
            R = rot_z(yaw)
            boxes.append(Box(center, R, wlh))
            n = int(points_per_frame * rs.uniform(0.8, 1.2))
            n_obj = int(n * rs.uniform(0.03, 0.12))
            obj = _surface(rs, n_obj, wlh) @ R.T + center
            n_cl = n // 6
            cl = []
            for c_xy, dims in clutter:
                q = _surface(rs, n_cl // len(clutter), (dims[1], dims[0], dims[2]))
                q[:, :2] += c_xy
                q[:, 2] += -1.7 + dims[2] / 2
                cl.append(q)
            n_g = n - n_obj - sum(len(q) for q in cl)
            g = np.stack([center[0] + rs.uniform(-14, 14, size=n_g), center[1] + rs.uniform(-14, 14, size=n_g),
                          -1.7 + rs.normal(0, 0.02, size=n_g)], 1)
            pts = np.concatenate([obj, g] + cl) + rs.normal(0, 0.01, size=(n, 3))
            pts = pts[rs.permutation(n)]
            clouds.append(np.ascontiguousarray(pts.T.astype(np.float32)))
            center = center + R @ np.array([speed, 0.0, 0.0])
            yaw += yaw_rate
        out.append((clouds, boxes))
    return out


def pad_frames(tracks, frame, cap):
    """Frame `frame` of every tracklet as one padded batch: points (T, cap, 3) float32 (zeros beyond n) + counts (T,) int32."""
    T = len(tracks)
    pts = np.zeros((T, cap, 3), np.float32)
    cnt = np.zeros((T,), np.int32)
    for t, (clouds, _) in enumerate(tracks):
        c = clouds[frame]
        n = c.shape[1]
        if n > cap:
            raise ValueError("frame has %d points, capacity %d" % (n, cap))
        pts[t, :n] = c.T
        cnt[t] = n
    return pts, cnt
