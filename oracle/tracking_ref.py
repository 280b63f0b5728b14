"""CPU restatement of the reference's per-frame tracking pre/post-processing.  TEST INFRASTRUCTURE.

What it restates (numpy, no pyquaternion -- boxes are (center, rotation matrix, wlh) in float64):

    crop_pc / crop_center_pc          ptt/datasets/kitti/kitti_tracking_utils.py:277-340
    get_model ("firstandprevious")    :219-237, tools/eval_utils/eval_tracking_utils.py:187-229
    regularize_pc (istrain=False)     :342-367   -- np.random.seed(1) + np.random.randint(0, n, size) every call
    get_box_by_offset                 :192-216   -- incl. its np.random.uniform(-1, 1) clamps
    the frame loop                    tools/eval_utils/eval_tracking_utils.py:77-120,140-274 (REF_BOX previous_result)

Arithmetic decisions recorded here (they are what the CUDA kernels reproduce bit for bit):
  * clouds are float32 (3, n) as KITTI loads them (kitti_dataset_tracking.py:304); PointCloud.translate / rotate assign
    float64 results back INTO the float32 array, i.e. every step rounds to float32 once (NumPy >= 2 promotion: the
    float64 scalar / matrix wins, the in-place assignment rounds);
  * comparisons against the crop bounds are exact float64 comparisons of the float32 coordinates;
  * a 3-term dot product r0*x + r1*y + r2*z is evaluated left to right in float64 without FMA (the reference calls BLAS,
    whose order is unspecified; the result is rounded to float32 afterwards, so the two agree except in double-rounding
    corner cases);
  * the random numbers are the MT19937 stream of seed 1 consumed exactly as numpy's legacy `randint` (masked rejection
    of 32-bit outputs) and `uniform` (two outputs -> 53-bit double) consume it; `mt_pos` is the position in that stream.

Pinned against the reference's own functions (with a minimal pyquaternion shim) by tests/test_tracking_cpu.py.
"""
import numpy as np

MT_STREAM_LEN = 1 << 15


def mt19937_stream(seed=1, n=MT_STREAM_LEN):
    """First n 32-bit outputs of MT19937 seeded like np.random.seed(seed)."""
    bg = np.random.MT19937()
    bg._legacy_seeding(seed)
    return bg.random_raw(n).astype(np.uint32)


_STREAM = None


def stream():
    global _STREAM
    if _STREAM is None:
        _STREAM = mt19937_stream(1)
    return _STREAM


def randint_seed1(n, size):
    """np.random.seed(1); np.random.randint(0, n, size) -> (indices int64, raw outputs consumed)."""
    raw = stream()
    rng = n - 1
    mask = rng
    for s in (1, 2, 4, 8, 16):
        mask |= mask >> s
    v = raw & np.uint32(mask)
    pos = np.nonzero(v <= rng)[0][:size]
    assert len(pos) == size, "MT stream table too short"
    return v[pos].astype(np.int64), int(pos[-1]) + 1


def uniform_pm1(mt_pos):
    """np.random.uniform(-1, 1) at stream position mt_pos -> (value, new position)."""
    raw = stream()
    a, b = int(raw[mt_pos]) >> 5, int(raw[mt_pos + 1]) >> 6
    return -1.0 + 2.0 * ((a * 67108864.0 + b) / 9007199254740992.0), mt_pos + 2


class Box:
    """center (3,), R (3,3) rotation matrix, wlh (3,): all float64."""

    def __init__(self, center, R, wlh):
        self.center = np.asarray(center, np.float64).copy()
        self.R = np.asarray(R, np.float64).copy()
        self.wlh = np.asarray(wlh, np.float64).copy()

    def copy(self):
        return Box(self.center, self.R, self.wlh)

    def as_row(self):
        return np.concatenate([self.center, self.R.reshape(-1), self.wlh])

    @staticmethod
    def from_row(row):
        return Box(row[0:3], np.asarray(row[3:12]).reshape(3, 3), row[12:15])


def rot_z(theta):
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _dot3(r, x, y, z):
    return (r[0] * x + r[1] * y) + r[2] * z


def corner_bounds(center, R, wlh, scale, offset):
    """max / min over Box.corners() (:170-189) of the box with wlh * scale, +- offset."""
    w, l, h = wlh * scale
    xs = l / 2 * np.array([1, 1, 1, 1, -1, -1, -1, -1], np.float64)
    ys = w / 2 * np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float64)
    zs = h / 2 * np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float64)
    c = np.stack([_dot3(R[i], xs, ys, zs) + center[i] for i in range(3)])
    return c.max(1) + offset, c.min(1) - offset


def crop_mask(pts, maxi, mini):
    """crop_pc (:277-297): strict inequalities; pts (3, n) float32 compared as float64."""
    p = pts.astype(np.float64)
    m = np.ones(pts.shape[1], bool)
    for i in range(3):
        m &= (p[i] > mini[i]) & (p[i] < maxi[i])
    return m


def crop_center_pc(pts, box, offset, scale, second_offset):
    """crop_center_pc (:300-340) without labels.  pts (3, n) float32 -> (3, m) float32 in the box frame.
    second_offset = offset + 0.6 * wlh[1] with a gt_box (the search area, :323), offset without (the template, :333)."""
    maxi, mini = corner_bounds(box.center, box.R, box.wlh, 4 * scale, 2 * offset)
    p = pts[:, crop_mask(pts, maxi, mini)].astype(np.float32)
    trans = -box.center
    q = np.empty_like(p)
    for i in range(3):
        q[i] = (p[i].astype(np.float64) + trans[i]).astype(np.float32)          # PointCloud.translate
    rt = box.R.T
    x, y, z = (q[i].astype(np.float64) for i in range(3))
    r = np.stack([_dot3(rt[i], x, y, z) for i in range(3)]).astype(np.float32)    # PointCloud.rotate
    # the box itself is now at the origin with identity orientation
    maxi2, mini2 = corner_bounds(np.zeros(3), np.eye(3), box.wlh, scale, second_offset)
    return r[:, crop_mask(r, maxi2, mini2)]


def regularize_pc(pts, size, mt_pos):
    """regularize_pc(istrain=False) (:342-367): (3, n) float32 -> ((size, 3) float32, new mt_pos)."""
    n = pts.shape[1]
    if n > 2:
        if n != size:
            idx, mt_pos = randint_seed1(n, size)
            pts = pts[:, idx]
        return np.ascontiguousarray(pts.T.astype(np.float32)), mt_pos
    return np.zeros((size, 3), np.float32), mt_pos


def search_cloud(pts, ref_box, offset, scale, size, mt_pos):
    """prepare_search (eval_tracking_utils.py:154-185); the gt box shares its wlh with the tracked box."""
    crop = crop_center_pc(pts, ref_box, offset, scale, offset + ref_box.wlh[1] * 0.6)
    return regularize_pc(crop, size, mt_pos)


def template_cloud(sources, offset, scale, size, mt_pos):
    """prepare_template (:187-229): get_model over [(pts, box), ...] (first, previous), then regularize_pc."""
    parts = [crop_center_pc(p, b, offset, scale, offset) for p, b in sources]
    parts = [c for c in parts if c.shape[1] > 0]
    model = np.concatenate(parts, axis=1) if parts else np.zeros((3, 0), np.float32)
    return regularize_pc(model, size, mt_pos)


def box_by_offset(box, est, use_z, mt_pos):
    """get_box_by_offset (:192-216) on the best proposal est = (x, y, z, theta_degrees), a FLOAT32 row of pred_box_data
    (eval_tracking_utils.py:266-270).  NumPy >= 2 promotion: `offset[-1] * np.pi / 180` stays float32 (Python floats are
    weak), and a clamp's np.random.uniform(-1, 1) is stored back into the float32 row."""
    est = np.asarray(est, np.float32)
    o = [est[0], est[1], est[2]]
    if o[0] > box.wlh[0]:
        u, mt_pos = uniform_pm1(mt_pos)
        o[0] = np.float32(u)
    if o[1] > min(box.wlh[1], 2):
        u, mt_pos = uniform_pm1(mt_pos)
        o[1] = np.float32(u)
    oz = float(o[2]) if use_z else 0.0
    theta = float(np.float32(np.float32(est[3] * np.float32(np.pi)) / np.float32(180)))
    new = box.copy()
    new.R = box.R @ rot_z(theta)
    new.center = box.center + np.array([_dot3(box.R[i], float(o[0]), float(o[1]), oz) for i in range(3)])
    return new, mt_pos


def track(clouds, first_box, model, offset=0.0, scale=1.25, model_offset=0.0, model_scale=1.25, n_search=1024,
          n_template=512, use_z=True, mt_pos=0):
    """TrackingEvaluator.test_batch for one tracklet (:77-120): clouds = [(3, n_i) float32], first_box = BBs[0];
    model(search (1,Ns,3), template (1,Nt,3)) -> pred_box_data (64, 5).  Returns the result boxes (one per frame)."""
    results = [first_box.copy()]
    for i in range(1, len(clouds)):
        ref = results[-1]
        search, mt_pos = search_cloud(clouds[i], ref, offset, scale, n_search, mt_pos)
        template, mt_pos = template_cloud([(clouds[0], results[0]), (clouds[i - 1], results[i - 1])], model_offset,
                                          model_scale, n_template, mt_pos)
        est = model(search[None], template[None])
        best = est[np.argmax(est[:, 4])]
        box, mt_pos = box_by_offset(ref, best[:4], use_z, mt_pos)
        results.append(box)
    return results
