"""Minimal stand-in for `pyquaternion.Quaternion` (not installed in this image).  TEST INFRASTRUCTURE: it exists so that
the reference's own kitti_tracking_utils.py (crop_center_pc, get_model, regularize_pc, get_box_by_offset) can be
imported and run to PIN oracle/tracking_ref.py.  Only what those functions use: construction from axis / angle
(radians) or a rotation matrix, `*`, `.inverse`, `.rotation_matrix`, `.elements`.  Hamilton convention, unit
quaternions, float64 -- the published pyquaternion semantics."""
import numpy as np


class Quaternion:
    def __init__(self, *args, **kw):
        if "matrix" in kw:
            self.q = self._from_matrix(np.asarray(kw["matrix"], np.float64)[:3, :3])
        elif "axis" in kw:
            angle = kw.get("radians", kw.get("angle"))
            if angle is None and "degrees" in kw:
                angle = np.deg2rad(kw["degrees"])
            axis = np.asarray(kw["axis"], np.float64)
            axis = axis / np.linalg.norm(axis)
            half = float(angle) / 2.0
            self.q = np.concatenate([[np.cos(half)], np.sin(half) * axis])
        elif len(args) == 1:
            a = args[0]
            self.q = a.q.copy() if isinstance(a, Quaternion) else np.asarray(a, np.float64).copy()
        elif len(args) == 4:
            self.q = np.asarray(args, np.float64)
        else:
            self.q = np.array([1.0, 0.0, 0.0, 0.0])

    @staticmethod
    def _from_matrix(m):
        t = np.trace(m)
        if t > 0:
            s = np.sqrt(t + 1.0) * 2
            q = [0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s]
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q = [(m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s]
        elif m[1, 1] > m[2, 2]:
            s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q = [(m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s]
        else:
            s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q = [(m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s]
        return np.asarray(q, np.float64)

    @property
    def elements(self):
        return self.q

    @property
    def inverse(self):
        w, x, y, z = self.q
        return Quaternion(np.array([w, -x, -y, -z]) / np.dot(self.q, self.q))

    def __mul__(self, o):
        a1, b1, c1, d1 = self.q
        a2, b2, c2, d2 = o.q
        return Quaternion(np.array([a1 * a2 - b1 * b2 - c1 * c2 - d1 * d2, a1 * b2 + b1 * a2 + c1 * d2 - d1 * c2,
                                    a1 * c2 - b1 * d2 + c1 * a2 + d1 * b2, a1 * d2 + b1 * c2 - c1 * b2 + d1 * a2]))

    @property
    def rotation_matrix(self):
        w, x, y, z = self.q / np.linalg.norm(self.q)
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
