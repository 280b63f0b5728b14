"""CPU: the restatement of the reference's tracking pre/post-processing (oracle/tracking_ref.py) pinned against the
REFERENCE'S OWN functions (ptt/datasets/kitti/kitti_tracking_utils.py: crop_center_pc, get_model, regularize_pc,
get_box_by_offset), imported from the reference tree with a minimal pyquaternion stand-in, and the MT19937 / numpy
`randint` / `uniform` emulation against numpy itself."""
import copy
import sys

import numpy as np
import pytest

from oracle import refload, tracking_ref as tr
from ptt_b200 import synth_tracks


def test_mt_stream_and_randint_emulation_match_numpy():
    raw = np.random.RandomState(1).randint(0, 2 ** 32, size=4096, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(raw, tr.stream()[:4096])
    for n in list(range(3, 70)) + [127, 128, 129, 511, 512, 513, 1000, 1023, 1025, 2049, 5000, 16385, 70000]:
        for size in (1024, 512):
            np.random.seed(1)
            want = np.random.randint(low=0, high=n, size=size, dtype=np.int64)
            got, used = tr.randint_seed1(n, size)
            assert np.array_equal(got, want), (n, size)
            u, _ = tr.uniform_pm1(used)
            assert u == np.random.uniform(-1, 1), (n, size)          # the stream position is right, too


@pytest.fixture(scope="module")
def ref_utils():
    if not refload.available():
        pytest.skip("no reference tree")
    refload.load()                                   # sys.path: reference + shims (incl. pyquaternion)
    # the file itself, by path: importing it as ptt.datasets.kitti.* would pull in the dataset classes (skimage, pandas, ...)
    import importlib.util
    import os
    path = os.path.join(refload.REFERENCE_ROOT, "ptt", "datasets", "kitti", "kitti_tracking_utils.py")
    spec = importlib.util.spec_from_file_location("ref_kitti_tracking_utils", path)
    ku = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ku)
    return ku


def _ref_box(ku, box):
    from pyquaternion import Quaternion
    return ku.Box(box.center.tolist(), box.wlh.tolist(), Quaternion(matrix=box.R))


@pytest.mark.reference
def test_crop_regularize_and_box_update_match_the_reference_functions(ref_utils):
    ku = ref_utils
    tracks = synth_tracks.make_tracklets(4, 6, seed=3, points_per_frame=3000)
    for t, (clouds, boxes) in enumerate(tracks):
        for i in range(1, len(clouds)):
            ref_box = boxes[i - 1] if t % 2 else synth_tracks.jitter(boxes[i - 1], seed=10 * t + i)
            rb = _ref_box(ku, ref_box)
            # ---- search area: crop_center_pc with a gt box + regularize_pc (eval_tracking_utils.py:154-185)
            pc, _, _ = ku.crop_center_pc(ku.PointCloud(clouds[i].copy()), rb, _ref_box(ku, boxes[i]), offset=0.0, scale=1.25)
            want = ku.regularize_pc(pc, 1024, istrain=False)
            got, pos = tr.search_cloud(clouds[i], ref_box, 0.0, 1.25, 1024, 0)
            assert np.array_equal(got, want.astype(np.float32)), (t, i)
            # ---- template: get_model over (first, previous) + regularize_pc (:187-229)
            model = ku.get_model([ku.PointCloud(clouds[0].copy()), ku.PointCloud(clouds[i - 1].copy())],
                                 [_ref_box(ku, boxes[0]), rb], offset=0.0, scale=1.25)
            want_t = ku.regularize_pc(model, 512, istrain=False)
            got_t, pos = tr.template_cloud([(clouds[0], boxes[0]), (clouds[i - 1], ref_box)], 0.0, 1.25, 512, pos)
            assert np.array_equal(got_t, want_t.astype(np.float32)), (t, i)
            # ---- post_process: get_box_by_offset incl. its random clamps (np.random continues after regularize_pc)
            rs = np.random.RandomState(100 * t + i)
            for est in (np.float32([0.1, -0.2, 0.05, 3.0]), np.float32([5.0, 0.3, 0.1, -7.5]), np.float32([0.2, 9.0, -0.1, 40.0]),
                        rs.normal(0, 2, 4).astype(np.float32)):
                for use_z in (True, False):
                    state = np.random.get_state()
                    want_b = ku.get_box_by_offset(copy.deepcopy(rb), est.copy(), use_z)
                    np.random.set_state(state)
                    got_b, _ = tr.box_by_offset(ref_box, est, use_z, pos)
                    np.testing.assert_allclose(got_b.center, want_b.center, rtol=0, atol=1e-12)
                    np.testing.assert_allclose(got_b.R, want_b.rotation_matrix, rtol=0, atol=1e-12)


def test_degenerate_crops():
    clouds, boxes = synth_tracks.make_tracklets(1, 2, seed=5, points_per_frame=2000)[0]
    far = boxes[0].copy()
    far.center = far.center + 100.0                      # nothing inside: <= 2 points -> zeros (:359-360)
    got, pos = tr.search_cloud(clouds[1], far, 0.0, 1.25, 1024, 7)
    assert got.shape == (1024, 3) and not got.any() and pos == 7
    few = clouds[1][:, :1024]
    same, pos = tr.regularize_pc(few, 1024, 7)            # n == size: no resampling, no reseed
    assert np.array_equal(same, few.T) and pos == 7
