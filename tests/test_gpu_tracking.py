"""GPU: the batched tracking loop (ptt_b200.tracking.BatchedTracker) against the CPU restatement of
the reference's per-frame pre / post-processing (oracle/tracking_ref.py, itself pinned against the reference's own
functions by tests/test_tracking_cpu.py).

Every frame is checked with the GPU's own box state as the starting point (a tracker is a feedback loop: a 1e-6
difference in a predicted box may move a point across a crop boundary a few frames later, so whole-loop equality is not
a meaningful bar; per-frame equality from identical state is, and it is exact): the regularised search and template
clouds bit for bit, the seed-1 random stream position exactly, the updated boxes to 1e-12."""
import numpy as np
import pytest
import torch

from conftest import t
from oracle import tracking_ref as tr
from ptt_b200 import ops, synth, synth_tracks, tracking

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _boxes_tensor(boxes):
    return torch.from_numpy(np.stack([b.as_row() for b in boxes]))


def test_crop_and_regularize_kernels_vs_oracle():
    tracks = synth_tracks.make_tracklets(6, 3, seed=11, points_per_frame=5000)
    cap = 6200
    mt = ops.mt19937_stream(tracking.MT_LEN, device=DEV)
    assert np.array_equal(mt.cpu().numpy().view(np.uint32), tr.stream()[: tracking.MT_LEN])
    for search, (off, scale, size) in ((True, (0.0, 1.25, 1024)), (False, (0.0, 1.25, 512)), (True, (0.3, 1.0, 256))):
        boxes = [synth_tracks.jitter(b[1], seed=k) for k, (_, b) in enumerate(tracks)]
        boxes[4].center += 50.0                                    # nothing inside the crop -> zeros
        pts, cnt = synth_tracks.pad_frames(tracks, 1, cap)
        out = torch.empty(len(tracks), cap, 3, device=DEV)
        out_cnt = torch.empty(len(tracks), dtype=torch.int32, device=DEV)
        ops.track_crop([(t(pts).to(DEV), t(cnt).to(DEV), _boxes_tensor(boxes).to(DEV))], off, scale, search, out, out_cnt)
        reg = torch.empty(len(tracks), size, 3, device=DEV)
        mt_pos = torch.full((len(tracks),), 5, dtype=torch.int32, device=DEV)
        ops.track_regularize(out, out_cnt, size, mt, mt_pos, reg)
        for k, (clouds, _) in enumerate(tracks):
            second = off + boxes[k].wlh[1] * 0.6 if search else off
            want = tr.crop_center_pc(clouds[1], boxes[k], off, scale, second)
            n = int(out_cnt[k])
            assert n == want.shape[1], (search, k)
            assert np.array_equal(out[k, :n].cpu().numpy(), want.T), (search, k)
            want_reg, pos = tr.regularize_pc(want, size, 5)
            assert np.array_equal(reg[k].cpu().numpy(), want_reg), (search, k)
            assert int(mt_pos[k]) == pos
        assert int(out_cnt[4]) == 0 and not reg[4].any()


def test_box_update_kernel_vs_oracle():
    rs = np.random.RandomState(0)
    T = 64
    boxes = [synth_tracks.Box(rs.normal(0, 10, 3), synth_tracks.rot_z(rs.uniform(-3, 3)), (1.6, 3.9, 1.56)) for _ in range(T)]
    est = rs.normal(0, 1.5, size=(T, 5)).astype(np.float32)
    est[:8, 0] = 5.0                                               # > wlh[0]: first random clamp
    est[4:12, 1] = 3.0                                             # > min(wlh[1], 2): second random clamp
    mt = ops.mt19937_stream(tracking.MT_LEN, device=DEV)
    for use_z in (True, False):
        state = _boxes_tensor(boxes).to(DEV)
        pos0 = rs.randint(0, 3000, size=T).astype(np.int32)
        mt_pos = t(pos0).to(DEV)
        results = torch.zeros(4, T, 15, dtype=torch.float64, device=DEV)
        frame_idx = torch.full((1,), 2, dtype=torch.int32, device=DEV)
        ops.track_update(t(est).to(DEV), state, use_z, mt, mt_pos, results, frame_idx)
        assert int(frame_idx) == 3 and torch.equal(results[2], state) and not results[3].any()
        for k in range(T):
            want, pos = tr.box_by_offset(boxes[k], est[k, :4], use_z, int(pos0[k]))
            np.testing.assert_allclose(state[k].cpu().numpy(), want.as_row(), rtol=0, atol=1e-12)
            assert int(mt_pos[k]) == pos


def test_batched_tracker_frame_by_frame_vs_oracle():
    T, F, cap = 5, 6, 4200
    tracks = synth_tracks.make_tracklets(T, F, seed=21, points_per_frame=3400)
    sd = synth.full_model_state_dict(0)
    trk = tracking.BatchedTracker(sd, T, cap, F, device=DEV)
    first = [b[0] for _, b in tracks]
    pts0, cnt0 = synth_tracks.pad_frames(tracks, 0, cap)
    trk.reset(t(pts0), t(cnt0), _boxes_tensor(first))
    for i in range(1, F):
        trk.stream.synchronize()
        before = trk.state.cpu().numpy().copy()
        pos_before = trk.mt_pos.cpu().numpy().copy()
        pts, cnt = synth_tracks.pad_frames(tracks, i, cap)
        trk.step(t(pts).pin_memory(), t(cnt).pin_memory())
        search, template, out = trk.last_inputs()
        after = trk.state.cpu().numpy()
        est = out["pred_box_data"].cpu().numpy()
        assert np.array_equal(out["best_box"].cpu().numpy(), est[np.arange(T), est[:, :, 4].argmax(1)])
        for k, (clouds, boxes) in enumerate(tracks):
            ref = synth_tracks.Box.from_row(before[k])
            s, pos = tr.search_cloud(clouds[i], ref, 0.0, 1.25, 1024, int(pos_before[k]))
            assert np.array_equal(search[k].cpu().numpy(), s), (i, k)
            m, pos = tr.template_cloud([(clouds[0], boxes[0]), (clouds[i - 1], ref)], 0.0, 1.25, 512, pos)
            assert np.array_equal(template[k].cpu().numpy(), m), (i, k)
            want, pos = tr.box_by_offset(ref, out["best_box"][k, :4].cpu().numpy(), True, pos)
            np.testing.assert_allclose(after[k], want.as_row(), rtol=0, atol=1e-12)
            assert int(trk.mt_pos[k]) == pos
    res = trk.results()
    assert res.shape == (F, T, 15) and torch.equal(res[0].cpu(), _boxes_tensor(first)) and torch.equal(res[-1], trk.state)


def test_run_tracklets_pipeline_equals_stepping():
    T, F, cap = 3, 5, 3000
    tracks = synth_tracks.make_tracklets(T, F, seed=31, points_per_frame=2300)
    sd = synth.full_model_state_dict(1)
    first = _boxes_tensor([b[0] for _, b in tracks])
    frames = [tuple(t(x).pin_memory() for x in synth_tracks.pad_frames(tracks, i, cap)) for i in range(F)]
    a = tracking.BatchedTracker(sd, T, cap, F, device=DEV)
    a.reset(frames[0][0], frames[0][1], first)
    for pts, cnt in frames[1:]:
        a.step(pts, cnt)
    b = tracking.BatchedTracker(sd, T, cap, F, device=DEV)
    b.reset(frames[0][0], frames[0][1], first)
    got = tracking.run_tracklets(b, frames[1:])
    assert torch.equal(got, a.results())
    assert (got[1:, :, :3] != got[:-1, :, :3]).any()                # the boxes do move
