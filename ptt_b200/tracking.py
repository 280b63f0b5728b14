"""Batched multi-tracklet tracking loop on one B200.

The reference tracks ONE object at a time: per frame it crops the raw cloud around the previous result box in numpy,
resamples it to a fixed size, copies it to the GPU, runs the model at batch 1, copies `pred_box_data` back, picks the best
proposal and moves the box on the host (tools/eval_utils/eval_tracking_utils.py:77-120, 140-274).  Frame t+1 needs frame
t's box, so a tracklet is strictly sequential -- but TRACKLETS are independent.  `BatchedTracker` advances T of them in
lockstep, one frame per step, with everything between the raw clouds and the new boxes on the device and inside ONE
CUDA graph:

    crop around the previous box (search area; template = first frame + previous frame)    ptt_track_crop
    seeded resampling to 1024 / 512 points (np.random.seed(1) + randint, bit-exact)        ptt_track_regularize
    the whole tracker forward, batch T                                                     HotPath.forward_full
    best proposal (argmax of the score column) -> get_box_by_offset                        ptt_track_update

The only per-frame host traffic is the H2D copy of the T raw clouds (padded to `max_points`), staged one frame ahead
on a copy stream; the boxes of all frames stay on the device until `results()` is called.

Configuration = the YAML's DATA_CONFIG / TEST keys (kitti_models/ptt.yaml:8-17,149-150): SEARCH/MODEL_BB_OFFSET and
_SCALE, SEARCH/TEMPLATE_INPUT_SIZE, USE_Z_AXIS, REF_BOX = previous_result, SHAPE_AGGREGATION = firstandprevious.
"""
import torch

from . import hotpath, ops

MT_LEN = 1 << 15          # outputs of the seed-1 MT19937 stream kept on the device (a resampling consumes <= ~2.2 * size)


class BatchedTracker:
    def __init__(self, state_dict, n_tracklets, max_points, max_frames, device="cuda", cfg=None, search_size=1024,
                 template_size=512, search_offset=0.0, search_scale=1.25, model_offset=0.0, model_scale=1.25, use_z=True):
        self.T, self.cap, self.F = int(n_tracklets), int(max_points), int(max_frames)
        self.device = torch.device(device)
        self.hp = hotpath.HotPath(state_dict, cfg=cfg, device=self.device)
        if not self.hp.full:
            raise RuntimeError("BatchedTracker needs the whole tracker's parameters (similarity module and heads)")
        self.p = dict(search_size=int(search_size), template_size=int(template_size), search_offset=float(search_offset),
                      search_scale=float(search_scale), model_offset=float(model_offset), model_scale=float(model_scale),
                      use_z=bool(use_z))
        d, T, cap = self.device, self.T, self.cap
        f32, i32, f64 = torch.float32, torch.int32, torch.float64
        self.mt = ops.mt19937_stream(MT_LEN, seed=1, device=d)
        z = lambda *shape, dtype=f32: torch.zeros(*shape, dtype=dtype, device=d)
        # static buffers of the per-frame graph
        self.cur_pts, self.cur_cnt = z(T, cap, 3), z(T, dtype=i32)
        self.prev_pts, self.prev_cnt = z(T, cap, 3), z(T, dtype=i32)
        self.first_crop, self.first_cnt = z(T, cap, 3), z(T, dtype=i32)       # frame 0 cropped by results[0]: constant
        self.tmp_s, self.cnt_s = z(T, cap, 3), z(T, dtype=i32)
        self.tmp_t, self.cnt_t = z(T, 2 * cap, 3), z(T, dtype=i32)
        self.search, self.template = z(T, self.p["search_size"], 3), z(T, self.p["template_size"], 3)
        self.state = z(T, 15, dtype=f64)
        self.mt_pos = z(T, dtype=i32)
        self.results_buf = z(self.F, T, 15, dtype=f64)
        self.frame_idx = z(1, dtype=i32)
        self.stage = [z(T, cap, 3), z(T, cap, 3)]                              # H2D staging, one frame ahead
        self.stage_cnt = [z(T, dtype=i32), z(T, dtype=i32)]
        self.stream = torch.cuda.Stream(d)
        self.copy_stream = torch.cuda.Stream(d)
        self._staged = [None, None]
        self._next_slot = 0
        self._graph = None
        self._out = None
        self.frames_done = 0

    # ------------------------------------------------------------------------------------------------
    def _frame(self):
        """One frame for all tracklets on the current stream (the body of the captured graph)."""
        p = self.p
        # prepare_search (eval_tracking_utils.py:154-185): REF_BOX = previous_result
        ops.track_crop([(self.cur_pts, self.cur_cnt, self.state)], p["search_offset"], p["search_scale"], True,
                       self.tmp_s, self.cnt_s)
        ops.track_regularize(self.tmp_s, self.cnt_s, p["search_size"], self.mt, self.mt_pos, self.search)
        # prepare_template (:187-229): SHAPE_AGGREGATION = firstandprevious -> get_model([PC_0, PC_{i-1}], [BB_0, BB_{i-1}])
        ops.track_crop([(self.first_crop, self.first_cnt, None), (self.prev_pts, self.prev_cnt, self.state)],
                       p["model_offset"], p["model_scale"], False, self.tmp_t, self.cnt_t)
        ops.track_regularize(self.tmp_t, self.cnt_t, p["template_size"], self.mt, self.mt_pos, self.template)
        out = self.hp.forward_full(self.search, self.template)                 # model_inference (:231-264)
        # post_process (:266-274): the best proposal was selected on the device (forward_full: best_box)
        ops.track_update(out["best_box"], self.state, p["use_z"], self.mt, self.mt_pos, self.results_buf, self.frame_idx)
        self.prev_pts.copy_(self.cur_pts)
        self.prev_cnt.copy_(self.cur_cnt)
        return out

    def _capture(self):
        keep = [t.clone() for t in (self.state, self.mt_pos, self.results_buf, self.frame_idx, self.prev_pts, self.prev_cnt)]
        self.hp._ws_scope = ("tracker", self.T)              # the graph owns its workspaces (HotPath._workspace)
        for _ in range(2):                                    # warm-up: kernel attributes, allocator pools
            self._frame()
        torch.cuda.current_stream().synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=torch.cuda.current_stream()):
            self._out = self._frame()
        self.hp._ws_scope = None
        for dst, src in zip((self.state, self.mt_pos, self.results_buf, self.frame_idx, self.prev_pts, self.prev_cnt), keep):
            dst.copy_(src)
        self._graph = graph

    # ------------------------------------------------------------------------------------------------
    def reset(self, first_points, first_counts, first_boxes):
        """Frame 0 of every tracklet: points (T,max_points,3) float32, counts (T,) int32, boxes (T,15) float64 (the
        ground-truth boxes BBs[0]; they ARE results[0], eval_tracking_utils.py:95-99).  Host or device tensors."""
        d = self.device
        with torch.cuda.stream(self.stream):
            self.prev_pts.copy_(first_points.to(d, non_blocking=True))
            self.prev_cnt.copy_(first_counts.to(d, non_blocking=True))
            self.state.copy_(first_boxes.to(d, non_blocking=True))
            self.mt_pos.zero_()
            self.results_buf.zero_()
            self.results_buf[0].copy_(self.state)
            self.frame_idx.fill_(1)
            # the first-frame part of the template never changes: crop it once (get_model's first source)
            ops.track_crop([(self.prev_pts, self.prev_cnt, self.state)], self.p["model_offset"], self.p["model_scale"], False,
                           self.first_crop, self.first_cnt)
        self.frames_done = 1
        self._staged = [None, None]

    def stage_frame(self, points_host, counts_host):
        """Start the H2D copy of the NEXT frame's raw clouds (pinned host tensors) on the copy stream."""
        slot = self._next_slot
        self._next_slot ^= 1
        if self._staged[slot] is not None:
            raise RuntimeError("both staging slots hold frames that have not been consumed by step()")
        with torch.cuda.stream(self.copy_stream):
            self.stage[slot].copy_(points_host, non_blocking=True)
            self.stage_cnt[slot].copy_(counts_host, non_blocking=True)
            ev = self.copy_stream.record_event()
        self._staged[slot] = ev
        return slot

    def step(self, points=None, counts=None, slot=None):
        """Advance every tracklet by one frame.  Either pass the frame (host or device tensors) or the `slot` a previous
        stage_frame() returned.  Asynchronous: returns as soon as the work is enqueued."""
        if self.frames_done >= self.F:
            raise RuntimeError("max_frames reached")
        with torch.cuda.stream(self.stream):
            if slot is not None:
                self.stream.wait_event(self._staged[slot])
                self.cur_pts.copy_(self.stage[slot])
                self.cur_cnt.copy_(self.stage_cnt[slot])
                done = self.stream.record_event()
                self.copy_stream.wait_event(done)            # the slot may be overwritten once this copy has run
                self._staged[slot] = None
            else:
                self.cur_pts.copy_(points.to(self.device, non_blocking=True))
                self.cur_cnt.copy_(counts.to(self.device, non_blocking=True))
            if self._graph is None:
                self._capture()
            self._graph.replay()
        self.frames_done += 1

    def results(self):
        """(frames_done, T, 15) float64 boxes so far (device tensor; waits for the enqueued frames)."""
        self.stream.synchronize()
        return self.results_buf[: self.frames_done]

    def last_inputs(self):
        """The regularised search / template clouds and the model outputs of the last frame (device; for tests)."""
        self.stream.synchronize()
        return self.search, self.template, self._out


def run_tracklets(tracker, frames):
    """frames: iterable of (points (T,cap,3) pinned float32, counts (T,) pinned int32) for frames 1, 2, ...; the H2D copy
    of frame i+1 overlaps the compute of frame i.  Returns tracker.results()."""
    it = iter(frames)
    nxt = next(it, None)
    slot = tracker.stage_frame(*nxt) if nxt is not None else None
    while nxt is not None:
        cur_slot = slot
        nxt = next(it, None)
        tracker.step(slot=cur_slot)
        if nxt is not None:
            slot = tracker.stage_frame(*nxt)
    return tracker.results()
