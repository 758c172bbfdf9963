"""Mirror of ``load_traj.py`` (reference :9-280): ``DataLoader`` with the reference's constructor
signature, attributes and ``next_step()`` result structure, plus the device-side padded scene
batching the north_star asks for (``device_table`` / ``scene_batch``: CSV -> (frame, ped)-sorted
table resident in HBM -> ``mmt_scene_batch_f32`` windows [S,N,F,2] + valid mask).

Differences from the reference, all deliberate: the data root is a parameter (the reference
hard-codes ``/home/siri0005/...``, load_traj.py:20); the frame dictionary is kept in memory instead
of being pickled into the data directory (:234-256); nothing else changes the values returned.
"""
from __future__ import annotations

import glob
import math
import os

import numpy as np

DATASET_DIRS = ['eth/hotel/', 'eth/univ/', 'ucy/zara/zara01/', 'ucy/zara/zara02/', 'ucy/univ/',
                'town_center.csv', 'annotation_tc.txt']          # load_traj.py:25-33


def default_data_root():
    """``data/`` of the working directory if it exists, else the tables shipped at the repository root."""
    if os.path.isdir("data"):
        return "data"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


class DataLoader():
    def __init__(self, args, datasets=(0, 1, 2, 3, 4, 5, 6), sel=None, start=0, processFrame=False, infer=False,
                 parent_dir=None, csv=None):
        parent_dir = parent_dir or getattr(args, "data_root", None) or default_data_root()
        self.data_dirs = [os.path.join(parent_dir, d) for d in DATASET_DIRS]
        self.used_data_dirs = [self.data_dirs[x] for x in datasets]
        self.infer = infer
        self.numDatasets = len(self.data_dirs)
        self.data_dir = parent_dir
        self.batch_size, self.seq_length = args.batch_size, args.seq_length
        self.pred_len, self.obs_len = args.pred_len, args.obs_len
        self.diff = self.obs_len
        self.current_dir = self.used_data_dirs[start]
        if csv is not None:                      # in-memory table (tests, synthetic data)
            self.dataset_pointer = sel if sel is not None else 0
            self.sel_file = "<memory>"
            self._load_array(np.asarray(csv, np.float64), val=infer)
        elif os.path.isdir(self.current_dir):
            # the reference takes the alphabetically first *.csv (load_traj.py:77-86); the tables shipped under data/
            # are gzip-compressed (numpy reads them transparently): order by the name without the .gz suffix
            files = sorted(glob.glob(self.current_dir + "*.csv") + glob.glob(self.current_dir + "*.csv.gz"),
                           key=lambda f: f[:-3] if f.endswith(".gz") else f)
            if sel is None:
                sel = 0 if len(files) == 1 else int(input('select which file you want for loading:'))
            self.dataset_pointer = sel
            self.sel_file = files[int(sel)]
            self.load_dataset(self.sel_file, val=infer)
        else:
            self.dataset_pointer = start
            self.sel_file = self.current_dir
            self.load_dataset(self.current_dir, val=infer)
        self.frame_preprocess(self.sel_file, seed=self.seed)
        self.num_batches = int((len(self.frameList) / self.seq_length) / self.batch_size)       # :104
        self.valid_num_batches = self.num_batches
        self.valid_frame_pointer = self.seed

    # ---- load_traj.py:114-150
    def load_dataset(self, data_file, val=False):
        self._load_array(np.genfromtxt(fname=data_file, delimiter=','), val)

    def _load_array(self, raw, val=False):
        self.raw_data = raw
        self.len = raw.shape[1]
        self.max = int(raw.shape[1] * 0.7)
        self.val_max = int(raw.shape[1] * 0.3)
        self.val_data = raw[:, self.max:self.max + self.val_max]
        self.tr_data = raw[:, 0:self.max]
        part = self.val_data if val else self.tr_data
        self.frameList = part[0, :]
        self.pedsPerFrameList = part[0:4, :]
        self.vislet = part[4:6, :]
        self.seed = self.frameList[0]
        self.frame_pointer = self.seed

    def load_trajectories(self, data_file=None):
        return self.trajectories

    # ---- load_traj.py:234-256: {frame: [{ped: [row2, row3]}, ...]} for frames seed, seed+diff, ...
    def frame_preprocess(self, data_file=None, seed=0):
        cols = self.pedsPerFrameList
        frames = cols[0].astype(np.int64)
        data = {int(f): {} for f in np.unique(frames)}
        order = np.argsort(frames, kind="stable")
        bounds = np.searchsorted(frames[order], np.arange(int(self.seed), int(frames.max()) + 1, self.diff))
        ends = np.searchsorted(frames[order], np.arange(int(self.seed), int(frames.max()) + 1, self.diff), side="right")
        for f, lo, hi in zip(range(int(self.seed), int(frames.max()) + 1, self.diff), bounds, ends):
            data[f] = [{cols[1, c]: [cols[2, c], cols[3, c]]} for c in order[lo:hi]]
        self.trajectories = data
        self.frame_pointer = self.seed      # the reference leaves the pointer past the end and resets on the
        return data                         # first empty batch (train.py:63-67); start reset instead

    # ---- load_traj.py:153-224
    def next_step(self, targets=None):
        targets = {} if targets is None else targets
        x_batch, window = {}, {}
        pc = 1
        max_idx = max(self.frameList)
        max_log = math.log(max_idx, self.diff)
        idx = self.frame_pointer
        for _ in range(self.batch_size + 1):
            room = max_idx - (idx + 1)
            if room <= 0:
                break
            if math.log(abs(room), self.diff) > max_log:
                self.tick_frame_pointer(valid=False)
                continue
            start = int(self.frame_pointer)
            for f in range(start, int(self.frame_pointer + self.batch_size * self.obs_len), self.diff):
                if f not in self.trajectories:
                    break
                window[f] = self.trajectories[f]
                idx = f
            keys = list(window)
            cursor = 0                                   # position of the reference's second iterator
            for f in keys:
                idx = f
                frame = self.trajectories[f]
                if len(frame):
                    x_batch[f] = frame
                    if pc % self.obs_len == 0:
                        if cursor >= len(keys):
                            break
                        tgt_frame = self.trajectories[keys[cursor]]
                        cursor += 1
                        for _rep in range(int(self.pred_len)):
                            for item in tgt_frame:
                                (pid, pos), = item.items()
                                targets.setdefault(int(pid), []).append(pos)
                pc += 1
                if cursor >= len(keys):
                    break
                cursor += 1
            self.frame_pointer += self.diff
        return x_batch, targets, self.frame_pointer

    def tick_frame_pointer(self, valid=False, incr=8):
        if not valid:
            self.frame_pointer += incr

    def reset_data_pointer(self, valid=False, dataset_pointer=0, frame_pointer=0):
        if not valid:
            self.frame_pointer = self.seed
        else:
            self.dataset_pointer = dataset_pointer
            self.frame_pointer = frame_pointer
            self.valid_frame_pointer = frame_pointer

    # ---------------------------------------------------------------------------------------------
    # device-side padded scene batching (new; replaces next_step + ConstructGraph on the batched path)
    def device_table(self, device="cuda", val=False):
        """(frame, ped)-sorted table of the current split on the device:
        frame_ids[nf] i32, row_start[nf+1] i32, ped[M] i32, xy[M,2] f32, vis[M,2] f32 | None."""
        import torch
        part = self.val_data if val else self.tr_data
        fr, ped = part[0].astype(np.int64), part[1].astype(np.int64)
        order = np.lexsort((ped, fr))
        fr, ped = fr[order], ped[order]
        xy = np.ascontiguousarray(part[2:4, order].T.astype(np.float32))
        vis = np.ascontiguousarray(part[4:6, order].T.astype(np.float32)) if part.shape[0] >= 6 else None
        frame_ids, first = np.unique(fr, return_index=True)
        row_start = np.concatenate([first, [len(fr)]]).astype(np.int32)
        t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
        return dict(frame_ids=t(frame_ids.astype(np.int32)), row_start=t(row_start), ped=t(ped.astype(np.int32)),
                    xy=t(xy), vis=t(vis), stride=int(np.min(np.diff(frame_ids))) if len(frame_ids) > 1 else 1)

    def scene_batch(self, table, N, F=None, hop=1):
        """All windows of F frames (default obs_len + pred_len) starting every ``hop`` frames ->
        (pos[S,N,F,2], vis[S,N,F,2] | None, valid[S,N], ped_of_slot[S,N]) on the device."""
        from . import ops
        F = F or (self.obs_len + self.pred_len)
        wins = table["frame_ids"][:max(0, table["frame_ids"].shape[0] - F + 1):hop].contiguous()
        return ops.scene_batch(table["frame_ids"], table["row_start"], table["ped"], table["xy"], table["vis"], wins,
                               N, F, table["stride"])
