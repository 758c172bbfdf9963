"""CPU oracle for device-side padded scene batching (TEST INFRASTRUCTURE ONLY).

Restates, as a clean window extraction, what ``load_traj.DataLoader.next_step``
(load_traj.py:153-224) and ``online_graph.ConstructGraph`` / ``Graph.setNodes``
(networkx_graph.py:30-73,114-129) assemble per batch: per-pedestrian position rows over a
window of frames.  The reference's accumulation defects (zero row on first
sighting, ``targets`` mutable default, node-axis slicing) are NOT reproduced; the rule is:
a pedestrian gets a slot -- in ascending id order of the window's first frame -- iff present in
all F frames of the window, up to N slots.
"""
from __future__ import annotations

import numpy as np


def table_from_csv(csv):
    """csv[4 or 6, M] (row0 frame, row1 ped, row2/3 position, row4/5 vislet; data/pixel_pos_format.md)
    -> table sorted by (frame, ped): frame_ids[nf], row_start[nf+1], ped[M], xy[M,2], vis[M,2]|None."""
    fr = csv[0].astype(np.int64)
    ped = csv[1].astype(np.int64)
    order = np.lexsort((ped, fr))
    fr, ped = fr[order], ped[order]
    xy = csv[2:4, order].T.astype(np.float32)
    vis = csv[4:6, order].T.astype(np.float32) if csv.shape[0] >= 6 else None
    frame_ids, start = np.unique(fr, return_index=True)
    row_start = np.concatenate([start, [len(fr)]]).astype(np.int32)
    return frame_ids.astype(np.int32), row_start, ped.astype(np.int32), np.ascontiguousarray(xy), \
        None if vis is None else np.ascontiguousarray(vis)


def scene_batch(frame_ids, row_start, ped, xy, vis, win_start, N, F, fstride):
    S = len(win_start)
    pos = np.zeros((S, N, F, 2), np.float32)
    vo = np.zeros((S, N, F, 2), np.float32) if vis is not None else None
    valid = np.zeros((S, N), np.uint8)
    slot_ped = np.full((S, N), -1, np.int32)
    index = {int(f): i for i, f in enumerate(frame_ids)}
    for s, w0 in enumerate(win_start):
        frames = [int(w0) + k * fstride for k in range(F)]
        if any(f not in index for f in frames):
            continue
        rows = []
        for f in frames:
            i = index[f]
            rows.append({int(ped[r]): r for r in range(row_start[i], row_start[i + 1])})
        n = 0
        i0 = index[frames[0]]
        for r0 in range(row_start[i0], row_start[i0 + 1]):
            p = int(ped[r0])
            if all(p in rk for rk in rows):
                if n < N:
                    valid[s, n] = 1
                    slot_ped[s, n] = p
                    for k, rk in enumerate(rows):
                        pos[s, n, k] = xy[rk[p]]
                        if vo is not None:
                            vo[s, n, k] = vis[rk[p]]
                n += 1
    return pos, vo, valid, slot_ped
