"""Observation sharding for multi-GPU bundle adjustment.

Observations arrive sorted by 3D line (``seq_utils.py:188-195``), so the observation
array is cut into ``world`` contiguous ranges at line boundaries, balanced by
observation count.  Every 3D line -- its observations, its 6x6 block, its update --
then lives on exactly one rank; camera parameters are replicated and only per-camera
partial sums cross NVLink.  Pure host-side integer logic (numpy), tested on CPU.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    rank: int
    world: int
    obs_index: np.ndarray   # (m_local,) indices into the caller's observation arrays, line-sorted
    line_start: int         # first global 3D line owned by this rank
    line_stop: int          # one past the last

    @property
    def n_lines(self):
        return self.line_stop - self.line_start


def plan_shards(lines_indices, n_lines: int, world: int) -> list[Shard]:
    li = np.asarray(lines_indices, dtype=np.int64)
    m = li.shape[0]
    if li.size and (li.min() < 0 or li.max() >= n_lines):
        raise ValueError("lines_indices out of range")
    if m > 1 and np.any(li[1:] < li[:-1]):
        order = np.argsort(li, kind="stable")
    else:
        order = np.arange(m, dtype=np.int64)
    line_ptr = np.zeros(n_lines + 1, dtype=np.int64)
    np.cumsum(np.bincount(li, minlength=n_lines), out=line_ptr[1:])
    cuts = [0]
    for r in range(1, world):
        target = (m * r) // world
        # the line boundary closest to the ideal cut, never before the previous cut
        j = int(np.searchsorted(line_ptr, target, side="left"))
        j = min(j, n_lines)
        if j > 0 and (target - line_ptr[j - 1]) <= (line_ptr[j] - target):
            j -= 1
        j = min(max(j, cuts[-1]), n_lines)
        cuts.append(j)
    cuts.append(n_lines)
    shards = []
    for r in range(world):
        l0, l1 = cuts[r], cuts[r + 1]
        shards.append(Shard(r, world, order[line_ptr[l0]:line_ptr[l1]], l0, l1))
    return shards


def local_problem(shard: Shard, camera_indices, lines_indices, n, p0, p1, points_3d):
    """Slice the caller's arrays down to one shard (local line ids start at 0)."""
    idx = shard.obs_index
    ci = np.asarray(camera_indices, dtype=np.int64)[idx]
    li = np.asarray(lines_indices, dtype=np.int64)[idx] - shard.line_start
    pts = np.asarray(points_3d, dtype=np.float64).reshape(-1, 3)[2 * shard.line_start:2 * shard.line_stop]
    return ci, li, np.asarray(n)[idx], np.asarray(p0)[idx], np.asarray(p1)[idx], pts
