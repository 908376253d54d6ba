"""Wavefront schedule of SPAIR's autoregressive cell sweep (host logic, no device code).

The reference visits the Hc x Wc cells strictly one after another in row-major order
(``models.py:68``) because each cell's heads consume the sampled latents of earlier cells
through ``_get_sequential_context`` (``models.py:292-320``).  With look-back radius L the
neighbours of (h, w) are (h-a, w+b) for a in 1..L, |b| <= L and (h, w-b) for b in 1..L, so
all cells with the same ``t = w + (L+1) * h`` are mutually independent and every dependency
has a smaller ``t``: the sweep needs ``Wc + (L+1)(Hc-1)`` steps instead of ``Hc*Wc``
(31 vs 121 at the default 11x11 grid, 46 vs 256 at 16x16, 94 vs 1024 at 32x32).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np


def context_offsets(n_lookback: int = 1) -> List[Tuple[int, int]]:
    """Neighbour offsets (dh, dw) in the reference's concat order (``models.py:292-307``):
    rows -L..0 outer, columns -L..L inner, minus the last L+1 entries (the cell itself and the
    cells to its right).  L=1 -> (-1,-1), (-1,0), (-1,+1), (0,-1)."""
    L = int(n_lookback)
    full = [(dh, dw) for dh in range(-L, 1) for dw in range(-L, L + 1)]
    return full[:-(L + 1)]


@dataclass
class WavefrontSchedule:
    Hc: int
    Wc: int
    n_lookback: int
    offsets: List[Tuple[int, int]]
    order: np.ndarray      # [HW] cell ids (h*Wc+w) in wavefront-major order
    wf_pos: np.ndarray     # [HW] position of each cell in ``order``
    starts: np.ndarray     # [T+1] start of each wavefront in ``order``
    missing: np.ndarray    # [HW, n_nb] bool, indexed by POSITION: neighbour falls outside the grid

    @property
    def n_wavefronts(self) -> int:
        return len(self.starts) - 1

    @property
    def max_cells(self) -> int:
        return int(np.max(np.diff(self.starts)))

    def cells_of(self, t: int) -> np.ndarray:
        return self.order[self.starts[t]:self.starts[t + 1]]


def build_schedule(Hc: int, Wc: int, n_lookback: int = 1) -> WavefrontSchedule:
    L = int(n_lookback)
    offsets = context_offsets(L)
    hh, ww = np.meshgrid(np.arange(Hc), np.arange(Wc), indexing="ij")
    t = (ww + (L + 1) * hh).reshape(-1)
    cell = (hh * Wc + ww).reshape(-1)
    # stable sort by wavefront index, ties by cell id (i.e. by h)
    order = cell[np.lexsort((cell, t))].astype(np.int32)
    t_sorted = t[order]
    T = int(t_sorted[-1]) + 1
    starts = np.searchsorted(t_sorted, np.arange(T + 1)).astype(np.int64)
    wf_pos = np.empty(Hc * Wc, np.int32)
    wf_pos[order] = np.arange(Hc * Wc, dtype=np.int32)
    missing = np.zeros((Hc * Wc, len(offsets)), bool)
    for pos, c in enumerate(order):
        h, w = divmod(int(c), Wc)
        for s, (dh, dw) in enumerate(offsets):
            nh, nw = h + dh, w + dw
            inside = 0 <= nh < Hc and 0 <= nw < Wc
            missing[pos, s] = not inside
            if inside and not t[nh * Wc + nw] < t[c]:
                raise AssertionError("wavefront schedule violates a context dependency")
    return WavefrontSchedule(Hc, Wc, L, offsets, order, wf_pos, starts, missing)
