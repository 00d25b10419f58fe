#!/usr/bin/env python
"""Cost model of the LU-SGS sweeps on an n^3 hexahedral block (DESIGN.md §4 "LU-SGS without barriers", §8 item 1).

Today: rows are swept hyperplane by hyperplane (level = i + j + k); a level of w rows moves w * B bytes and cannot start
before the previous one is visible, so   t_sweep = sum_levels max(w * B / BW, hop).
Planned: slabs of thickness s along k, one slab per thread-block cluster, sweep values exchanged through distributed shared
memory inside the cluster (hop_c per level) and only slab faces through L2; slab m runs s levels behind slab m-1:
  t_sweep = max(bytes / BW, ((n + n + s - 2) + (n_slabs - 1) * s) * hop_c).

    python tools/lusgs_model.py [n ...]
Prints the model next to the measured sweep times recorded in profiles/ (B200, this round)."""
import sys

import numpy as np

BW = 6392.8e9            # measured HBM copy bandwidth (MEASURED_PEAKS.json), B/s
B_ROW = (176 + 416 * 3) / 2.0   # bytes per row and sweep: (176 N + 416 F) / 2 sweeps with F = 3 N
HOP = 3.3e-6             # measured dependent hop through L2 (profiles/r01_lusgs_trace.txt)
HOP_CLUSTER = 1.0e-6     # planned: 0.11 us DSMEM signalling (B300_MICROARCH) + the ordered sums of one row
MEASURED_MS = {128: 2.16, 172: 3.45, 200: 4.4, 344: 13.4}   # both sweeps, ms


def level_widths(n):
    """rows per hyperplane i + j + k = l of an n^3 block"""
    c = np.ones(n, dtype=np.int64)
    w = np.convolve(np.convolve(c, c), c)
    return w


def sweep_today(n):
    w = level_widths(n)
    return float(np.maximum(w * B_ROW / BW, HOP).sum())


def sweep_clustered(n, n_clusters=18):
    s = int(np.ceil(n / n_clusters))
    n_slabs = int(np.ceil(n / s))
    chain = (n + n + s - 2) + (n_slabs - 1) * s
    return max(n ** 3 * B_ROW / BW, chain * HOP_CLUSTER), s, chain


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [128, 172, 200, 344]
    print(f"{'n':>5} {'levels':>7} {'bytes/BW ms':>12} {'model ms':>9} {'measured ms':>12} {'clustered ms':>13} {'slab':>5} {'chain':>6}")
    for n in sizes:
        bw_ms = 2e3 * n ** 3 * B_ROW / BW
        today = 2e3 * sweep_today(n)
        cl, s, chain = sweep_clustered(n)
        meas = MEASURED_MS.get(n)
        print(f"{n:5d} {3 * n - 2:7d} {bw_ms:12.2f} {today:9.2f} {meas if meas else float('nan'):12.2f} {2e3 * cl:13.2f} {s:5d} {chain:6d}")
