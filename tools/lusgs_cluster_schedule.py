#!/usr/bin/env python
"""Prototype of the clustered LU-SGS schedule planned for round 2 — host logic only, runs on CPU.

For an n^3 block in OpenFOAM cell order (i fastest) it builds the (slab, intra-slab level) schedule, checks the two
properties the device protocol relies on, and replays a scalar Gauss-Seidel forward sweep in schedule order to show that
the result is bit-identical to the sequential sweep in cell order (every row subtracts its lower neighbours in ascending
face id, exactly like lusgs.C:230-290):

  (P1) inside a slab every lower neighbour of a row of level l lies in level l - 1 of the same slab  -> one cluster barrier
       per level is enough;
  (P2) the lower neighbour across the slab face lies in the previous slab at intra-slab level l + s - 1, i.e. the previous
       slab produced it s - 1 levels before this slab needs it when slab m runs s levels behind slab m - 1.

    python tools/lusgs_cluster_schedule.py [n] [n_clusters]
"""
import sys

import numpy as np


def schedule(n, n_clusters):
    s = int(np.ceil(n / n_clusters))
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    cell = (i + n * (j + n * k)).ravel()
    slab = (k // s).ravel()
    level = (i + j + k % s).ravel()
    # global step at which a row is processed: slab m starts s levels after slab m - 1
    step = level + slab * s
    return s, cell, slab, level, step, (i.ravel(), j.ravel(), k.ravel())


def check_and_replay(n=24, n_clusters=6, seed=0):
    s, cell, slab, level, step, (i, j, k) = schedule(n, n_clusters)
    idx = lambda a, b, c: a + n * (b + n * c)
    lower = []                      # lower neighbours in ascending face id: (i,j,k-1), (i,j-1,k), (i-1,j,k)
    for (di, dj, dk) in ((0, 0, -1), (0, -1, 0), (-1, 0, 0)):
        ok = (i + di >= 0) & (j + dj >= 0) & (k + dk >= 0)
        nb = np.where(ok, idx(i + di, j + dj, k + dk), -1)
        lower.append(nb)
    pos_of_cell = np.empty(n ** 3, np.int64)
    pos_of_cell[cell] = np.arange(n ** 3)
    for nb in lower:
        has = nb >= 0
        q = pos_of_cell[nb[has]]
        same = slab[q] == slab[has]
        assert (level[q][same] == level[has][same] - 1).all(), "P1 violated"
        assert (slab[q][~same] == slab[has][~same] - 1).all() and (level[q][~same] == level[has][~same] + s - 1).all(), "P2 violated"
        assert (step[q] < step[has]).all(), "a neighbour would be needed before it exists"
    # replay: x_c = (b_c - sum_lower a_cq x_q) * rD_c, sequential in cell order vs in schedule order
    rng = np.random.default_rng(seed)
    a = [rng.standard_normal(n ** 3) for _ in lower]
    b, rD = rng.standard_normal(n ** 3), 1.0 / (4.0 + rng.random(n ** 3))

    def sweep(order):
        x = np.full(n ** 3, np.nan)
        for c in order:
            acc = b[c]
            for nb, coef in zip(lower, a):
                if nb[c] >= 0:
                    acc -= coef[c] * x[nb[c]]
            x[c] = acc * rD[c]
        return x

    seq = sweep(np.arange(n ** 3))
    sched = sweep(cell[np.lexsort((cell, step))])
    assert np.array_equal(seq, sched), "schedule order changes the result"
    chain = step.max() + 1
    return s, chain


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    nc = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    s, chain = check_and_replay(n, nc)
    print(f"n = {n}: slab thickness {s}, {int(np.ceil(n / s))} slabs, dependent chain {chain} steps (hyperplane sweep: {3 * n - 2} levels through L2); "
          f"P1, P2 hold and the replayed sweep is bit-identical to the sequential one")
