#!/usr/bin/env python
"""Round-2 design prototype (host side, numpy): symmetric-half TILED storage of the stencil matrix.

Every edge (i, j) carries ONE symmetric 3x3 block S (the Laplacian stencil: H_ii += S, H_jj += S, H_ij = H_ji = -S).
Views are cut into G groups; an edge belongs to tile (group(i), group(j)) with group(i) <= group(j).  A tile is streamed
once: d_e = S_e (x_i - x_j); y_i += d_e (row side, edges sorted by i inside the tile -> segmented reduction) and
y_j -= d_e (column side, through a per-tile permutation that sorts the tile's edges by j).  Per tile that yields partial
y slices for its two groups; a fixed-order pass adds the <= G partials of every view.

This script builds the layout for the bench workload, checks y = H x against the half-edge (full storage) product, and
prints the bytes a pass would stream next to today's 104 B/edge, plus the segment statistics the CUDA kernel will see."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from globalsfmpy_b200 import viewgraph as vg  # noqa: E402


def build(num_views, ei, ej, group_size):
    lo, hi = np.minimum(ei, ej), np.maximum(ei, ej)
    G = (num_views + group_size - 1) // group_size
    ga, gb = lo // group_size, hi // group_size
    tile = ga * G + gb
    order = np.lexsort((hi, lo, tile))             # tile, then row (i), then column (j)
    return dict(G=G, order=order, lo=lo[order], hi=hi[order], tile=tile[order], group_size=group_size)


def spmv_tiled(L, S, x):
    """y = offdiag-part of H x + the edge contributions to the diagonal, via d = S (x_i - x_j)."""
    d = np.einsum("eab,eb->ea", S[L["order"]], x[L["lo"]] - x[L["hi"]])
    y = np.zeros_like(x)
    np.add.at(y, L["lo"], d)
    np.add.at(y, L["hi"], -d)
    return y


def main():
    N, E = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (10000, 1000000)
    g = vg.synthetic_pose_graph(N, E, seed=56, noise_deg=1.0, outlier_fraction=0.1)
    rng = np.random.default_rng(0)
    A = rng.normal(size=(g.num_edges, 3, 3))
    S = A @ np.transpose(A, (0, 2, 1))            # symmetric PSD blocks
    x = rng.normal(size=(N, 3))
    # reference: full (half-edge) storage
    y_ref = np.zeros_like(x)
    np.add.at(y_ref, g.edge_i, np.einsum("eab,eb->ea", S, x[g.edge_i] - x[g.edge_j]))
    np.add.at(y_ref, g.edge_j, np.einsum("eab,eb->ea", S, x[g.edge_j] - x[g.edge_i]))
    for group_size in (1250, 2500, 5000):
        L = build(N, g.edge_i, g.edge_j, group_size)
        y = spmv_tiled(L, S, x)
        err = np.abs(y - y_ref).max() / np.abs(y_ref).max()
        tiles, counts = np.unique(L["tile"], return_counts=True)
        # row segments inside a tile: runs of equal lo; column segments: distinct hi per tile
        row_runs = 1 + np.count_nonzero((np.diff(L["lo"]) != 0) | (np.diff(L["tile"]) != 0))
        col_keys = L["tile"].astype(np.int64) * N + L["hi"]
        col_runs = len(np.unique(col_keys))
        bytes_edge = 48 + 2 + 2 + 4                # S(6 doubles) + u16 row, u16 col within the group + u32 column-order permutation
        partial = len(tiles) * 2 * group_size * 24  # partial y slices written + read once each
        stream = g.num_edges * bytes_edge + len(tiles) * 2 * group_size * 24 + 2 * partial
        print(f"group_size {group_size:5d}: G = {L['G']:2d}, tiles = {len(tiles):3d} (edges/tile min {counts.min()} max {counts.max()}), "
              f"rel err {err:.1e}, row segments {row_runs} (avg {g.num_edges / row_runs:.1f} edges), column segments {col_runs}, "
              f"x slices per CTA {2 * group_size * 24 / 1024:.0f} KB, bytes per pass {stream / 1e6:.1f} MB "
              f"({stream / g.num_edges:.1f} B/edge; full storage today: {104 * g.num_edges / 1e6:.0f} MB)")


if __name__ == "__main__":
    main()
