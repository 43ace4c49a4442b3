#!/usr/bin/env python
"""Build tests/golden/madrid_metropolis.npz from the reference's shipped 1DSfM dataset.

Run HERE (the build container), where /root/reference is mounted; the GPU box has
no /root/reference, so the compact fixture is what travels.  The fixture is DATA
(parsed numbers), not reference source.

What is kept (all the rotation-averaging path reads):
  cc            int32[V]     view ids of the connected component        (cc.txt)
  num_images    int          number of lines of list.txt (view id == line index,
                             thirdparty/TheiaSfM/src/theia/io/read_1dsfm.cc:113-160)
  edge_ij       int32[E,2]   (view_id1, view_id2) of every EGs.txt line (read_1dsfm.cc:299-373)
  edge_R        float64[E,9] the row-major 3x3 exactly as parsed from the text
  edge_t        float64[E,3] the position column exactly as parsed
  num_verified_matches int32[E]  #tracks common to both views (read_1dsfm.cc:354-360)
  cov_ij        int32[C,2]   covariance_rot.txt ids                     (src/uncertainty.cpp:200-229)
  cov6          float64[C,6] C00 C11 C22 C01 C02 C12, bit-exact (uint64 text -> double)
  cov_rot       float64[C,3] the refined rotation stored beside it (unused by the solver)
"""
import sys
import numpy as np

SRC = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/datasets/Madrid_Metropolis"
OUT = sys.argv[2] if len(sys.argv) > 2 else __file__.rsplit("/", 1)[0] + "/madrid_metropolis.npz"

cc = np.array(open(f"{SRC}/cc.txt").read().split(), dtype=np.int32)
num_images = sum(1 for l in open(f"{SRC}/list.txt") if l.strip())

# tracks.txt: "<num_tracks>\n" then per track "<n> (view feat)*n"
tok = open(f"{SRC}/tracks.txt").read().split()
pos = 0
num_tracks = int(tok[pos]); pos += 1
view_tracks = {}
for t in range(num_tracks):
    n = int(tok[pos]); pos += 1
    views = tok[pos:pos + 2 * n:2]
    pos += 2 * n
    # Theia's Reconstruction::AddTrack rejects a track that sees one view twice.
    if len(set(views)) != len(views):
        continue
    for v in views:
        view_tracks.setdefault(int(v), set()).add(t)

eg = np.loadtxt(f"{SRC}/EGs.txt", dtype=np.float64)
edge_ij = eg[:, :2].astype(np.int32)
edge_R = np.ascontiguousarray(eg[:, 2:11])
edge_t = np.ascontiguousarray(eg[:, 11:14])
empty = set()
nvm = np.array([len(view_tracks.get(int(a), empty) & view_tracks.get(int(b), empty))
                for a, b in edge_ij], dtype=np.int32)

rows = [l.split() for l in open(f"{SRC}/covariance_rot.txt").read().splitlines()[2:] if l.strip()]
cov_ij = np.array([[int(r[0]), int(r[1])] for r in rows], dtype=np.int32)
bits = np.array([[int(x) for x in r[2:11]] for r in rows], dtype=np.uint64)
vals = bits.view(np.float64)
cov6 = np.ascontiguousarray(vals[:, :6])
cov_rot = np.ascontiguousarray(vals[:, 6:9])

np.savez_compressed(OUT, cc=cc, num_images=np.int64(num_images), edge_ij=edge_ij, edge_R=edge_R,
                    edge_t=edge_t, num_verified_matches=nvm, cov_ij=cov_ij, cov6=cov6, cov_rot=cov_rot)
print("views", len(cc), "images", num_images, "edges", len(edge_ij), "covs", len(cov_ij),
      "matches>=30:", int((nvm >= 30).sum()), "->", OUT)
