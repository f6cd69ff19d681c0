"""TEST INFRASTRUCTURE ONLY — CPU oracle, never imported by the product path.

Ground-truth Vietoris-Rips persistence (H0, H1, Z/2) by explicit boundary-matrix
reduction.  This is the *definition-level* implementation: it materialises every
vertex, edge and triangle, orders them with Ripser's simplexwise total order and
runs the textbook left-to-right column reduction.  It is O(#triangles^2) in the
worst case and meant for N <= ~64 points.

parity status: "parity unpinned" — the reference repository holds no golden
vector for the ripser boundary; `ripser` itself is a third-party
dependency (`ripser>=0.6`, /root/reference/requirements.txt:5) that is absent
from this image.  The semantics restated here follow SURVEY.md Appendix A.1:

* values are float32 (ripser.py casts the distance matrix to float32 before the
  C++ call; reference call sites /root/reference/scripts/utils.py:131,140 and
  /root/reference/scripts/tda_eeg_classification_v2.py:170-175);
* only the upper triangle D[i, j], i < j is read;
* edge (i > j) has index C(i,2)+j, triangle (a > b > c) has index
  C(a,3)+C(b,2)+c; a simplex's diameter is the max of its float32 edge lengths;
* filtration total order: diameter ascending, then dimension ascending, then
  combinatorial index DESCENDING;
* simplices with diameter > thresh are absent (NaN compares false => absent);
* H0 diagram: (0, d) per merging edge with d != 0 in filtration order, then one
  (0, inf) per surviving component; H1 diagram: rows in DESCENDING birth-edge
  filtration order, zero-persistence pairs dropped, unpaired cycles (d, inf).

Because the persistence pairing is unique once the simplex order is fixed, any
other correct algorithm (oracle/rips_cpu.cpp, the CUDA engine) must reproduce
these (birth simplex, death simplex) index pairs bit for bit.
"""
from __future__ import annotations

import numpy as np


def _c2(i: int) -> int:
    return i * (i - 1) // 2


def _c3(i: int) -> int:
    return i * (i - 1) * (i - 2) // 6


def upper_triangle_f32(dm: np.ndarray) -> np.ndarray:
    """float32 matrix whose [i, j] (i<j) entries are what ripser would read."""
    dm = np.asarray(dm)
    n = dm.shape[0]
    out = np.zeros((n, n), dtype=np.float32)
    iu = np.triu_indices(n, 1)
    out[iu] = dm[iu].astype(np.float32)
    out.T[iu] = out[iu]
    return out


def rips_h01_naive(dm, thresh=np.inf):
    """Return dict(dgms=[H0,H1] float64 (k,2); pairs=[H0,H1] int64 (k,2)).

    pairs[0][k] = (birth vertex, death edge index) ; essential -> death = -1.
    pairs[1][k] = (birth edge index, death triangle index) ; essential -> -1.
    H0 birth vertex follows the elder rule under the total order above (vertex
    with the larger index is the elder one; the younger component's eldest
    vertex is the one that dies).
    """
    d = upper_triangle_f32(dm)
    n = d.shape[0]
    thr = np.float32(thresh) if np.isfinite(thresh) else np.float32(np.inf)

    # ---- simplices -------------------------------------------------------
    edges = []  # (diam, index, i, j)
    for i in range(n):
        for j in range(i):
            v = d[j, i]
            if v <= thr:  # NaN -> False -> edge absent
                edges.append((float(v), _c2(i) + j, i, j))
    # order: diameter ascending, index descending
    edges.sort(key=lambda e: (e[0], -e[1]))
    edge_pos = {e[1]: p for p, e in enumerate(edges)}
    have = np.zeros((n, n), dtype=bool)
    for _, _, i, j in edges:
        have[i, j] = have[j, i] = True

    tris = []
    for a in range(n):
        for b in range(a):
            if not have[a, b]:
                continue
            for c in range(b):
                if have[a, c] and have[b, c]:
                    diam = max(d[b, a], d[c, a], d[c, b])
                    tris.append((float(diam), _c3(a) + _c2(b) + c, a, b, c))
    tris.sort(key=lambda t: (t[0], -t[1]))

    # ---- H0: reduce the edge boundary matrix over vertices -----------------
    # vertex filtration position: all diameter 0, index descending
    vpos = {v: (n - 1 - v) for v in range(n)}
    low_owner = {}
    h0_pairs, h0_dgm = [], []
    positive_edge = []
    ecols = []
    for (diam, idx, i, j) in edges:
        col = (1 << vpos[i]) ^ (1 << vpos[j])
        while col:
            low = col.bit_length() - 1
            o = low_owner.get(low)
            if o is None:
                break
            col ^= ecols[o]
        ecols.append(col)
        if col:
            low = col.bit_length() - 1
            low_owner[low] = len(ecols) - 1
            positive_edge.append(False)
            h0_pairs.append((n - 1 - low, idx, diam))
        else:
            positive_edge.append(True)
    dg0 = [(0.0, dd) for (_, _, dd) in h0_pairs if dd != 0.0]
    pr0 = [(v, e) for (v, e, dd) in h0_pairs if dd != 0.0]
    killed = {v for (v, _, _) in h0_pairs}
    for v in range(n):
        if v not in killed:
            dg0.append((0.0, np.inf))
            pr0.append((v, -1))

    # ---- H1: reduce the triangle boundary matrix over edges ---------------
    low_owner = {}
    tcols = []
    death_of_edge = {}
    for (diam, idx, a, b, c) in tris:
        col = 0
        for (x, y) in ((a, b), (a, c), (b, c)):
            col ^= 1 << edge_pos[_c2(x) + y]
        while col:
            low = col.bit_length() - 1
            o = low_owner.get(low)
            if o is None:
                break
            col ^= tcols[o]
        tcols.append(col)
        if col:
            low = col.bit_length() - 1
            low_owner[low] = len(tcols) - 1
            death_of_edge[low] = (idx, diam)
    dg1, pr1 = [], []
    for p in range(len(edges) - 1, -1, -1):
        if not positive_edge[p]:
            continue
        diam, idx, _, _ = edges[p]
        if p in death_of_edge:
            tidx, tdiam = death_of_edge[p]
            if tdiam > diam:
                dg1.append((diam, tdiam))
                pr1.append((idx, tidx))
        else:
            dg1.append((diam, np.inf))
            pr1.append((idx, -1))

    def arr(x, dt):
        return np.asarray(x, dtype=dt).reshape(-1, 2)

    return {
        "dgms": [arr(dg0, np.float64), arr(dg1, np.float64)],
        "pairs": [arr(pr0, np.int64), arr(pr1, np.int64)],
        "num_edges": len(edges),
    }
