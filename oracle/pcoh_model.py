"""TEST INFRASTRUCTURE ONLY — executable model of the algorithm the CUDA Rips engine runs.

The CUDA kernel (tda_eeg_audio_b200/csrc/rips_small.cu) does not reduce a boundary matrix and
does not port Ripser's heap-based column reduction.  It runs *persistent cohomology by cocycle
annotation* in one sweep over the sorted edge list, with every cocycle stored as per-vertex
adjacency bitmasks so that a whole "group" of triangles is evaluated with one XOR/AND:

  sweep edges e=(i,j) in filtration order (d ascending, index descending)
    * union-find says "merging"  -> H0 pair, nothing else to do
    * otherwise e is a cycle-creating edge.  G = adj[i] & adj[j] is the set of apexes v whose
      triangles (i,j,v) enter the filtration right after e (e is their youngest edge); the
      triangle index is monotone in v, so "descending index" == "descending v".
        - G empty      -> a real H1 class is born: new cocycle phi = indicator(e)
        - G non-empty  -> e and its top triangle form a zero-persistence (apparent) pair.
                          For each live cocycle l, x_l = (phi_l.row[i] ^ phi_l.row[j]) & G holds
                          phi_l(i,v)+phi_l(j,v) for every apex at once; phi_l(e) := x_l[v_top]
                          (the unique extension that stays a cocycle on the top triangle) and
                          c_l = x_l ^ (phi_l(e) ? G : 0) is the coboundary of phi_l on all the
                          other triangles of the group.  If every c_l is 0 (the common case)
                          nothing dies.  Otherwise walk apexes downwards: at apex v the youngest
                          live l with c_l[v]=1 dies (pair: its birth edge, triangle (i,j,v)); every
                          other l' with c_l'[v]=1 absorbs it: phi_l' ^= phi_l, c_l' ^= c_l.
    * edges of equal length (a tie run) are handled in the exact simplexwise order instead:
      all edges of the run first, then the union of their groups in descending triangle index.

This file is the CPU statement of that sweep (python ints as bitmasks).  It is checked against
oracle/rips_naive.py in tests/ so that the *algorithm* is validated without a GPU; the kernel is
then checked against oracle/rips_cpu.cpp on the GPU.
"""
from __future__ import annotations

import numpy as np


def _c2(i):
    return i * (i - 1) // 2


def _c3(i):
    return i * (i - 1) * (i - 2) // 6


def tri_index(i, j, v):
    a, b, c = sorted((i, j, v), reverse=True)
    return _c3(a) + _c2(b) + c


def rips_h01_pcoh(dm, thresh=np.inf, stats=None):
    dm = np.asarray(dm)
    n = dm.shape[0]
    thr = np.float32(thresh)
    edges = []
    for i in range(1, n):
        for j in range(i):
            v = np.float32(dm[j, i])
            if v <= thr:
                edges.append((float(v), _c2(i) + j, i, j))
    edges.sort(key=lambda e: (e[0], -e[1]))
    m = len(edges)

    comp = list(range(n))
    eldest = list(range(n))
    adj = [0] * n
    live = []  # dicts: rank, phi(list of n ints)
    born = []  # (rank, death_tri, death_diam) filled on death
    h0 = []
    max_live = 0
    n_xor = 0
    n_events = 0

    def h0_step(d, idx, i, j):
        a, b = comp[i], comp[j]
        if a == b:
            return False
        ea, eb = eldest[a], eldest[b]
        h0.append((min(ea, eb), idx, d))
        for v in range(n):
            if comp[v] == a:
                comp[v] = b
        eldest[b] = max(ea, eb)
        return True

    def new_class(rank, i, j):
        phi = [0] * n
        phi[i] |= 1 << j
        phi[j] |= 1 << i
        rec = {"rank": rank, "phi": phi, "death": None}
        live.append(rec)
        born.append(rec)

    def kill_at(cands, tri, diam):
        """cands: live records with coboundary 1 on the triangle; youngest dies, others absorb."""
        nonlocal n_xor, n_events
        n_events += 1
        dying = max(cands, key=lambda r: r["rank"])
        dying["death"] = (tri, diam)
        live.remove(dying)
        for r in cands:
            if r is not dying:
                r["phi"] = [x ^ y for x, y in zip(r["phi"], dying["phi"])]
                n_xor += 1
        return dying

    r = 0
    while r < m:
        r1 = r + 1
        while r1 < m and edges[r1][0] == edges[r][0]:
            r1 += 1
        d = edges[r][0]
        if r1 - r == 1:
            # ---------------- fast path: one edge, its group evaluated with bitmasks
            _, idx, i, j = edges[r]
            G = adj[i] & adj[j]
            merging = h0_step(d, idx, i, j)
            if not merging:
                if G == 0:
                    new_class(r, i, j)
                else:
                    vtop = G.bit_length() - 1
                    cm = {}
                    for rec in live:
                        x = (rec["phi"][i] ^ rec["phi"][j]) & G
                        if (x >> vtop) & 1:
                            rec["phi"][i] |= 1 << j
                            rec["phi"][j] |= 1 << i
                            x ^= G
                        cm[id(rec)] = x
                    while True:
                        top = max((c.bit_length() - 1 for c in cm.values()), default=-1)
                        if top < 0:
                            break
                        cands = [rec for rec in live if (cm[id(rec)] >> top) & 1]
                        dying = kill_at(cands, tri_index(i, j, top), d)
                        cd = cm.pop(id(dying))
                        for rec in cands:
                            if rec is not dying:
                                cm[id(rec)] ^= cd
            adj[i] |= 1 << j
            adj[j] |= 1 << i
        else:
            # ---------------- tie run: exact simplexwise order, with apparent pairs inside the run
            # pass 1: all edges of the run enter (H0 decisions in rank order)
            run = []
            runadj = [0] * n
            for p in range(r, r1):
                _, idx, i, j = edges[p]
                merging = h0_step(d, idx, i, j)
                adj[i] |= 1 << j
                adj[j] |= 1 << i
                runadj[i] |= 1 << j
                runadj[j] |= 1 << i
                run.append((p, idx, i, j, merging))

            def earlier(x, y, idx_e):
                # is edge (x,y) earlier in the filtration than the run edge with index idx_e ?
                if not (runadj[x] >> y) & 1:
                    return True
                a, b = max(x, y), min(x, y)
                return _c2(a) + b > idx_e

            # pass 2: a cycle-creating run edge whose FIRST cofacet (largest apex among the
            # triangles present after the run) has it as youngest edge forms an apparent pair:
            # no slot; the cocycles are extended over it when that triangle is reached.
            defv = {}
            groups = {}
            for (p, idx, i, j, merging) in run:
                cand = adj[i] & adj[j]
                g = 0
                v = 0
                c = cand
                while c:
                    if (c & 1) and earlier(i, v, idx) and earlier(j, v, idx):
                        g |= 1 << v
                    c >>= 1
                    v += 1
                groups[(i, j)] = g
                if merging:
                    continue
                vt = cand.bit_length() - 1
                if cand and earlier(i, vt, idx) and earlier(j, vt, idx):
                    defv[(i, j)] = vt
                else:
                    new_class(p, i, j)
            tris = []
            for (p, idx, i, j, merging) in run:
                g = groups[(i, j)]
                v = 0
                while g:
                    if g & 1:
                        tris.append((tri_index(i, j, v), i, j, v))
                    g >>= 1
                    v += 1
            tris.sort(reverse=True)

            def val(rec, x, y):
                return (rec["phi"][x] >> y) & 1

            for (t, i, j, v) in tris:
                if defv.get((i, j)) == v:
                    # defining triangle of the apparent edge (i,j): phi(i,j) := phi(i,v)+phi(j,v)
                    for rec in live:
                        if val(rec, i, v) ^ val(rec, j, v):
                            rec["phi"][i] |= 1 << j
                            rec["phi"][j] |= 1 << i
                    continue
                cands = [rec for rec in live if val(rec, i, j) ^ val(rec, i, v) ^ val(rec, j, v)]
                if cands:
                    kill_at(cands, t, d)
        max_live = max(max_live, len(live))
        r = r1

    dg0 = [(0.0, dd) for (_, _, dd) in h0 if dd != 0.0]
    pr0 = [(v, e) for (v, e, dd) in h0 if dd != 0.0]
    for v in sorted(eldest[c] for c in set(comp)):
        dg0.append((0.0, np.inf))
        pr0.append((v, -1))
    dg1, pr1 = [], []
    for rec in sorted(born, key=lambda q: -q["rank"]):
        bd, bidx = edges[rec["rank"]][0], edges[rec["rank"]][1]
        if rec["death"] is None:
            dg1.append((bd, np.inf))
            pr1.append((bidx, -1))
        elif rec["death"][1] > bd:
            dg1.append((bd, rec["death"][1]))
            pr1.append((bidx, rec["death"][0]))
    if stats is not None:
        stats.update(max_live=max_live, n_xor=n_xor, n_events=n_events, n_born=len(born))

    def arr(x, dt):
        return np.asarray(x, dtype=dt).reshape(-1, 2)

    return {"dgms": [arr(dg0, np.float64), arr(dg1, np.float64)],
            "pairs": [arr(pr0, np.int64), arr(pr1, np.int64)]}
