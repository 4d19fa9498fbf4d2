"""TEST INFRASTRUCTURE ONLY - float64 CPU statement of the ANALYTIC oriented-3D IoU mode.

The reference has no analytic IoU (its obb_iou is the sampled estimator, instances.py:573-613;
SURVEY.md F2).  BASELINE.json's north_star nevertheless asks for a gravity-aligned BEV
Sutherland-Hodgman kernel, so the product offers it as a second mode and this file is the oracle
that defines it: for two boxes sharing an axis (|a_i . b_j| >= 1 - 1e-6, gravity axis = local Y
preferred), IoU = area(footprint_A ∩ footprint_B) * overlap along the shared axis / union volume,
with box frames recovered from the float32 corners ((v0+v6)/2, v1-v0, v3-v0, v4-v0; vertex order of
boxes.py:756-766).  Pairs without a shared axis return None (the product falls back to the sampled
estimator).  "parity unpinned": there is no reference output for this mode; it is validated against
the sampled estimator's discretisation error only.
"""
import numpy as np


def _frame(c):
    c = c.astype(np.float64)
    cen = 0.5 * (c[0] + c[6])
    ax, half = [], []
    for o in (1, 3, 4):
        e = c[o] - c[0]
        ln = np.sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])
        half.append(0.5 * ln)
        ax.append(e / ln if ln > 0 else np.zeros(3))
    return cen, np.array(ax), np.array(half)


def _clip(poly, ha, hq):
    for e in range(4):
        lim = ha if e < 2 else hq
        sgn = -1.0 if (e & 1) else 1.0
        k = 0 if e < 2 else 1
        out = []
        n = len(poly)
        for i in range(n):
            p, q = poly[i], poly[(i + 1) % n]
            ci, cj = sgn * p[k], sgn * q[k]
            ini, inj = ci <= lim, cj <= lim
            if ini:
                out.append(p)
            if ini != inj:
                t = (lim - ci) / (cj - ci)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
        poly = out
        if not poly:
            break
    return poly


def iou_pair(ca, cb):
    cA, aA, hA = _frame(ca)
    cB, aB, hB = _frame(cb)
    ia = ib = -1
    for a in (1, 0, 2):
        for b in (1, 0, 2):
            if abs(float(aA[a] @ aB[b])) >= 1.0 - 1e-6:
                ia, ib = a, b
                break
        if ia >= 0:
            break
    if ia < 0:
        return None
    u = aA[ia]
    pa, qa, mb, nb = (ia + 1) % 3, (ia + 2) % 3, (ib + 1) % 3, (ib + 2) % 3
    dc = cB - cA
    hb = float(dc @ u)
    oh = min(hA[ia], hb + hB[ib]) - max(-hA[ia], hb - hB[ib])
    if oh <= 0:
        return 0.0
    P, Q = aA[pa], aA[qa]
    cp, cq = float(dc @ P), float(dc @ Q)
    mp, mq = float(aB[mb] @ P) * hB[mb], float(aB[mb] @ Q) * hB[mb]
    np_, nq = float(aB[nb] @ P) * hB[nb], float(aB[nb] @ Q) * hB[nb]
    poly = [(cp - mp - np_, cq - mq - nq), (cp + mp - np_, cq + mq - nq), (cp + mp + np_, cq + mq + nq),
            (cp - mp + np_, cq - mq + nq)]
    poly = _clip(poly, hA[pa], hA[qa])
    area = 0.0
    for i in range(len(poly)):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % len(poly)]
        area += x0 * y1 - x1 * y0
    vi = 0.5 * abs(area) * oh
    vA, vB = 8.0 * hA[0] * hA[1] * hA[2], 8.0 * hB[0] * hB[1] * hB[2]
    return vi / (vA + vB - vi)


def iou_matrix(corners_a, corners_b):
    """float64 [M,N]; AABB-disjoint pairs are 0 (their analytic IoU is exactly 0)."""
    A, B = np.asarray(corners_a, np.float32), np.asarray(corners_b, np.float32)
    loA, hiA, loB, hiB = A.min(1), A.max(1), B.min(1), B.max(1)
    m = np.float32(1e-4)
    cand = np.all((loA[:, None, :] <= hiB[None, :, :] + m) & (loB[None, :, :] <= hiA[:, None, :] + m), axis=2)
    out = np.zeros((A.shape[0], B.shape[0]))
    for i, j in zip(*np.nonzero(cand)):
        v = iou_pair(A[i], B[j])
        assert v is not None, "pair without a shared axis"
        out[i, j] = v
    return out
