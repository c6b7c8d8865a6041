"""TEST INFRASTRUCTURE ONLY -- numpy restatement of pieces of the reference that the compiled oracle
(oracle/_ref) cannot pin because the reference itself is racy there.

particleVelocityToGridThread / centeredParamsToGridThread write their validity flags into
std::vector<bool> grids from several ThreadPool ranges at once (flipsolver2d.cpp:1373-1375,1429);
vector<bool> packs 64 flags per word, so flags that share a word with a range boundary are lost
nondeterministically. The functions below restate the flag rule itself in float32 arithmetic
(numpy float32 ops are IEEE single, no FMA), for small cases."""
import numpy as np

F = np.float32


def bspline(v):
    """simmath::bSpline (mathfuncs.cpp:32-37)"""
    v = np.abs(v).astype(F)
    a = (F(0.75) - v * v) * (v < F(0.5)).astype(F)
    h = F(1.5) - v
    b = (F(0.5) * h * h) * ((v >= F(0.5)) & (v < F(1.5))).astype(F)
    return (a + b).astype(F)


def quadratic_bspline(x, y):
    """simmath::quadraticBSpline (mathfuncs.cpp:39-43); bSpline(0) = 0.75"""
    return (bspline(x) * bspline(y) * F(0.75)).astype(F)


def p2g_validity(pos, I, J):
    """Cells (i, j) that receive at least one particle with weightU > 1e-9 && weightV > 1e-9
    (flipsolver2d.cpp:1352-1368). Returns a bool (I, J) array; the U flag lives at U(i, j), the V
    flag at V(i, j)."""
    out = np.zeros((I, J), bool)
    px, py = pos[:, 0].astype(F), pos[:, 1].astype(F)
    ci, cj = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            i, j = ci + di, cj + dj
            ok = (i >= 0) & (i < I) & (j >= 0) & (j < J)
            fi, fj = i.astype(F), j.astype(F)
            wu = quadratic_bspline(px - fi, py - (fj + F(0.5)))
            wv = quadratic_bspline(px - (fi + F(0.5)), py - fj)
            hit = ok & (wu > F(1e-9)) & (wv > F(1e-9))
            out[i[hit], j[hit]] = True
    return out


def centered_known(pos, I, J, water=True):
    """knownCenteredParams rule: w > 1e-9 (water, flipsolver2d.cpp:1416-1423) or |w| > 1e-6
    (nbflip / smoke, nbflipsolver.cpp:481-489)."""
    out = np.zeros((I, J), bool)
    px, py = pos[:, 0].astype(F), pos[:, 1].astype(F)
    ci, cj = np.floor(px).astype(np.int64), np.floor(py).astype(np.int64)
    for di in (-1, 0, 1, 2):
        for dj in (-1, 0, 1, 2):
            i, j = ci + di, cj + dj
            ok = (i >= 0) & (i < I) & (j >= 0) & (j < J)
            w = quadratic_bspline(px - i.astype(F), py - j.astype(F))
            hit = ok & ((w > F(1e-9)) if water else (np.abs(w) > F(1e-6)))
            out[i[hit], j[hit]] = True
    return out
