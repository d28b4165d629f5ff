"""shared helpers for the tests"""
import numpy as np


def random_rays(lv, n, rng, lo=None, hi=None):
    """n random (origin cell, direction, physical origin) triples inside [lo,hi) (default: interior)"""
    lo = np.array(lv.interior_lo if lo is None else lo)
    hi = np.array(lv.interior_hi if hi is None else hi)
    cells = np.stack([rng.integers(lo[d], hi[d], n) for d in range(3)], 1).astype(np.int32)
    z = rng.uniform(-1, 1, n)
    th = rng.uniform(0, 2 * np.pi, n)
    r = np.sqrt(1 - z * z)
    dirs = np.stack([r * np.cos(th), r * np.sin(th), z], 1)
    dx = np.array(lv.dx)
    org = np.array(lv.anchor) + dx * cells + dx * rng.uniform(0, 1, (n, 3))
    return cells, dirs, org


def add_random_walls(lv, frac, rng, wall_type=8, abskg=0.7, sigmaT4=None):
    """turn a random fraction of interior cells into intrusions (exercise the cellType masking)"""
    ct = lv.cellType
    interior = np.zeros(ct.shape, bool)
    interior[1:-1, 1:-1, 1:-1] = True
    pick = interior & (rng.uniform(size=ct.shape) < frac)
    ct[pick] = wall_type
    lv.abskg[pick] = abskg
    if sigmaT4 is not None:
        lv.sigmaT4[pick] = sigmaT4
    return int(pick.sum())


def mc_stats(runs):
    """per-cell mean / std over a list of divQ arrays and the seed-to-seed relative L2 spread"""
    a = np.stack(runs)
    mu = a.mean(0)
    sd = a.std(0, ddof=1)
    nrm = np.linalg.norm(mu)
    spread = [np.linalg.norm(x - mu) / nrm for x in a]
    return mu, sd, spread


def loo_control(runs, sel=None):
    """Leave-one-out control of the Monte-Carlo acceptance test: for every oracle run k compare it with the mean / std
    of the OTHER runs exactly the way a candidate is compared.  Returns (frac3, frac5, l2) lists -- the reference's own
    seed-to-seed behaviour under the same statistic."""
    a = np.stack(runs)
    K = len(runs)
    f3, f5, l2 = [], [], []
    for k in range(K):
        others = np.delete(a, k, axis=0)
        mu = others.mean(0)
        sd = others.std(0, ddof=1)
        x = a[k]
        if sel is not None:
            mu_s, sd_s, x_s = mu[sel], sd[sel], x[sel]
        else:
            mu_s, sd_s, x_s = mu, sd, x
        z = np.abs(x_s - mu_s) / (sd_s * np.sqrt(1 + 1 / (K - 1)) + 1e-300)
        f3.append(float((z <= 3.0).mean()))
        f5.append(float((z <= 5.0).mean()))
        l2.append(float(np.linalg.norm(x - mu) / np.linalg.norm(mu)))
    return f3, f5, l2


def candidate_stats(x, runs, sel=None):
    a = np.stack(runs)
    K = len(runs)
    mu = a.mean(0)
    sd = a.std(0, ddof=1)
    if sel is not None:
        z = np.abs(x[sel] - mu[sel]) / (sd[sel] * np.sqrt(1 + 1 / K) + 1e-300)
    else:
        z = np.abs(x - mu) / (sd * np.sqrt(1 + 1 / K) + 1e-300)
    return float((z <= 3.0).mean()), float((z <= 5.0).mean()), float(np.linalg.norm(x - mu) / np.linalg.norm(mu))


def centreline_rms(po, solve, N=41, timesteps=(1, 2, 3)):
    """RMS over the x, y and z centre lines of several solves of sqrt(mean((divQ - exact)^2)) against the Burns & Christon exact
    solution (RMCRT.m:368-413).  With ray batching the cells of one centre line share their ray directions, so the error along ONE line
    is a handful of draws rather than N: same expectation as for independent rays, much wider spread -- hence the average.
    solve(ts) -> divQ over the (N+2)^3 window"""
    sq = []
    mid = N // 2 + 1
    for ts in timesteps:
        d = solve(ts)
        ex = po.centreline_L2(d, N)[2]
        for ln in (d[mid, mid, 1:N + 1], d[mid, 1:N + 1, mid], d[1:N + 1, mid, mid]):
            sq.append(np.mean((ln - ex) ** 2))
    return float(np.sqrt(np.mean(sq)))
