"""Parity at the BASELINE sizes (SURVEY 8c acceptance, unrelaxed) -- the CUDA path through the C ABI against the committed Monte-Carlo
fixtures tests/golden/atsize_*.npz (made by tests/golden/make_atsize_golden.py from the reference's own per-ray routines / the oracle).

Per slab of origin cells of the full-size level, with K = 8 reference seeds giving per-cell mean mu and standard deviation s:
  (i)   per cell |divQ - mu| <= 3 s sqrt(1 + 1/K): at least as often as the reference's own runs manage under the same statistic
        (leave-one-out control stored in the fixture; s comes from 8 runs, so normal theory's 99.7 % is ~98 % for the reference itself);
        the same at 5 s;
  (ii)  global relative L2 distance from mu no larger than the largest of the reference's own seed-to-seed distances;
  (iii) centre-line L2 against the Burns & Christon exact solution (RMCRT.m:22-38, 368-413) no larger than the reference's worst;
  (iv)  no bias: the mean over the slab agrees within 4 standard errors;
  masks (cellType skip, boundFlux.p, set of written Stencil7 slots): bit-exact.
No slack factors.  The library defaults are used (rayCoherence 0: independent Philox rays per (cell, ray, timestep)); the opt-in batched
modes are held to (i), (ii), (iv) and measured -- not asserted -- on error correlation (test_error_correlation).
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(case):
    f = np.load(os.path.join(GOLD, f"atsize_{case}.npz"))
    return f, json.loads(str(f["meta"]))


def make(**kw):
    from uintah_b200 import RMCRT
    return RMCRT(device=0, **kw)


def accept(d, mu, sd, K, ctl, what, flow=None, draws_per_cell=1.0):
    """criteria (i), (ii), (iv); returns the numbers for the report.  draws_per_cell < 1 (opt-in batched modes only): the cells of a warp
    tile share their rays, so the fractions of (i) are estimated from cells * draws_per_cell independent draws -- three binomial standard
    deviations of that smaller sample are allowed; the default sampling is held to the reference's own minimum without any allowance."""
    mu = mu.astype(np.float64); sd = sd.astype(np.float64)
    sel = np.ones(mu.shape, bool) if flow is None else flow
    z = np.abs(d[sel] - mu[sel]) / (sd[sel] * np.sqrt(1 + 1 / K) + 1e-300)
    f3, f5 = float((z <= 3).mean()), float((z <= 5).mean())
    l2 = float(np.linalg.norm(d[sel] - mu[sel]) / np.linalg.norm(mu[sel]))
    bias = float((d[sel] - mu[sel]).mean())
    se = float(np.sqrt((sd[sel] ** 2).mean() * (1 + 1 / K) / sel.sum()))
    rep = dict(what=what, f3=f3, f3_ref_min=min(ctl["loo_f3"]), f5=f5, f5_ref_min=min(ctl["loo_f5"]), l2=l2, l2_ref_max=max(ctl["loo_l2"]),
               bias_in_se=bias / se)
    print(json.dumps(rep))
    n_eff = sel.sum() * draws_per_cell
    a3 = 0.0 if draws_per_cell >= 1 else 3 * np.sqrt(0.975 * 0.025 / n_eff)
    a5 = 0.0 if draws_per_cell >= 1 else 3 * np.sqrt(0.998 * 0.002 / n_eff)
    assert f3 >= min(ctl["loo_f3"]) - a3, rep
    assert f5 >= min(ctl["loo_f5"]) - a5, rep
    assert l2 <= max(ctl["loo_l2"]) * (1.0 if draws_per_cell >= 1 else 1.0 + 3 / np.sqrt(2 * n_eff)), rep
    return rep


def centre_lines(po, d, N, zi):
    mid = N // 2
    ex = po.burns_christon_exact((np.arange(N) + 0.5) / N)
    return [float(np.sqrt(np.mean((d[zi, mid, :] - ex) ** 2))), float(np.sqrt(np.mean((d[zi, :, mid] - ex) ** 2)))]


@pytest.mark.parametrize("case,half,coh,bm", [("c2", False, 0, 3), ("c3", True, 0, 3), ("c3", False, 0, 3), ("c2", False, 0, 0), ("c2", True, 0, 0),
                                              ("c3", True, 0, 0), ("c2", False, 0, 2), ("c2", False, 2, 3), ("c3", True, 2, 3)])
def test_single_level_at_size(gpu_lib, po, case, half, coh, bm):
    """C2 (128^3) and C3 (256^3) Burns & Christon, 100 rays / cell, float properties; float (8-B) and half-packed (4-B) level records.
    bm = blockedMarch: 3 = the library default (divq_sorted_kernel on the blocked fp16 copy of either record kind), 0 = the one-cell
    kernel on the level's own records (float: no fp16 anywhere), 2 = the blocked copy without the sort"""
    fx, meta = load(case)
    N, K = meta["N"], meta["K"]
    lv = po.burns_christon(N, dtype=np.float32)
    g = make(nDivQRays=100, threshold=0.05, seed=20261018, half_records=half, rayCoherence=coh, blockedMarch=bm)
    g.upload_level(0, lv)
    reps = []
    for si, (z0, z1) in enumerate(meta["slabs"]):
        out = g.trace_divq(0, (0, 0, z0), (N, N, z1), out_lo=(0, 0, z0), out_hi=(N, N, z1), timestep=7)
        d = out["divQ"]
        assert np.all(d != 0) and np.all(out["radiationVolq"] > 0)
        rep = accept(d, fx[f"mu{si}"], fx[f"sd{si}"], K, meta[f"control{si}"], f"{case} half={half} coh={coh} bm={bm} slab [{z0},{z1})",
                     draws_per_cell=1.0 if coh == 0 else 1.0 / 128)
        assert abs(rep["bias_in_se"]) <= (4 if coh == 0 else 4 * np.sqrt(128)), rep     # batched modes: one tile = one draw
        if z0 <= N // 2 < z1 and coh == 0:       # (iii): both centre lines of the plane through cell N/2
            ref = np.array(meta["centreline_L2"])
            mine = centre_lines(po, d, N, N // 2 - z0)
            print(json.dumps(dict(what=f"{case} centre lines", gpu=mine, ref_max=float(ref.max()), ref_mean=float(ref.mean()))))
            assert max(mine) <= ref.max(), (mine, ref.tolist())
        reps.append(rep)
    g.close()


def test_steps_per_ray_at_size(gpu_lib, po):
    """executed ray-cell steps of the first leg at 128^3 agree with the reference's count on the same slab to 0.2 % (Monte-Carlo spread
    of the mean over 3.3e6 rays is 0.04 %)"""
    fx, meta = load("c2")
    N = meta["N"]
    lv = po.burns_christon(N, dtype=np.float32)
    g = make(nDivQRays=100, threshold=0.05, seed=5)
    g.upload_level(0, lv)
    K = meta["K"]
    for si, (z0, z1) in enumerate(meta["slabs"]):
        g.reset_counters()
        g.trace_divq(0, (0, 0, z0), (N, N, z1), out_lo=(0, 0, z0), out_hi=(N, N, z1))
        ref = np.mean(meta["steps"][si * K:(si + 1) * K])
        assert abs(g.counters().nRaySteps - ref) <= 2e-3 * ref, (g.counters().nRaySteps, ref)


@pytest.mark.parametrize("half", [True, False])
def test_data_onion_at_size(gpu_lib, po, half):
    """C4: fine 256^3 / coarse 64^3 (RR 4), halo 4, patch_based ROI; the first 128^3 fine patch, compared on two slabs"""
    fx, meta = load("c4")
    N, K = meta["N"], meta["K"]
    fine = po.burns_christon(N, dtype=np.float32)
    fine.rr = (4, 4, 4)
    coarse = po.coarsen_level(fine, (4, 4, 4))
    g = make(nDivQRays=100, threshold=0.05, seed=77, haloCells=tuple(meta["halo"]), half_records=half)
    g.upload_level(0, coarse)
    g.upload_level(1, fine)
    plo, phi = tuple(meta["patch"][0]), tuple(meta["patch"][1])
    out = g.trace_dataonion(2, plo, phi, out_lo=plo, out_hi=phi, timestep=3)
    d = out["divQ"]
    assert np.all(d != 0)
    for si, (z0, z1) in enumerate(meta["slabs"]):
        rep = accept(d[z0:z1], fx[f"mu{si}"], fx[f"sd{si}"], K, meta[f"control{si}"], f"c4 half={half} slab [{z0},{z1})")
        assert abs(rep["bias_in_se"]) <= 4, rep
    g.close()


@pytest.mark.parametrize("half", [True, False])
def test_boundary_flux_reflections_at_size(gpu_lib, po, half):
    """C5: 256^3 benchmark 2, walls 1000 K with x- / y- mirrors, solveBoundaryFlux + allowReflect, 100 + 100 rays, Threshold 0.001"""
    fx, meta = load("c5")
    N, K = meta["N"], meta["K"]
    y1 = meta["y_range"][1]
    lv = po.parallel_plates(N, dtype=np.float32, mirror_faces=("x-", "y-"))
    g = make(nDivQRays=100, nFluxRays=100, threshold=0.001, solveBoundaryFlux=1, allowReflect=1, seed=31, half_records=half)
    g.upload_level(0, lv)
    for si, (z0, z1) in enumerate(meta["slabs"]):
        plo, phi = (0, 0, z0), (N, y1, z1)
        out = g.trace_divq(0, plo, phi, out_lo=plo, out_hi=phi, timestep=2)
        rep = accept(out["divQ"], fx[f"mu{si}"], fx[f"sd{si}"], K, meta[f"control{si}"], f"c5 half={half} slab [{z0},{z1}) divQ")
        assert abs(rep["bias_in_se"]) <= 4, rep
        S, fmu, fsd = out["boundFlux"], fx[f"fmu{si}"].astype(np.float64), fx[f"fsd{si}"].astype(np.float64)
        assert np.array_equal(S[..., 6], fmu[..., 6])                          # boundFlux.p: bit-exact
        assert np.array_equal(S[..., :6] != 0, fmu[..., :6] != 0)             # set of written faces: bit-exact
        m = fmu[..., :6] != 0
        if m.any():
            z = np.abs(S[..., :6][m] - fmu[..., :6][m]) / (fsd[..., :6][m] * np.sqrt(1 + 1 / K) + 1e-300)
            f3 = float((z <= 3).mean())
            relbias = float((S[..., :6][m].mean() - fmu[..., :6][m].mean()) / fmu[..., :6][m].mean())
            print(json.dumps(dict(what=f"c5 half={half} slab [{z0},{z1}) fluxes", faces=int(m.sum()), f3=f3, rel_bias=relbias)))
            assert f3 >= 0.975, f3           # t distribution with 7 degrees of freedom: P(|t| <= 3) = 0.980
            assert abs(relbias) < 2e-3, relbias
    g.close()


def tile_stats(err, sd, K, tile=(2, 4, 16)):
    """variance of the tile-mean error over the (z, y, x) tiles of a slab, in units of the independent-ray expectation
    mean(s^2) (1 + 1/K) / cells-per-tile; and the normalised nearest-neighbour error correlation along x"""
    tz, ty, tx = tile
    nz, ny, nx = err.shape
    e = err[:nz // tz * tz, :ny // ty * ty, :nx // tx * tx]
    s2 = (sd.astype(np.float64) ** 2)[:e.shape[0], :e.shape[1], :e.shape[2]]
    t = e.reshape(e.shape[0] // tz, tz, e.shape[1] // ty, ty, e.shape[2] // tx, tx).mean(axis=(1, 3, 5))
    expect = s2.mean() * (1 + 1 / K) / (tz * ty * tx)
    zn = err / (sd * np.sqrt(1 + 1 / K))
    return float(t.var() / expect), float((zn[:, :, 1:] * zn[:, :, :-1]).mean())


def test_error_correlation(gpu_lib, po):
    """The price of ray batching, measured at 128^3 x 100 rays against the reference's per-cell mean: variance of the error averaged over
    the 16 x 4 x 2 cells of a warp tile relative to what independent rays give, and the correlation of the errors of x-neighbours.
    Default (rayCoherence 0, the reference's sampling): ratio ~1, correlation ~0 -- asserted.  rayCoherence 1 / 2: reported (of order the
    128 cells of a tile, correlation ~0.9) and asserted only to be what the header says, so the trade stays visible."""
    fx, meta = load("c2")
    N, K = meta["N"], meta["K"]
    z0, z1 = meta["slabs"][0]
    mu, sd = fx["mu0"].astype(np.float64), fx["sd0"].astype(np.float64)
    lv = po.burns_christon(N, dtype=np.float32)
    res = {}
    for coh in (0, 1, 2):
        g = make(nDivQRays=100, threshold=0.05, seed=99, rayCoherence=coh)
        g.upload_level(0, lv)
        ratios, corrs = [], []
        for ts in (1, 2, 3):
            d = g.trace_divq(0, (0, 0, z0), (N, N, z1), out_lo=(0, 0, z0), out_hi=(N, N, z1), timestep=ts)["divQ"]
            r, c = tile_stats(d - mu, sd, K)
            ratios.append(r); corrs.append(c)
        res[coh] = (float(np.mean(ratios)), float(np.mean(corrs)))
        g.close()
    print(json.dumps({"tile_mean_variance_over_independent": {k: v[0] for k, v in res.items()},
                      "x_neighbour_error_correlation": {k: v[1] for k, v in res.items()}}))
    out_dir = os.path.join(os.path.dirname(GOLD), "..", "gpurun_out")
    if os.path.isdir(out_dir):
        json.dump({str(k): {"tile_mean_variance_over_independent": v[0], "x_neighbour_error_correlation": v[1]} for k, v in res.items()},
                  open(os.path.join(out_dir, "error_correlation.json"), "w"))
    assert res[0][0] < 1.5 and abs(res[0][1]) < 0.05, res      # the default is the reference's sampling
    assert res[2][0] > 20 and res[2][1] > 0.5, res              # documented price of the opt-in lockstep mode
