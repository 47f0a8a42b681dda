"""Generate tests/golden/macho_*.npz by running the REFERENCE'S OWN compiled RADEX routines.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference and x86-64) and only when asked
to (RADEX_RUN_REF_BINARY=1, see oracle/macho_ref.py: it executes the reference's machine code); the fixtures it writes
are committed so that every other run -- the default test-suite, the GPU box -- checks against them instead.

What is recorded (all produced by emcee/pyradex/radex/radex.so through oracle/macho_ref.py):
  macho_escprob.npz : escprob(tau) for the three geometries on a tau grid incl. branch edges
  macho_backrad.npz : backi/totalb for several tbg
  macho_readdata.npz: readdata() itself -- level/line tables, crate, ctot, totdens -- for both synthetic tables
  macho_solve_rotor21.npz : the solve loop for the second table (21 levels, partners H2 and e)
  macho_solve.npz   : for random config-2 style parameter draws, the state after pyradex's
                      run_radex loop (emcee/pyradex/core.py:896-925, reuse_last=False) around the
                      binary's matrix(): niter, xpop[41], tex[40], taul[40]; plus a few chained
                      reuse_last=True solves (the drivers' mode, emcee/emcee_radex.py:127).
Inputs the binary cannot produce itself (readdata needs libgfortran I/O): level/line tables and
crate/ctot come from oracle/radex_oracle.c's restatement of readdata on the synthetic co.dat and
are stored in the fixture as well, so the fixture is self-contained.

    python oracle/make_golden.py
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import macho_ref  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

MOLFILE = os.path.join(ROOT, "radex_emcee_b200", "data", "co.dat")
ROTOR = os.path.join(ROOT, "radex_emcee_b200", "data", "rotor21.dat")
OUT = os.path.join(ROOT, "tests", "golden")


def draw_params(rng, n, tbg):
    """Config-2 draws: log n~U[2,7], log T~U[log tbg,3], log N~U[15.5,19.5], 10<logN-logn<17.5."""
    out = []
    while len(out) < n:
        ln, lt, lN = rng.uniform(2, 7), rng.uniform(np.log10(tbg), 3), rng.uniform(15.5, 19.5)
        if 10.0 < lN - ln < 17.5:
            out.append((10 ** lt, 10 ** ln, 10 ** lN))
    return np.array(out)


def main():
    os.makedirs(OUT, exist_ok=True)
    o = Oracle(MOLFILE)
    r = macho_ref.RefRadex()
    r.load_tables(o.nlev, o.nline, o.eterm, o.gstat, o.iupp, o.ilow, o.aeinst, o.xnu, o.spfreq, o.eup)
    nl, nn = o.nlev, o.nline

    # ---- escprob -------------------------------------------------------------------------
    taus = np.concatenate([np.logspace(-7, 7, 281), -np.logspace(-7, 1.2, 83),
                           [0.0, 0.02, 0.019999999, 0.0200000001, 14.0, 13.999999, 14.000001, 0.2, 0.1999999,
                            100.0, 100.00001, 0.1 / 3, 0.0333333, 50.0 / 3, 16.6667, -14.0, -20.0]])
    beta = np.array([[r.escprob(t, m) for t in taus] for m in (1, 2, 3)])
    np.savez(os.path.join(OUT, "macho_escprob.npz"), tau=taus, methods=np.array([1, 2, 3]), beta=beta)

    # ---- backrad -------------------------------------------------------------------------
    tbgs = np.array([2.7315, 2.73, 2.7315 * 4.6345, 2.7315 * 5.243, 10.926, 0.05, 300.0])
    backi = []
    for t in tbgs:
        r.backrad(t)
        backi.append(r.dview("backi", macho_ref.MAXLINE)[:nn].copy())
        assert (r.dview("totalb", macho_ref.MAXLINE)[:nn] == backi[-1]).all()
        assert (r.dview("trj", macho_ref.MAXLINE)[:nn] == t).all()
    np.savez(os.path.join(OUT, "macho_backrad.npz"), tbg=tbgs, backi=np.array(backi), xnu=o.xnu)

    # ---- solves --------------------------------------------------------------------------
    rng = np.random.default_rng(20170914)
    cases = []
    for tbg, method, k in ((2.7315, 2, 10), (10.926, 2, 14), (2.7315 * 4.6345, 2, 8), (2.7315, 1, 5), (2.7315, 3, 5)):
        for (T, n, N) in draw_params(rng, k, tbg):
            cases.append((T, n, N, tbg, method))
    # the reference test-suite's own settings (test_radex.py:99-115,175-200), with fixed OPR 3
    cases += [(30.0, 1e4, 1e14, 2.73, 2), (20.0, 1e3, 1e15, 2.7315, 2), (25.0, 1e4, 1e14, 2.7315, 2)]
    cases = np.array(cases)
    res = dict(xpop=[], tex=[], tau=[], niter=[], crate=[], ctot=[], totdens=[], stub_calls=[])
    xr = r.dview("xpop", macho_ref.MAXLEV)
    tr = r.dview("tex", macho_ref.MAXLINE)
    ur = r.dview("taul", macho_ref.MAXLINE)

    def setup(T, n, N, tbg, method):
        o.set_physics(T, 0.25 * n, 0.75 * n)
        r.load_rates(o.crate, o.ctot, o.totdens)
        r.backrad(tbg)
        r.set_scalar("tkin", T)
        r.set_scalar("cdmol", N)
        r.set_scalar("deltav", 1e5)
        r.set_int("method", int(method))
        d = r.dview("density", 9)
        d[:] = 0
        d[1], d[2] = 0.25 * n, 0.75 * n

    for (T, n, N, tbg, method) in cases:
        setup(T, n, N, tbg, method)
        it = r.run_pyradex_loop(reuse_last=False)
        res["niter"].append(it)
        res["xpop"].append(xr[:nl].copy())
        res["tex"].append(tr[:nn].copy())
        res["tau"].append(ur[:nn].copy())
        res["crate"].append(o.crate.copy())
        res["ctot"].append(o.ctot.copy())
        res["totdens"].append(o.totdens)

    # chained history (reuse_last=True): walk through the first 12 LVG cases in order without reset
    chain = dict(xpop=[], tex=[], tau=[], niter=[])
    idx = [i for i, c in enumerate(cases) if c[4] == 2][:12]
    setup(*cases[idx[0]])
    r.run_pyradex_loop(reuse_last=False)
    for i in idx[1:]:
        setup(*cases[i])
        it = r.run_pyradex_loop(reuse_last=True)
        chain["niter"].append(it)
        chain["xpop"].append(xr[:nl].copy())
        chain["tex"].append(tr[:nn].copy())
        chain["tau"].append(ur[:nn].copy())

    np.savez_compressed(
        os.path.join(OUT, "macho_solve.npz"),
        cases=cases, niter=np.array(res["niter"]), xpop=np.array(res["xpop"]), tex=np.array(res["tex"]),
        tau=np.array(res["tau"]), crate=np.array(res["crate"]), ctot=np.array(res["ctot"]),
        totdens=np.array(res["totdens"]),
        chain_idx=np.array(idx[1:]), chain_niter=np.array(chain["niter"]), chain_xpop=np.array(chain["xpop"]),
        chain_tex=np.array(chain["tex"]), chain_tau=np.array(chain["tau"]),
        eterm=o.eterm, gstat=o.gstat, iupp=o.iupp, ilow=o.ilow, aeinst=o.aeinst, xnu=o.xnu)
    print("cases", len(cases), "niter", res["niter"])
    print("chain niter", chain["niter"])
    print("imports the binary called:", sorted(set(r.img.calls)))
    make_rotor(rng)
    make_readdata()


def make_readdata():
    """macho_readdata.npz: the reference's own readdata() (LAMDA parse, temperature interpolation incl. the clamps at
    both ends of the grid, partner mix, detailed balance, row sums) on the two synthetic tables: what it leaves in
    /imolec/, /rmolec/, /radi/ and crate / ctot / totdens for a list of (tkin, densities).  One fresh image per file:
    the binary keeps the previous file's collision tables in its COMMON blocks."""
    out = {}
    for tag, path, mixes in (("co", MOLFILE, ([0, 2.5e3, 7.5e3, 0, 0, 0, 0], [0, 1e6, 0, 0, 0, 0, 0], [0, 3.0, 1e2, 0, 0, 0, 0])),
                             ("rotor21", ROTOR, ([3e4, 0, 0, 5.0, 0, 0, 0], [1e3, 0, 0, 0, 0, 0, 0], [0, 0, 0, 40.0, 0, 0, 0]))):
        r = macho_ref.RefRadex()
        temps = [1.5, 2.0, 2.0000001, 9.99, 10.0, 12.85, 50.0, 77.7, 333.3, 499.9, 500.0, 2999.0, 3000.0, 5000.0, 9999.0]
        cases, crate, ctot, totdens = [], [], [], []
        for d in mixes:
            for T in temps:
                r.readdata(path, T, d)
                t = r.tables()
                cases.append([T] + list(d))
                crate.append(t["crate"])
                ctot.append(t["ctot"])
                totdens.append(t["totdens"])
        for k in ("eterm", "gstat", "iupp", "ilow", "aeinst", "eup", "xnu", "spfreq"):
            out["%s_%s" % (tag, k)] = t[k]
        out["%s_amass" % tag] = np.array(t["amass"])
        out["%s_cases" % tag] = np.array(cases)
        out["%s_crate" % tag] = np.array(crate)
        out["%s_ctot" % tag] = np.array(ctot)
        out["%s_totdens" % tag] = np.array(totdens)
        print("readdata", tag, "cases", len(cases), "imports called:", sorted(set(r.img.calls)))
    np.savez_compressed(os.path.join(OUT, "macho_readdata.npz"), **out)


def make_rotor(rng):
    """macho_solve_rotor21.npz: the same loop through the binary for the second synthetic table (21 levels, partners
    H2 and e on their own temperature grids): pins matrix()/lubksb for a matrix size other than CO's and for the
    collider mix of SURVEY.md 8(f) rank 4."""
    o = Oracle(ROTOR)
    r = macho_ref.RefRadex()
    r.load_tables(o.nlev, o.nline, o.eterm, o.gstat, o.iupp, o.ilow, o.aeinst, o.xnu, o.spfreq, o.eup)
    nl, nn = o.nlev, o.nline
    cases = []
    for tbg, method, k in ((2.7315, 2, 12), (10.926, 2, 8), (2.7315, 1, 5), (2.7315, 3, 5)):
        for _ in range(k):
            lt, ln, lN = rng.uniform(np.log10(max(tbg, 5.0)), 2.7), rng.uniform(2.5, 7.0), rng.uniform(11.5, 15.5)
            ne = 0.0 if rng.uniform() < 0.3 else 10 ** rng.uniform(-2.0, 2.0)
            cases.append((10 ** lt, 10 ** ln, ne, 10 ** lN, tbg, method))
    cases = np.array(cases)
    res = dict(xpop=[], tex=[], tau=[], niter=[], crate=[])
    xr, tr, ur = r.dview("xpop", macho_ref.MAXLEV), r.dview("tex", macho_ref.MAXLINE), r.dview("taul", macho_ref.MAXLINE)
    for (T, n, ne, N, tbg, method) in cases:
        d7 = np.array([n, 0, 0, ne, 0, 0, 0.0])
        o.set_physics_dens(T, d7)
        r.load_rates(o.crate, o.ctot, o.totdens)
        r.backrad(tbg)
        r.set_scalar("tkin", T)
        r.set_scalar("cdmol", N)
        r.set_scalar("deltav", 1e5)
        r.set_int("method", int(method))
        d = r.dview("density", 9)
        d[:] = 0
        d[:7] = d7
        # fresh state per case, as a new Radex object has it
        xr[:nl] = 0
        tr[:nn] = 0
        ur[:nn] = 0
        res["niter"].append(r.run_pyradex_loop(reuse_last=False))
        res["xpop"].append(xr[:nl].copy())
        res["tex"].append(tr[:nn].copy())
        res["tau"].append(ur[:nn].copy())
        res["crate"].append(o.crate.copy())
    np.savez_compressed(os.path.join(OUT, "macho_solve_rotor21.npz"), cases=cases, niter=np.array(res["niter"]),
                        xpop=np.array(res["xpop"]), tex=np.array(res["tex"]), tau=np.array(res["tau"]),
                        crate=np.array(res["crate"]))
    print("rotor21 cases", len(cases), "niter", res["niter"])


if __name__ == "__main__":
    main()
