import sys, numpy as np
d = np.load(sys.argv[1])
for m in (2, 1, 3):
    g = {k: d["got_%s_%d" % (k, m)] for k in ("xpop", "tex", "tau", "surf", "niter", "status")}
    r = {k: d["ref_%s_%d" % (k, m)] for k in ("xpop", "tex", "tau", "surf", "niter", "status")}
    P = d["P_%d" % m]
    n = len(P)
    cg, cr = g["niter"] < 200, r["niter"] < 200
    print("method", m, "n", n, "conv both", (cg & cr).sum(), "cap both", (~cg & ~cr).sum(), "mixed", (cg ^ cr).sum(),
          "niter diff hist (conv both):", np.bincount(np.minimum(np.abs(g["niter"] - r["niter"])[cg & cr], 10)))
    for label, sel in (("conv-both", cg & cr), ("cap-both", ~cg & ~cr), ("mixed", cg ^ cr)):
        if not sel.any():
            continue
        x, xr = g["xpop"][sel], r["xpop"][sel]
        with np.errstate(all="ignore"):
            ex = np.where(xr > 1e-9, np.abs(x - xr) / xr, 0).max(axis=1)
            s, sr = g["surf"][sel][:, :12], r["surf"][sel][:, :12]
            bright = np.abs(sr) > 1e-6 * np.nanmax(np.abs(sr), axis=1, keepdims=True)
            es = np.where(bright, np.abs(s - sr) / np.abs(sr), 0)
            es = np.nan_to_num(es, nan=0).max(axis=1)
        q = lambda e: "med %.1e 90%% %.1e 99%% %.1e max %.1e" % (np.median(e), np.quantile(e, .9), np.quantile(e, .99), e.max())
        print("  %-10s n=%4d  pops(x>1e-9): %s | flux J<=12: %s | frac flux>1e-5: %.3f" % (label, sel.sum(), q(ex), q(es), (es > 1e-5).mean()))
        worst = np.argsort(-es)[:3]
        for w in worst:
            i = np.where(sel)[0][w]
            print("      worst: T=%.1f n=%.2e N=%.2e niter g/r %d/%d es %.1e ex %.1e" % (P[i, 0], P[i, 1], P[i, 2], g["niter"][i], r["niter"][i], es[w], ex[w]))
