"""The reference's behaviour on strong masers, pinned with the CPU oracle (bit-exact to the reference binary): a NaN
escape probability (LVG, tau/2 <= -7: the logarithm of a negative number, radex.so@0xa9c0) makes the next call of matrix()
return EVERY level on the population floor, the relaxed populations then shrink by 0.7f per call and no longer sum to 1.
The CUDA kernels skip the elimination of such calls; this is what they rely on."""
import math

import numpy as np

import bench

F07, F03, MINPOP = float(np.float32(0.7)), float(np.float32(0.3)), 1e-20


def test_nan_escape_probability_puts_every_level_on_the_floor(oracle):
    # a model of the bench sweep that ends at the 200-call cap: T = 849 K, n(H2) = 8.3e3 cm^-3, N(CO) = 3.0e18 cm^-2
    T, nh2, N, i = [848.9765097178501], [8286.817438888651], [3.006119645216707e+18], 0
    oracle.set_physics(T[i], 0.25 * nh2[i], 0.75 * nh2[i])
    oracle.set_column(N[i], 1.0)
    oracle.set_method(2)
    oracle.backrad(bench.TBG)
    nl = oracle.nlev
    floor_calls, proper_calls, prev = 0, 0, None
    for it in range(60):
        # the escape probabilities this call will use: from the relaxed populations the previous call left
        if it > 0:
            x = oracle.xpop.copy()
            up, lo = oracle.iupp - 1, oracle.ilow - 1            # LAMDA numbers levels from 1
            gm, gn = oracle.gstat[up], oracle.gstat[lo]
            xt = oracle.xnu ** 3
            tau = (N[i] / 1e5) * (x[lo] * gm / gn - x[up]) / (float(np.float32(1.0645)) * 8.0 * 3.14159265 * xt / oracle.aeinst)
            beta = np.array([oracle.escprob(t) for t in tau])
            has_nan = bool(np.isnan(beta).any())
            assert has_nan == bool((tau * 0.5 <= -7.0).any())
        oracle.matrix(it)
        x_after = oracle.xpop.copy()
        if it > 0:
            on_floor = np.allclose(x_after, F03 * MINPOP + F07 * np.maximum(MINPOP, prev), rtol=1e-13, atol=0)
            assert on_floor == has_nan, (it, on_floor, has_nan)     # NaN in <=> every level on the floor out
            floor_calls += on_floor
            proper_calls += not on_floor
        prev = x_after
    assert floor_calls > 40 and proper_calls >= 5                  # 87 % of this model's calls
    assert abs(prev.sum() - 1.0) > 0.1                             # the reference's populations are not normalised here
    assert math.isfinite(prev.sum()) and (prev >= MINPOP).all() and len(prev) == nl


def test_share_of_floor_calls_in_the_sweep(oracle):
    """6 % of the sweep's models make such calls for most of their 200 calls: ~9 % of all calls of matrix()."""
    T, nh2, N = bench.draw(400, 0)
    total, floor = 0, 0
    for i in range(400):
        oracle.set_physics(T[i], 0.25 * nh2[i], 0.75 * nh2[i])
        oracle.set_column(N[i], 1.0)
        oracle.set_method(2)
        oracle.backrad(bench.TBG)
        prev = None
        for it in range(40):
            oracle.matrix(it)
            s = oracle.xpop.sum()
            if prev is not None and abs(s - F07 * prev) < 1e-12 * s + 1e-18:
                floor += 1
            prev = s
            total += 1
    assert 0.03 < floor / total < 0.2, floor / total
