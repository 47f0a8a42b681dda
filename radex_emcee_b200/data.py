"""Readers for the drivers' flux tables (reference: emcee/emcee_radex.py:183-240 read_data/get_source,
emcee/emcee_radex_2comp.py:247-279) without pandas/astropy tables.

``flux.dat`` rows: SOURCE z D_L line_width  11 x (CO flux, err)  2 x (CI flux, err)   (30 tokens)
``flux_for2p.dat`` rows add T_dust after D_L (31 tokens).  '#' lines and blank lines are skipped,
so the commented-out source of the 2-component file is dropped as in the reference.
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

DATA_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")


def _num(tok):
    try:
        return float(tok)
    except ValueError:          # pd.to_numeric(errors='coerce')
        return float("nan")


def read_data(filename, two_component=None):
    """Return an ordered mapping source -> dict of columns (the reference returns ``df.T``)."""
    rows = []
    with open(filename) as f:
        for line in f:
            s = line.strip()
            if s and not s.startswith("#"):
                rows.append(s.split())
    if not rows:
        raise ValueError("no data rows in %s" % filename)
    ncol = len(rows[0])
    if any(len(r) != ncol for r in rows):
        raise ValueError("Number of columns in data rows does not match the expected number of columns")
    if two_component is None:
        two_component = (ncol - 8) % 2 == 1
    fixed = ["SOURCE", "z", "D_L"] + (["T_d"] if two_component else []) + ["line_width"]
    nco = (ncol - len(fixed) - 4) // 2
    cols = list(fixed)
    for i in range(nco):
        cols += ["CO_J_%d" % (i + 1), "eCO_J_%d" % (i + 1)]
    cols += ["CI_1", "eCI_1", "CI_2", "eCI_2"]
    if len(cols) != ncol:
        raise ValueError("Number of columns in data rows does not match the expected number of columns")
    out = OrderedDict()
    for r in rows:
        out[r[0]] = {c: _num(t) for c, t in zip(cols[1:], r[1:])}
    return out


def get_source(source, data):
    """(z, line_width, Jup, flux, eflux) -- plus T_d as second item when the table has it."""
    row = data[source]
    keys = [k for k in row if "CO" in k and "eCO" not in k]
    sel = [(jlow + 1, row[k], row["e" + k]) for jlow, k in enumerate(keys) if np.isfinite(row[k])]
    jup = np.array([s[0] for s in sel], dtype=np.int64)
    flux = np.array([s[1] for s in sel], dtype=np.float64)
    eflux = np.array([s[2] for s in sel], dtype=np.float64)
    if "T_d" in row:
        return row["z"], row["T_d"], row["line_width"], jup, flux, eflux
    return row["z"], row["line_width"], jup, flux, eflux
