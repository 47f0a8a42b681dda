"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares, the
LAMDA loader agrees with the oracle's independent parser, and GPU entry points fail loudly here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import MOLFILE, ROOT
from radex_emcee_b200 import _lib


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "radex_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), s
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)


def test_default_opts():
    o = _lib.default_opts()
    assert (o.stop_rule, o.miniter, o.maxiter) == (0, 10, 200)
    assert o.abs_tol == 1e-16
    assert abs(o.fk_epi / 1.4387768775 - 1) < 1e-9 and abs(o.thc_epi / 3.97289171e-16 - 1) < 1e-8


def test_moldata_matches_oracle_parser(oracle):
    m = _lib.MolData(MOLFILE)
    assert (m.nlev, m.nline, m.npart) == (41, 40, 2)
    assert list(m.partner_id) == [2, 3] and list(m.ncoll) == [820, 820] and list(m.ntemp) == [25, 25]
    np.testing.assert_array_equal(m.eterm, oracle.eterm)
    np.testing.assert_array_equal(m.gstat, oracle.gstat)
    np.testing.assert_array_equal(m.iupp, oracle.iupp)
    np.testing.assert_array_equal(m.ilow, oracle.ilow)
    np.testing.assert_array_equal(m.aeinst, oracle.aeinst)
    np.testing.assert_array_equal(m.xnu, oracle.xnu)      # from level energies, not the GHz column
    np.testing.assert_array_equal(m.spfreq, oracle.spfreq)


def test_moldata_matches_reference_binary():
    """rb_moldata_load (the product's parser) against the tables the reference's own readdata() produced
    (tests/golden/macho_readdata.npz): levels, lines, xnu from the level energies -- bit for bit, both tables."""
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "macho_readdata.npz"))
    for tag, name, parts in (("co", "co.dat", [2, 3]), ("rotor21", "rotor21.dat", [1, 4])):
        m = _lib.MolData(os.path.join(ROOT, "radex_emcee_b200", "data", name))
        assert list(m.partner_id) == parts
        np.testing.assert_array_equal(m.eterm, g[tag + "_eterm"])
        np.testing.assert_array_equal(m.gstat, g[tag + "_gstat"])
        np.testing.assert_array_equal(m.iupp, g[tag + "_iupp"])           # 1-based at the C ABI, like the Fortran
        np.testing.assert_array_equal(m.ilow, g[tag + "_ilow"])
        np.testing.assert_array_equal(m.aeinst, g[tag + "_aeinst"])
        np.testing.assert_array_equal(m.spfreq, g[tag + "_spfreq"])
        np.testing.assert_array_equal(m.eup, g[tag + "_eup"])
        np.testing.assert_array_equal(m.xnu, g[tag + "_xnu"])
    m = _lib.MolData(os.path.join(ROOT, "radex_emcee_b200", "data", "rotor21.dat"))
    assert list(m.ncoll) == [210, 57] and list(m.ntemp) == [9, 6]


def test_moldata_errors(tmp_path):
    with pytest.raises(_lib.RadexB200Error, match="cannot open"):
        _lib.MolData(str(tmp_path / "missing.dat"))
    bad = tmp_path / "bad.dat"
    bad.write_text("!MOLECULE\nX\n!W\n1.0\n!N\n3\n!LEV\n1 0.0 1.0\n")
    with pytest.raises(_lib.RadexB200Error, match="parse error"):
        _lib.MolData(str(bad))


def test_no_cpu_fallback():
    """Without a GPU the context cannot be created and says so; nothing silently computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = _lib.MolData(MOLFILE)
    with pytest.raises(_lib.RadexB200Error):
        _lib.Context(m, 0)
    from radex_emcee_b200.radex import Radex
    R = Radex(species="co", density={"oH2": 750.0, "pH2": 250.0}, column=1e15, temperature=20.0)
    with pytest.raises(_lib.RadexB200Error):
        R.run_radex()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "radex_emcee_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("oracle/radex_oracle.c ro_", "").replace("oracle ro_", ""), f


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in _lib.py against what a C compiler makes of include/radex_b200.h (plain C, no CUDA needed)."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "radex_b200.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("rb_opts %zu\n", sizeof(rb_opts));   P(rb_opts, kernel); P(rb_opts, abs_tol); P(rb_opts, thc_epi); P(rb_opts, park_max); P(rb_opts, lnprob_pipe_min);
  printf("rb_obs %zu\n", sizeof(rb_obs));     P(rb_obs, jup); P(rb_obs, flux); P(rb_obs, eflux);
  printf("rb_source %zu\n", sizeof(rb_source)); P(rb_source, bounds); P(rb_source, tbg); P(rb_source, has_td); P(rb_source, t_d);
  printf("rb_split %zu\n", sizeof(rb_split)); P(rb_split, walkers_per_source); P(rb_split, block); P(rb_split, randomize); P(rb_split, seed);
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.rsplit(" ", 1) for line in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    for name, ct in (("rb_opts", _lib.rb_opts), ("rb_obs", _lib.rb_obs), ("rb_source", _lib.rb_source), ("rb_split", _lib.rb_split)):
        assert int(got[name]) == C.sizeof(ct), name
        for key, off in got.items():
            if key.startswith(name + "."):
                assert int(off) == getattr(ct, key.split(".")[1]).offset, key
