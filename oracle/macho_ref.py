"""Run the reference's own compiled RADEX routines (Mach-O x86-64) in this Linux container.

TEST INFRASTRUCTURE ONLY (never imported by the product path).

The reference ships RADEX only as a macOS f2py extension,
``/root/reference/emcee/pyradex/radex/radex.so``; the Fortran source is not
vendored.  The machine code is x86-64 / System-V ABI, the same ABI as this container, so this
module maps the image's ``__TEXT``/``__DATA`` segments into anonymous memory, applies the dyld
rebase/bind opcode streams itself, points libm/libc imports at the Linux libraries and every other
import (Python C-API, libgfortran I/O) at a stub, and then calls the *pure compute* routines

    escprob_(tau)         radex.so@0xa9c0
    backrad_()            radex.so@0x1be30    (tbg>0 branch: no I/O)
    matrix_(niter, conv)  radex.so@0x17f70    (-> lubksb_/sgeir_/sgefa_/sgesl_)

directly, exactly as pyradex does through f2py (emcee/pyradex/core.py:854,910,1024).  COMMON-block
member addresses are obtained the way f2py gets them: by calling the image's own
``f2pyinit<block>_`` routines with a callback.  ``readdata_`` needs libgfortran's formatted I/O
and is NOT run; the tables it would fill (crate/ctot, level and line data) are written into the
COMMON blocks by the caller, which is what pins only matrix/escprob/backrad to the binary.

Used by ``oracle/make_golden.py`` to generate ``tests/golden/macho_*.npz`` and by
``tests/test_oracle_vs_binary.py`` (skipped when /root/reference is absent, e.g. on the GPU box).
"""
from __future__ import annotations

import ctypes
import ctypes.util
import os
import struct

import numpy as np

REF_SO = "/root/reference/emcee/pyradex/radex/radex.so"

_libc = ctypes.CDLL(None, use_errno=True)
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")

PROT_RWX = 0x1 | 0x2 | 0x4
MAP_PRIVATE, MAP_ANONYMOUS, MAP_NORESERVE = 0x02, 0x20, 0x4000

# dimensions of this build's radex.inc
MAXLEV, MAXLINE, MAXPART = 2999, 99999, 9


def _uleb(buf, p):
    r = s = 0
    while True:
        b = buf[p]
        p += 1
        r |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            return r, p


def _sleb(buf, p):
    r = s = 0
    while True:
        b = buf[p]
        p += 1
        r |= (b & 0x7F) << s
        s += 7
        if not b & 0x80:
            if b & 0x40:
                r -= 1 << s
            return r, p


class MachoImage:
    """Minimal loader for one MH_BUNDLE with LC_DYLD_INFO_ONLY fixups."""

    def __init__(self, path=REF_SO, verbose=False):
        self.f = open(path, "rb").read()
        self.verbose = verbose
        self.segs = []
        self.syms = {}
        self.calls = []          # names of stubbed imports that were actually called
        self._keep = []          # keep ctypes callbacks / buffers alive
        self._parse()
        self._map()
        self._rebase()
        self._bind(self.bind_off, self.bind_size, lazy=False)
        self._bind(self.lazy_off, self.lazy_size, lazy=True)

    # ---- parsing -------------------------------------------------------------------------
    def _parse(self):
        f = self.f
        magic, cpu, _, ftype, ncmds, _, _, _ = struct.unpack("<IiiIIIII", f[:32])
        if magic != 0xFEEDFACF or cpu != 0x01000007:
            raise RuntimeError("not a 64-bit x86-64 Mach-O image")
        off = 32
        for _ in range(ncmds):
            cmd, csz = struct.unpack("<II", f[off:off + 8])
            if cmd == 0x19:
                name = f[off + 8:off + 24].rstrip(b"\0").decode()
                vmaddr, vmsize, fileoff, filesize = struct.unpack("<QQQQ", f[off + 24:off + 56])
                self.segs.append((name, vmaddr, vmsize, fileoff, filesize))
            elif cmd in (0x22, 0x80000022):
                (self.rebase_off, self.rebase_size, self.bind_off, self.bind_size, _, _,
                 self.lazy_off, self.lazy_size, _, _) = struct.unpack("<10I", f[off + 8:off + 48])
            elif cmd == 0x2:
                symoff, nsyms, stroff, _ = struct.unpack("<4I", f[off + 8:off + 24])
                for i in range(nsyms):
                    strx, typ, sect, _, val = struct.unpack("<IBBHQ", f[symoff + 16 * i:symoff + 16 * i + 16])
                    nm = f[stroff + strx:f.index(b"\0", stroff + strx)].decode()
                    if (typ & 0xE0) == 0 and (typ & 0x0E) == 0x0E:
                        self.syms[nm] = val
            off += csz
        self.vmsize = max(v + s for _, v, s, _, _ in self.segs)

    def _map(self):
        _libc.mmap.restype = ctypes.c_void_p
        _libc.mmap.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                               ctypes.c_int, ctypes.c_long]
        base = _libc.mmap(None, self.vmsize, PROT_RWX, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0)
        if base in (None, ctypes.c_void_p(-1).value):
            raise OSError(ctypes.get_errno(), "mmap failed")
        self.base = base
        for name, vmaddr, vmsize, fileoff, filesize in self.segs:
            if name == "__LINKEDIT" or filesize == 0:
                continue
            ctypes.memmove(base + vmaddr, self.f[fileoff:fileoff + filesize], filesize)

    def _rd64(self, addr):
        return ctypes.c_uint64.from_address(addr).value

    def _wr64(self, addr, v):
        ctypes.c_uint64.from_address(addr).value = v & 0xFFFFFFFFFFFFFFFF

    def _rebase(self):
        buf = self.f[self.rebase_off:self.rebase_off + self.rebase_size]
        p = 0
        seg = segoff = 0
        while p < len(buf):
            b = buf[p]
            p += 1
            op, imm = b & 0xF0, b & 0x0F
            if op == 0x00:
                break
            elif op == 0x10:
                pass
            elif op == 0x20:
                seg = imm
                segoff, p = _uleb(buf, p)
            elif op == 0x30:
                v, p = _uleb(buf, p)
                segoff += v
            elif op == 0x40:
                segoff += imm * 8
            elif op in (0x50, 0x60):
                n = imm
                if op == 0x60:
                    n, p = _uleb(buf, p)
                for _ in range(n):
                    a = self.base + self.segs[seg][1] + segoff
                    self._wr64(a, self._rd64(a) + self.base)
                    segoff += 8
            elif op == 0x70:
                a = self.base + self.segs[seg][1] + segoff
                self._wr64(a, self._rd64(a) + self.base)
                v, p = _uleb(buf, p)
                segoff += 8 + v
            elif op == 0x80:
                cnt, p = _uleb(buf, p)
                skip, p = _uleb(buf, p)
                for _ in range(cnt):
                    a = self.base + self.segs[seg][1] + segoff
                    self._wr64(a, self._rd64(a) + self.base)
                    segoff += 8 + skip
            else:
                raise RuntimeError("bad rebase opcode %#x" % b)

    # ---- import resolution ---------------------------------------------------------------
    _LIBM = ("exp", "log", "log10", "pow", "sqrt")
    _LIBC = ("memset", "memcpy", "memcmp", "strlen", "strcmp", "strncpy", "malloc", "free",
             "snprintf", "sprintf")

    def _resolve(self, name):
        bare = name[1:] if name.startswith("_") else name
        if bare in self._LIBM:
            return ctypes.cast(getattr(_libm, bare), ctypes.c_void_p).value
        if bare in self._LIBC:
            return ctypes.cast(getattr(_libc, bare), ctypes.c_void_p).value
        if bare == "__bzero":
            return ctypes.cast(_libc.bzero, ctypes.c_void_p).value
        if bare in ("__stack_chk_guard", "__stderrp", "_Py_NoneStruct") or bare.startswith("PyExc_") \
                or bare.endswith("_Type"):
            buf = ctypes.create_string_buffer(256)
            self._keep.append(buf)
            return ctypes.addressof(buf)

        def stub(a, b, c, d, _n=bare):
            self.calls.append(_n)
            if _n in ("_gfortran_stop_string", "__stack_chk_fail"):
                raise RuntimeError("reference binary called " + _n)
            return 0

        cb = ctypes.CFUNCTYPE(ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                              ctypes.c_void_p)(stub)
        self._keep.append(cb)
        return ctypes.cast(cb, ctypes.c_void_p).value

    def _bind(self, off, size, lazy):
        buf = self.f[off:off + size]
        p = 0
        seg = segoff = 0
        name = None
        addend = 0
        cache = getattr(self, "_cache", {})
        self._cache = cache

        def do_bind():
            if name not in cache:
                cache[name] = self._resolve(name)
            self._wr64(self.base + self.segs[seg][1] + segoff, cache[name] + addend)

        while p < len(buf):
            b = buf[p]
            p += 1
            op, imm = b & 0xF0, b & 0x0F
            if op == 0x00:            # DONE (lazy streams have one per entry)
                if not lazy:
                    break
            elif op in (0x10, 0x30):  # dylib ordinal imm / special
                pass
            elif op == 0x20:
                _, p = _uleb(buf, p)
            elif op == 0x40:
                e = buf.index(b"\0", p)
                name = buf[p:e].decode()
                p = e + 1
            elif op == 0x50:
                pass
            elif op == 0x60:
                addend, p = _sleb(buf, p)
            elif op == 0x70:
                seg = imm
                segoff, p = _uleb(buf, p)
            elif op == 0x80:
                v, p = _uleb(buf, p)
                segoff += v
            elif op == 0x90:
                do_bind()
                segoff += 8
            elif op == 0xA0:
                do_bind()
                v, p = _uleb(buf, p)
                segoff += 8 + v
            elif op == 0xB0:
                do_bind()
                segoff += 8 + imm * 8
            elif op == 0xC0:
                cnt, p = _uleb(buf, p)
                skip, p = _uleb(buf, p)
                for _ in range(cnt):
                    do_bind()
                    segoff += 8 + skip
            else:
                raise RuntimeError("bad bind opcode %#x" % b)

    def addr(self, sym):
        return self.base + self.syms[sym]


class RefRadex:
    """The reference's RADEX COMMON blocks + compute routines, as pyradex sees them through f2py."""

    BLOCKS = {
        # block -> ordered member names (f2py docstring of the module)
        "cphys": ["density", "tkin", "tbg", "cdmol", "deltav", "totdens"],
        "collie": ["crate", "ctot", "xpop"],
        "radi": ["xnu", "taul", "tex", "backi", "totalb", "spfreq", "trj"],
        "imolec": ["nlev", "nline", "ncoll", "npart", "ntemp", "iupp", "ilow"],
        "rmolec": ["amass", "eterm", "gstat", "aeinst", "eup"],
        "setup": ["radat", "method", "version", "logfile"],
        "freq": ["fmin", "fmax"],
        "dbg": ["debug"],
    }

    def __init__(self, verbose=False):
        self.img = MachoImage(verbose=verbose)
        self.a = {}
        for blk, names in self.BLOCKS.items():
            got = []

            def setup(*args, _got=got):
                _got.extend(args)
                return 0

            n = len(names)
            cbt = ctypes.CFUNCTYPE(ctypes.c_long, *([ctypes.c_void_p] * n))
            cb = cbt(setup)
            fn = ctypes.CFUNCTYPE(None, ctypes.c_void_p)(self.img.addr("_f2pyinit%s_" % blk))
            fn(ctypes.cast(cb, ctypes.c_void_p))
            if len(got) != n:
                raise RuntimeError("f2pyinit%s_ passed %d members, expected %d" % (blk, len(got), n))
            for nm, ad in zip(names, got):
                self.a[nm] = ad
        self._escprob = ctypes.CFUNCTYPE(ctypes.c_double, ctypes.POINTER(ctypes.c_double))(
            self.img.addr("_escprob_"))
        self._matrix = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int))(
            self.img.addr("_matrix_"))
        self._backrad = ctypes.CFUNCTYPE(None)(self.img.addr("_backrad_"))

    # ---- typed views of COMMON members ---------------------------------------------------
    def dview(self, name, n):
        return np.ctypeslib.as_array((ctypes.c_double * n).from_address(self.a[name]))

    def iview(self, name, n):
        return np.ctypeslib.as_array((ctypes.c_int32 * n).from_address(self.a[name]))

    def set_scalar(self, name, v):
        ctypes.c_double.from_address(self.a[name]).value = float(v)

    def get_scalar(self, name):
        return ctypes.c_double.from_address(self.a[name]).value

    def set_int(self, name, v):
        ctypes.c_int32.from_address(self.a[name]).value = int(v)

    def get_int(self, name):
        return ctypes.c_int32.from_address(self.a[name]).value

    # ---- routines ------------------------------------------------------------------------
    def escprob(self, tau, method=2):
        self.set_int("method", method)
        return self._escprob(ctypes.byref(ctypes.c_double(tau)))

    def backrad(self, tbg):
        self.set_scalar("tbg", tbg)
        self._backrad()

    def matrix(self, niter):
        conv = ctypes.c_int(0)
        self._matrix(ctypes.byref(ctypes.c_int(niter)), ctypes.byref(conv))
        return conv.value

    # ---- what readdata would have filled ---------------------------------------------------
    def load_tables(self, nlev, nline, eterm, gstat, iupp, ilow, aeinst, xnu, spfreq, eup):
        """Write level/line tables (1-based iupp/ilow) into /imolec/, /rmolec/, /radi/."""
        self.set_int("nlev", nlev)
        self.set_int("nline", nline)
        self.dview("eterm", MAXLEV)[:nlev] = eterm
        self.dview("gstat", MAXLEV)[:nlev] = gstat
        self.iview("iupp", MAXLINE)[:nline] = iupp
        self.iview("ilow", MAXLINE)[:nline] = ilow
        self.dview("aeinst", MAXLINE)[:nline] = aeinst
        self.dview("eup", MAXLINE)[:nline] = eup
        self.dview("xnu", MAXLINE)[:nline] = xnu
        self.dview("spfreq", MAXLINE)[:nline] = spfreq

    def load_rates(self, crate, ctot, totdens):
        """crate[i, j] = rate i -> j (0-based, nlev x nlev); Fortran crate(i+1, j+1), column-major."""
        nlev = self.get_int("nlev")
        cr = self.dview("crate", MAXLEV * MAXLEV).reshape(MAXLEV, MAXLEV)  # cr[j, i] == crate(i+1, j+1)
        cr[:nlev, :nlev] = np.asarray(crate).T
        self.dview("ctot", MAXLEV)[:nlev] = ctot
        self.set_scalar("totdens", totdens)

    def run_pyradex_loop(self, reuse_last=False, miniter=10, maxiter=200, abs_tol=1e-16, rel_tol=1e-8, trace=None):
        """The python loop of emcee/pyradex/core.py:896-925 around the binary's matrix()."""
        # level_population is the full 2999-long COMMON array in pyradex; the sums below run over
        # all of it (numpy pairwise summation), which fixes the rounding of the 1e-16 stop test.
        xpop = self.dview("xpop", MAXLEV)
        nlev = self.get_int("nlev")
        it = 1 if reuse_last else 0
        last = xpop.copy()
        while True:
            if it >= maxiter:
                break
            self.matrix(it)
            if trace is not None:
                trace.append(xpop[:nlev].copy())
            level_diff = np.abs(last - xpop)
            with np.errstate(all="ignore"):
                frac_level_diff = level_diff / xpop
            if ((level_diff.sum() < abs_tol) or (frac_level_diff.sum() < rel_tol)) and it > miniter:
                break
            last = xpop.copy()
            it += 1
        return it


def available():
    return os.path.exists(REF_SO) and os.uname().machine == "x86_64"


if __name__ == "__main__":
    r = RefRadex(verbose=True)
    for m, nm in ((2, "lvg"), (1, "sphere"), (3, "slab")):
        print(nm, [r.escprob(t, m) for t in (1e-3, 0.5, 5.0, 20.0, 200.0, -0.5)])
    print("stub calls:", r.img.calls)
