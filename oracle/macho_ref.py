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
``f2pyinit<block>_`` routines with a callback.

``readdata_`` (radex.so@0x1cf90; core.py:570,744,887) is run as well: it reaches libgfortran only for
OPEN / list-directed and ``(a)`` READs / CLOSE of the LAMDA file, and ``GfortranIO`` below serves exactly those
(st_parameter layouts read off the disassembly: flags @0, unit @4; OPEN file @0x30 / length @0x2c; READ flag 0x80 =
list-directed, 0x1000 = format @0x48 / length @0x50).  That pins the parse, the temperature interpolation, the
partner mix and the detailed balance -- ``crate``/``ctot`` -- to the reference's own machine code.

SAFETY.  The image is third-party machine code from an untrusted tree.  It is mapped and run only when the caller
opts in (environment variable RADEX_RUN_REF_BINARY=1: ``oracle/make_golden.py`` and the two ``*_live_binary``
tests), only if its SHA-256 equals the pinned digest of the file that was disassembled for SURVEY.md, with its
segments mapped W^X (``__TEXT`` read+execute, ``__DATA`` read+write: never writable and executable at once) and
with every import bound to a stub in this file -- the image has no libc of its own.  The default test-suite relies on
the checked-in ``tests/golden/*.npz`` only.
"""
from __future__ import annotations

import ctypes
import ctypes.util
import os
import struct

import numpy as np

REF_SO = "/root/reference/emcee/pyradex/radex/radex.so"
REF_SHA256 = "7cec685080ae12c462322e282461178564f13a00c644e63b1fcd2d3c58224d58"
OPT_IN_ENV = "RADEX_RUN_REF_BINARY"

_libc = ctypes.CDLL(None, use_errno=True)
_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")

PROT_R, PROT_W, PROT_X = 0x1, 0x2, 0x4
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


class GfortranIO:
    """What readdata_ needs of libgfortran.3's I/O: OPEN(unit, file, status='old', err=), READ(unit,*) items,
    READ(unit,'(a)') string, CLOSE.  Records are the lines of the file; a list-directed READ starts on a new record,
    takes blank/comma separated items across as many records as it needs and leaves the file after the last record it
    touched; a READ without items skips one record.  WRITEs (debug prints) are swallowed."""

    def __init__(self, img):
        self.img = img
        self.units = {}
        self.cur = None
        self.error = None       # first exception raised inside a callback (ctypes would only print it)
        CB = ctypes.CFUNCTYPE(ctypes.c_long, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p)
        def guard(fn):
            def run(*a):
                try:
                    return fn(*a)
                except Exception as e:      # noqa: BLE001 -- reported by RefRadex.readdata after the call returns
                    if self.error is None:
                        self.error = e
                    return 0
            return run

        self._cbs = {"_gfortran_st_open": CB(guard(self.st_open)), "_gfortran_st_close": CB(guard(self.st_close)),
                     "_gfortran_st_read": CB(guard(self.st_read)), "_gfortran_st_read_done": CB(guard(self.st_read_done)),
                     "_gfortran_transfer_real": CB(guard(self.transfer_real)),
                     "_gfortran_transfer_integer": CB(guard(self.transfer_integer)),
                     "_gfortran_transfer_character": CB(guard(self.transfer_character))}
        self.table = {k: ctypes.cast(v, ctypes.c_void_p).value for k, v in self._cbs.items()}

    @staticmethod
    def _i32(addr):
        return ctypes.c_int32.from_address(addr).value

    def st_open(self, p, *_):
        unit, flen = self._i32(p + 4), self._i32(p + 0x2C)
        fptr = ctypes.c_uint64.from_address(p + 0x30).value
        name = ctypes.string_at(fptr, flen).decode("latin-1").rstrip()
        try:
            with open(name, "r", errors="replace") as f:
                self.units[unit] = {"lines": f.read().split("\n"), "idx": 0, "name": name}
        except OSError:
            ctypes.c_int32.from_address(p).value |= 1      # LIBRETURN_ERROR: the err= branch
        return 0

    def st_close(self, p, *_):
        self.units.pop(self._i32(p + 4), None)
        return 0

    def st_read(self, p, *_):
        flags, unit = self._i32(p), self._i32(p + 4)
        if unit not in self.units:
            raise RuntimeError("READ from unit %d, which is not open" % unit)
        self.cur = {"u": self.units[unit], "list": bool(flags & 0x80), "touched": False, "tokens": []}
        if not self.cur["list"]:
            # formatted: the edit descriptors readdata uses are (a) and (i1,a); one record, consumed left to right
            if not flags & 0x1000:
                raise RuntimeError("READ that is neither list-directed nor formatted")
            fptr, flen = ctypes.c_uint64.from_address(p + 0x48).value, self._i32(p + 0x50)
            fmt = ctypes.string_at(fptr, flen).decode("latin-1").strip().lower()
            import re
            items = [re.fullmatch(r"([ai])(\d*)", t.strip()) for t in fmt.strip("()").split(",")]
            if not fmt.startswith("(") or any(m is None for m in items):
                raise RuntimeError("unsupported FORMAT %r" % fmt)
            self.cur.update(fmt=[(m.group(1), int(m.group(2) or 0)) for m in items], record=None, pos=0)
        return 0

    def _field(self, default_width):
        """Next field of a formatted READ: (descriptor letter, text)."""
        c = self.cur
        if c["record"] is None:
            c["record"] = self._next_record()
        if not c["fmt"]:
            raise RuntimeError("more items than edit descriptors")
        kind, w = c["fmt"].pop(0)
        w = w or default_width
        text = c["record"][c["pos"]:c["pos"] + w]
        c["pos"] += w
        return kind, text

    def _next_record(self):
        u = self.cur["u"]
        if u["idx"] >= len(u["lines"]):
            raise RuntimeError("end of file on %s" % u["name"])
        line = u["lines"][u["idx"]]
        u["idx"] += 1
        self.cur["touched"] = True
        return line

    def _token(self):
        while not self.cur["tokens"]:
            self.cur["tokens"] = self._next_record().replace(",", " ").split()
        return self.cur["tokens"].pop(0)

    def transfer_real(self, p, ptr, kind, *_):
        if not self.cur["list"]:
            raise RuntimeError("formatted READ of a real: not something readdata does")
        v = float(self._token().lower().replace("d", "e"))
        (ctypes.c_double if kind == 8 else ctypes.c_float).from_address(ptr).value = v
        return 0

    def transfer_integer(self, p, ptr, kind, *_):
        if not self.cur["list"]:
            d, text = self._field(0)
            if d != "i":
                raise RuntimeError("integer item under an A descriptor")
            (ctypes.c_int64 if kind == 8 else ctypes.c_int32).from_address(ptr).value = int(text.strip() or 0)
            return 0
        t = self._token()
        try:
            v = int(t)
        except ValueError:
            v = int(float(t.lower().replace("d", "e")))
        (ctypes.c_int64 if kind == 8 else ctypes.c_int32).from_address(ptr).value = v
        return 0

    def transfer_character(self, p, ptr, length, *_):
        if self.cur["list"]:
            text = self._token()
        else:
            d, text = self._field(length)
            if d != "a":
                raise RuntimeError("character item under an I descriptor")
        raw = text.encode("latin-1")[:length].ljust(length, b" ")
        ctypes.memmove(ptr, raw, length)
        return 0

    def st_read_done(self, p, *_):
        if not self.cur["touched"]:
            self._next_record()
        self.cur = None
        return 0


class MachoImage:
    """Minimal loader for one MH_BUNDLE with LC_DYLD_INFO_ONLY fixups."""

    def __init__(self, path=REF_SO, verbose=False):
        if os.environ.get(OPT_IN_ENV) != "1":
            raise RuntimeError("running the reference's machine code is opt-in: set %s=1" % OPT_IN_ENV)
        self.f = open(path, "rb").read()
        import hashlib
        digest = hashlib.sha256(self.f).hexdigest()
        if digest != REF_SHA256:
            raise RuntimeError("%s is not the image this loader was written for (sha256 %s)" % (path, digest))
        self.verbose = verbose
        self.segs = []
        self.syms = {}
        self.calls = []          # names of stubbed imports that were actually called
        self._keep = []          # keep ctypes callbacks / buffers alive
        self._parse()
        self._map()
        self._rebase()
        self.io = GfortranIO(self)
        self._bind(self.bind_off, self.bind_size, lazy=False)
        self._bind(self.lazy_off, self.lazy_size, lazy=True)
        self._protect()

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
        base = _libc.mmap(None, self.vmsize, PROT_R | PROT_W, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0)
        if base in (None, ctypes.c_void_p(-1).value):
            raise OSError(ctypes.get_errno(), "mmap failed")
        self.base = base
        for name, vmaddr, vmsize, fileoff, filesize in self.segs:
            if name == "__LINKEDIT" or filesize == 0:
                continue
            ctypes.memmove(base + vmaddr, self.f[fileoff:fileoff + filesize], filesize)

    def _protect(self):
        """W^X: after the fixups __TEXT becomes read+execute; everything else stays read+write, not executable."""
        _libc.mprotect.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        for name, vmaddr, vmsize, fileoff, filesize in self.segs:
            if name == "__TEXT":
                size = (vmsize + 4095) & ~4095
                if _libc.mprotect(self.base + vmaddr, size, PROT_R | PROT_X) != 0:
                    raise OSError(ctypes.get_errno(), "mprotect failed")

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
        if bare in self.io.table:
            return self.io.table[bare]
        if bare in ("__stack_chk_guard", "__stderrp", "_Py_NoneStruct") or bare.startswith("PyExc_") \
                or bare.endswith("_Type"):
            buf = ctypes.create_string_buffer(256)
            self._keep.append(buf)
            return ctypes.addressof(buf)

        def stub(a, b, c, d, _n=bare):
            self.calls.append(_n)
            if _n == "_gfortran_stop_string":
                # Fortran STOP does not return: there is no frame to unwind to from inside the image, so say why
                # and end the process (the live-binary legs are opt-in test infrastructure)
                try:
                    msg = ctypes.string_at(a, min(int(b or 0), 200)).decode("latin-1")
                except Exception:      # noqa: BLE001
                    msg = "?"
                import sys
                sys.stderr.write("reference binary executed STOP '%s'; I/O emulation error: %r\n" % (msg, self.io.error))
                sys.stderr.flush()
                os._exit(70)
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
        "impex": ["outfile", "molfile", "specref"],
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
        self._readdata = ctypes.CFUNCTYPE(None)(self.img.addr("_readdata_"))

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

    def readdata(self, molfile, tkin, density):
        """The reference's own readdata(): parse `molfile`, interpolate the rates to tkin, mix the partners with
        density[id - 1] (pyradex writes cphys.density and tkin, then calls radex.readdata(): core.py:525-570)."""
        raw = os.fsencode(molfile)
        if len(raw) > 120:
            raise ValueError("molfile path longer than the 120 characters of impex.molfile")
        ctypes.memmove(self.a["molfile"], raw.ljust(120, b" "), 120)
        self.set_scalar("tkin", tkin)
        d = self.dview("density", MAXPART)
        d[:] = 0.0
        d[:len(density)] = density
        self.set_int("debug", 0)
        self.img.io.error = None
        del self.img.calls[:]
        self._readdata()
        if self.img.io.error is not None:
            raise RuntimeError("I/O emulation failed inside readdata_: %r" % (self.img.io.error,))
        bad = [c for c in self.img.calls if "stop" in c or "pause" in c]
        if bad:
            raise RuntimeError("readdata_ stopped: %s" % bad)

    def tables(self):
        """What readdata left in /imolec/, /rmolec/, /radi/, /collie/ (0-based copies)."""
        nlev, nline = self.get_int("nlev"), self.get_int("nline")
        cr = self.dview("crate", MAXLEV * MAXLEV).reshape(MAXLEV, MAXLEV)
        return dict(nlev=nlev, nline=nline, npart=self.get_int("npart"),
                    eterm=self.dview("eterm", MAXLEV)[:nlev].copy(), gstat=self.dview("gstat", MAXLEV)[:nlev].copy(),
                    iupp=self.iview("iupp", MAXLINE)[:nline].copy(), ilow=self.iview("ilow", MAXLINE)[:nline].copy(),
                    aeinst=self.dview("aeinst", MAXLINE)[:nline].copy(), eup=self.dview("eup", MAXLINE)[:nline].copy(),
                    xnu=self.dview("xnu", MAXLINE)[:nline].copy(), spfreq=self.dview("spfreq", MAXLINE)[:nline].copy(),
                    amass=self.get_scalar("amass"), totdens=self.get_scalar("totdens"),
                    crate=cr[:nlev, :nlev].T.copy(), ctot=self.dview("ctot", MAXLEV)[:nlev].copy())

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
    """The live-binary legs run only where the reference tree exists, on x86-64, and when asked for."""
    return os.path.exists(REF_SO) and os.uname().machine == "x86_64" and os.environ.get(OPT_IN_ENV) == "1"


if __name__ == "__main__":
    r = RefRadex(verbose=True)
    for m, nm in ((2, "lvg"), (1, "sphere"), (3, "slab")):
        print(nm, [r.escprob(t, m) for t in (1e-3, 0.5, 5.0, 20.0, 200.0, -0.5)])
    print("stub calls:", r.img.calls)
