/*
 * radex_oracle.c -- CPU restatement of the reference's RADEX/pyradex hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under radex_emcee_b200/ may link, import or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Parity status: PINNED for escprob / backrad / matrix (+ the LINPACK LU behind it) against the
 * reference's own compiled Fortran (emcee/pyradex/radex/radex.so, run in this container through
 * oracle/macho_ref.py; fixtures in tests/golden/macho_*.npz, made by oracle/make_golden.py).
 * UNPINNED for readdata (needs libgfortran I/O, cannot be run) and for the Python-level pieces
 * nothing in the reference's tests pins (source_line_surfbrightness, model_lvg, lnlike, lnprior,
 * lnprob) -- there the Python source is the spec and every function below cites it.  The
 * reference's known-answer tests (emcee/pyradex/tests/test_radex.py:99-115,175-200) need the real
 * LAMDA co.dat, which is absent, so they cannot be evaluated here.
 *
 * Citations are relative to /root/reference/.  "radex.so@0x..." = symbol address in
 * emcee/pyradex/radex/radex.so.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define RO_MAXPART 9

/* ---- constants as stored in the binary's constant pool ------------ */
static const double RO_FK = 1.4387809925261357;      /* h c / k, radex.inc values          */
static const double RO_THC = 3.972907393443411e-16;  /* 2 h c                              */
static const double RO_PI = 3.14159265;              /* truncated pi of radex.inc          */
static const double RO_MINPOP = 1e-20;
#define RO_F(x) ((double)(x##f))                     /* single-precision literal promoted  */

typedef struct {
  int nlev, nline, npart;
  double amass;
  double *eterm, *gstat;             /* [nlev]  cm^-1, weights                    */
  int *iupp, *ilow;                  /* [nline] 1-based                           */
  double *aeinst, *spfreq, *eup, *xnu; /* [nline]; xnu = eterm(iupp)-eterm(ilow)  */
  int part_id[RO_MAXPART], ncoll[RO_MAXPART], ntemp[RO_MAXPART];
  double *temp[RO_MAXPART];          /* [ntemp]                                   */
  int *lcu[RO_MAXPART], *lcl[RO_MAXPART]; /* [ncoll] 1-based                      */
  double *coll[RO_MAXPART];          /* [ncoll][ntemp] downward rates cm^3 s^-1   */
} ro_mol;

/* the part of the Fortran COMMON state one solve touches */
typedef struct {
  const ro_mol *mol;
  int method;                        /* 1 sphere, 2 LVG, 3 slab (core.py:690-700) */
  double tkin, tbg, cdmol, deltav, totdens;
  double density[RO_MAXPART];        /* index = LAMDA partner id - 1: H2,pH2,oH2,e,H,He,H+ */
  double *crate, *ctot;              /* [nlev*nlev] crate[i*nlev+j] = rate i->j ; [nlev]   */
  double *xpop;                      /* [nlev]                                            */
  double *tex, *taul, *backi, *totalb, *trj; /* [nline]                                   */
  double *yrate, *rhs, *lu;          /* work (nlev+1)^2, nlev+1, nlev^2                   */
  int *ipvt;
  int nthick, nfat;
  double tsum;
} ro_state;

/* ------------------------------------------------------------------------------------------- */
/* LAMDA reader: restates Fortran readdata's parse (radex.so@0x1cf90).  */
static char *ro_line(FILE *f, char *buf, int n) { return fgets(buf, n, f); }

void ro_mol_free(ro_mol *m) {
  if (!m) return;
  free(m->eterm); free(m->gstat); free(m->iupp); free(m->ilow);
  free(m->aeinst); free(m->spfreq); free(m->eup); free(m->xnu);
  for (int p = 0; p < RO_MAXPART; ++p) { free(m->temp[p]); free(m->lcu[p]); free(m->lcl[p]); free(m->coll[p]); }
  free(m);
}

ro_mol *ro_mol_load(const char *path) {
  FILE *f = fopen(path, "r");
  if (!f) return NULL;
  enum { N = 1 << 16 };
  char *buf = (char *)malloc(N);
  ro_mol *m = (ro_mol *)calloc(1, sizeof(ro_mol));
  int ok = 0;
  do {
    /* every block is preceded by one '!' comment line which the reader skips blindly */
    if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) break;           /* name           */
    if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) break;           /* weight         */
    m->amass = atof(buf);
    if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) break;           /* nlev           */
    m->nlev = atoi(buf);
    if (m->nlev < 2) break;
    m->eterm = (double *)calloc(m->nlev, sizeof(double));
    m->gstat = (double *)calloc(m->nlev, sizeof(double));
    if (!ro_line(f, buf, N)) break;
    int bad = 0;
    for (int i = 0; i < m->nlev; ++i) {
      int idx;
      if (!ro_line(f, buf, N) || sscanf(buf, "%d %lf %lf", &idx, &m->eterm[i], &m->gstat[i]) != 3) { bad = 1; break; }
    }
    if (bad) break;
    if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) break;           /* nline          */
    m->nline = atoi(buf);
    m->iupp = (int *)calloc(m->nline, sizeof(int));
    m->ilow = (int *)calloc(m->nline, sizeof(int));
    m->aeinst = (double *)calloc(m->nline, sizeof(double));
    m->spfreq = (double *)calloc(m->nline, sizeof(double));
    m->eup = (double *)calloc(m->nline, sizeof(double));
    m->xnu = (double *)calloc(m->nline, sizeof(double));
    if (!ro_line(f, buf, N)) break;
    for (int i = 0; i < m->nline; ++i) {
      int idx;
      if (!ro_line(f, buf, N) ||
          sscanf(buf, "%d %d %d %lf %lf %lf", &idx, &m->iupp[i], &m->ilow[i], &m->aeinst[i], &m->spfreq[i], &m->eup[i]) != 6) { bad = 1; break; }
      /* line frequency comes from the level energies, not from the GHz column */
      m->xnu[i] = m->eterm[m->iupp[i] - 1] - m->eterm[m->ilow[i] - 1];
    }
    if (bad) break;
    if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) break;           /* npart          */
    m->npart = atoi(buf);
    if (m->npart < 1 || m->npart > RO_MAXPART) break;
    for (int p = 0; p < m->npart && !bad; ++p) {
      if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) { bad = 1; break; }
      m->part_id[p] = atoi(buf);                                     /* leading integer */
      if (m->part_id[p] < 1 || m->part_id[p] > 7) { bad = 1; break; }
      if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) { bad = 1; break; }
      m->ncoll[p] = atoi(buf);
      if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) { bad = 1; break; }
      m->ntemp[p] = atoi(buf);
      if (!ro_line(f, buf, N) || !ro_line(f, buf, N)) { bad = 1; break; }
      m->temp[p] = (double *)calloc(m->ntemp[p], sizeof(double));
      {
        char *s = buf, *e;
        for (int t = 0; t < m->ntemp[p]; ++t) { m->temp[p][t] = strtod(s, &e); if (e == s) { bad = 1; break; } s = e; }
      }
      if (bad) break;
      if (!ro_line(f, buf, N)) { bad = 1; break; }
      m->lcu[p] = (int *)calloc(m->ncoll[p], sizeof(int));
      m->lcl[p] = (int *)calloc(m->ncoll[p], sizeof(int));
      m->coll[p] = (double *)calloc((size_t)m->ncoll[p] * m->ntemp[p], sizeof(double));
      for (int c = 0; c < m->ncoll[p]; ++c) {
        if (!ro_line(f, buf, N)) { bad = 1; break; }
        char *s = buf, *e;
        (void)strtol(s, &e, 10); s = e;
        m->lcu[p][c] = (int)strtol(s, &e, 10); s = e;
        m->lcl[p][c] = (int)strtol(s, &e, 10); s = e;
        for (int t = 0; t < m->ntemp[p]; ++t) {
          m->coll[p][(size_t)c * m->ntemp[p] + t] = strtod(s, &e);
          if (e == s) { bad = 1; break; }
          s = e;
        }
        if (bad) break;
      }
    }
    if (bad) break;
    ok = 1;
  } while (0);
  fclose(f);
  free(buf);
  if (!ok) { ro_mol_free(m); return NULL; }
  return m;
}

int ro_mol_nlev(const ro_mol *m) { return m->nlev; }
int ro_mol_nline(const ro_mol *m) { return m->nline; }
int ro_mol_npart(const ro_mol *m) { return m->npart; }
void ro_mol_get_partner_ids(const ro_mol *m, int *ids) { for (int p = 0; p < m->npart; ++p) ids[p] = m->part_id[p]; }
void ro_mol_get_levels(const ro_mol *m, double *eterm, double *gstat) {
  memcpy(eterm, m->eterm, sizeof(double) * m->nlev);
  memcpy(gstat, m->gstat, sizeof(double) * m->nlev);
}
void ro_mol_get_lines(const ro_mol *m, int *iupp, int *ilow, double *aeinst, double *spfreq, double *eup, double *xnu) {
  memcpy(iupp, m->iupp, sizeof(int) * m->nline);
  memcpy(ilow, m->ilow, sizeof(int) * m->nline);
  memcpy(aeinst, m->aeinst, sizeof(double) * m->nline);
  memcpy(spfreq, m->spfreq, sizeof(double) * m->nline);
  memcpy(eup, m->eup, sizeof(double) * m->nline);
  memcpy(xnu, m->xnu, sizeof(double) * m->nline);
}

/* ------------------------------------------------------------------------------------------- */
ro_state *ro_state_new(const ro_mol *mol) {
  ro_state *s = (ro_state *)calloc(1, sizeof(ro_state));
  int nl = mol->nlev, nn = mol->nline, np = nl + 1;
  s->mol = mol;
  s->method = 2;
  s->deltav = 1e5;                   /* 1 km/s in cm/s (core.py:447-454) */
  s->crate = (double *)calloc((size_t)nl * nl, sizeof(double));
  s->ctot = (double *)calloc(nl, sizeof(double));
  s->xpop = (double *)calloc(nl, sizeof(double));
  s->tex = (double *)calloc(nn, sizeof(double));
  s->taul = (double *)calloc(nn, sizeof(double));
  s->backi = (double *)calloc(nn, sizeof(double));
  s->totalb = (double *)calloc(nn, sizeof(double));
  s->trj = (double *)calloc(nn, sizeof(double));
  s->yrate = (double *)calloc((size_t)np * np, sizeof(double));
  s->rhs = (double *)calloc(np, sizeof(double));
  s->lu = (double *)calloc((size_t)nl * nl, sizeof(double));
  s->ipvt = (int *)calloc(np, sizeof(int));
  return s;
}

void ro_state_free(ro_state *s) {
  if (!s) return;
  free(s->crate); free(s->ctot); free(s->xpop); free(s->tex); free(s->taul);
  free(s->backi); free(s->totalb); free(s->trj); free(s->yrate); free(s->rhs); free(s->lu); free(s->ipvt);
  free(s);
}

/* accessors for ctypes */
double *ro_state_xpop(ro_state *s) { return s->xpop; }
double *ro_state_tex(ro_state *s) { return s->tex; }
double *ro_state_taul(ro_state *s) { return s->taul; }
double *ro_state_backi(ro_state *s) { return s->backi; }
double *ro_state_totalb(ro_state *s) { return s->totalb; }
double *ro_state_crate(ro_state *s) { return s->crate; }
double *ro_state_ctot(ro_state *s) { return s->ctot; }
double ro_state_totdens(const ro_state *s) { return s->totdens; }
int ro_state_nthick(const ro_state *s) { return s->nthick; }
void ro_state_set_method(ro_state *s, int method) { s->method = method; }
void ro_state_set_column(ro_state *s, double cdmol, double deltav_cms) { s->cdmol = cdmol; s->deltav = deltav_cms; }

/* ------------------------------------------------------------------------------------------- */
/* readdata's numerical part: T-interpolation, partner mix, detailed balance, ctot.
 * (radex.so@0x1cf90; SURVEY.md 3.3 "readdata()").  density[] is indexed by LAMDA id - 1, the
 * layout pyradex writes into cphys.density (core.py:525-561); totdens = sum (core.py:565).   */
void ro_set_physics(ro_state *s, double tkin, const double *density /*[7..9]*/, int ndens) {
  const ro_mol *m = s->mol;
  int nl = m->nlev;
  s->tkin = tkin;
  memset(s->density, 0, sizeof(s->density));
  for (int i = 0; i < ndens && i < RO_MAXPART; ++i) s->density[i] = density[i];
  s->totdens = 0.0;
  for (int i = 0; i < RO_MAXPART; ++i) s->totdens += s->density[i];
  for (int i = 0; i < nl * nl; ++i) s->crate[i] = 0.0;
  for (int p = 0; p < m->npart; ++p) {
    const double *T = m->temp[p];
    int nt = m->ntemp[p];
    double dens = s->density[m->part_id[p] - 1];
    for (int c = 0; c < m->ncoll[p]; ++c) {
      const double *r = m->coll[p] + (size_t)c * nt;
      double v;
      if (tkin <= T[0]) {
        v = r[0];                                  /* clamped, no extrapolation */
      } else if (tkin >= T[nt - 1]) {
        v = r[nt - 1];
      } else {
        v = r[nt - 1];
        for (int t = 0; t < nt - 1; ++t) {
          if (tkin > T[t] && tkin <= T[t + 1]) {
            double fint = (tkin - T[t]) / (T[t + 1] - T[t]);
            v = r[t] + fint * (r[t + 1] - r[t]);
            if (v < 0.0) v = r[t];
            break;
          }
        }
      }
      int iu = m->lcu[p][c] - 1, il = m->lcl[p][c] - 1;
      s->crate[iu * nl + il] += dens * v;
    }
  }
  /* upward rates from detailed balance */
  for (int iu = 0; iu < nl; ++iu)
    for (int il = 0; il < nl; ++il) {
      double ediff = m->eterm[iu] - m->eterm[il];
      if (ediff > 0.0) {
        double x = RO_FK * ediff / tkin;
        if (x >= 160.0) s->crate[il * nl + iu] = 0.0;
        else s->crate[il * nl + iu] = m->gstat[iu] / m->gstat[il] * exp(-x) * s->crate[iu * nl + il];
      }
    }
  for (int i = 0; i < nl; ++i) {
    double t = 0.0;
    for (int j = 0; j < nl; ++j) t += s->crate[i * nl + j];
    s->ctot[i] = t;
  }
}

/* backrad, tbg > 0 branch (radex.so@0x1be30; SURVEY.md 3.3 "backrad()"). */
void ro_backrad(ro_state *s, double tbg) {
  const ro_mol *m = s->mol;
  s->tbg = tbg;
  for (int l = 0; l < m->nline; ++l) {
    double hnu = RO_FK * m->xnu[l] / tbg;
    double v;
    if (hnu >= 160.0) v = 1.0e-30;   /* eps */
    else v = RO_THC * pow(m->xnu[l], 3.0) / (exp(hnu) - 1.0);   /* xnu**3. compiles to pow() */
    s->backi[l] = v;
    s->totalb[l] = v;
    s->trj[l] = tbg;
  }
}

/* escprob(tau) (radex.so@0xa9c0).  Constants in binary order. */
double ro_escprob(double tau, int method) {
  double taur = tau / 2.0, beta;
  if (method == 1) {          /* uniform sphere */
    if (fabs(taur) < RO_F(0.1)) {
      beta = 1.0 - 0.75 * taur + (taur * taur) / 2.5 - (taur * taur * taur) / 6.0 + (taur * taur * taur * taur) / 17.5;
    } else if (fabs(taur) > 50.0) {
      beta = 0.75 / taur;
    } else {
      beta = 0.75 / taur * (1.0 - 1.0 / (2.0 * (taur * taur)) + (1.0 / taur + 1.0 / (2.0 * (taur * taur))) * exp(-2.0 * taur));
    }
  } else if (method == 2) {   /* expanding sphere = LVG */
    if (fabs(taur) < RO_F(0.01)) {
      beta = 1.0;
    } else if (fabs(taur) < 7.0) {
      beta = 2.0 * (1.0 - exp(-RO_F(2.34) * taur)) / (RO_F(4.68) * taur);
    } else {
      beta = 2.0 / (taur * 4.0 * sqrt(log(taur / sqrt(RO_PI))));
    }
  } else {                    /* slab */
    if (fabs(3.0 * tau) < RO_F(0.1)) {
      beta = 1.0 - 1.5 * (tau + tau * tau);
    } else if (fabs(3.0 * tau) > 50.0) {
      beta = 1.0 / (3.0 * tau);
    } else {
      beta = (1.0 - exp(-3.0 * tau)) / (3.0 * tau);
    }
  }
  return beta;
}

/* LINPACK-style LU with partial pivoting, column-major a[i + j*lda] (sgefa/sgesl behind
 * lubksb -> sgeir, radex.so@0x17cb0,0x16d50,0xf3d0,0xdb70; SGEIR's residual pass only
 * estimates accuracy and does not touch the solution).                       */
static int ro_gefa(double *a, int lda, int n, int *ipvt) {
  int info = 0;
  for (int k = 0; k < n - 1; ++k) {
    int l = k;
    double amax = fabs(a[k + k * lda]);
    for (int i = k + 1; i < n; ++i) {
      double v = fabs(a[i + k * lda]);
      if (v > amax) { amax = v; l = i; }
    }
    ipvt[k] = l;
    if (a[l + k * lda] == 0.0) { info = k + 1; continue; }
    if (l != k) { double t = a[l + k * lda]; a[l + k * lda] = a[k + k * lda]; a[k + k * lda] = t; }
    double t = -1.0 / a[k + k * lda];
    for (int i = k + 1; i < n; ++i) a[i + k * lda] *= t;
    for (int j = k + 1; j < n; ++j) {
      double tj = a[l + j * lda];
      if (l != k) { a[l + j * lda] = a[k + j * lda]; a[k + j * lda] = tj; }
      for (int i = k + 1; i < n; ++i) a[i + j * lda] += tj * a[i + k * lda];
    }
  }
  ipvt[n - 1] = n - 1;
  if (a[(n - 1) + (n - 1) * lda] == 0.0) info = n;
  return info;
}

static void ro_gesl(const double *a, int lda, int n, const int *ipvt, double *b) {
  for (int k = 0; k < n - 1; ++k) {
    int l = ipvt[k];
    double t = b[l];
    if (l != k) { b[l] = b[k]; b[k] = t; }
    for (int i = k + 1; i < n; ++i) b[i] += t * a[i + k * lda];
  }
  for (int k = n - 1; k >= 0; --k) {
    b[k] /= a[k + k * lda];
    double t = -b[k];
    for (int i = 0; i < k; ++i) b[i] += t * a[i + k * lda];
  }
}

/* One call of Fortran matrix(niter, conv) (radex.so@0x17f70).
 * Returns the conv flag the Fortran computes (and f2py drops).                               */
int ro_matrix(ro_state *s, int niter) {
  const ro_mol *m = s->mol;
  const int nl = m->nlev, nn = m->nline, np = nl + 1;
  double *y = s->yrate;              /* column-major y[i + j*np] == yrate(i+1, j+1) */
  double *rhs = s->rhs;
  const double eps_td = 1.0e-30 * s->totdens;   /* 1.0d-30 literal: verified bitwise against the binary yrate */
  int conv = 0;
#define Y(i, j) y[(i) + (size_t)(j) * np]
  for (int i = 0; i < nl; ++i) {
    for (int j = 0; j < nl; ++j) Y(i, j) = -eps_td;
    Y(np - 1, i) = 1.0;
    rhs[i] = eps_td;
    Y(i, np - 1) = eps_td;
  }
  rhs[np - 1] = eps_td;
  Y(np - 1, np - 1) = 0.0;
  double cddv = 0.0;
  if (niter == 0) {
    for (int l = 0; l < nn; ++l) {
      int mu = m->iupp[l] - 1, n = m->ilow[l] - 1;
      double etr = RO_FK * m->xnu[l] / s->trj[l];
      double exr = (etr >= 160.0) ? 0.0 : 1.0 / (exp(etr) - 1.0);
      double a = m->aeinst[l], gm = m->gstat[mu], gn = m->gstat[n];
      Y(mu, mu) += a * (1.0 + exr);
      Y(n, n) += a * gm * exr / gn;      /* niter=0 branch has no parentheses: ((a*gm)*exr)/gn */
      Y(mu, n) -= a * (gm / gn) * exr;
      Y(n, mu) -= a * (1.0 + exr);
    }
  } else {
    cddv = s->cdmol / s->deltav;
    s->nthick = 0;
    s->nfat = 0;
    for (int l = 0; l < nn; ++l) {
      int mu = m->iupp[l] - 1, n = m->ilow[l] - 1;
      double xt = pow(m->xnu[l], 3.0);
      double a = m->aeinst[l], gm = m->gstat[mu], gn = m->gstat[n];
      s->taul[l] = cddv * (s->xpop[n] * gm / gn - s->xpop[mu]) / (/*fgaus*/ RO_F(1.0645) * 8.0 * RO_PI * xt / a);
      if (s->taul[l] > 1.0e-2) s->nthick++;
      if (s->taul[l] > 1.0e5) s->nfat++;
      double beta = ro_escprob(s->taul[l], s->method);
      double bnu = s->totalb[l] * beta;
      double exr = bnu / (RO_THC * xt);
      Y(mu, mu) += a * (beta + exr);
      Y(n, n) += a * (gm * exr / gn);
      Y(mu, n) -= a * (gm / gn) * exr;
      Y(n, mu) -= a * (beta + exr);
    }
  }
  for (int i = 0; i < nl; ++i) {
    Y(i, i) += s->ctot[i];
    for (int j = 0; j < nl; ++j)
      if (i != j) Y(i, j) -= s->crate[j * nl + i];
  }
  /* lubksb as compiled into this build (radex.so@0x17cb0, disassembled): it does NOT solve the
   * (nlev+1)^2 system.  It builds a reduced nlev x nlev system from rows 1..nlev-1 of columns
   * 1..nlev of yrate, overwrites row nlev with 1.0 (conservation), uses rhs = (0,...,0,1) and
   * hands that to sgeir (LU with partial pivoting; sgeir's residual pass only estimates
   * accuracy).  The caller's rhs(1..nlev) receive the solution; rhs(nplus) is left alone.     */
  {
    double *w = s->lu;             /* nl x nl, column-major */
    for (int j = 0; j < nl; ++j) {
      for (int i = 0; i < nl - 1; ++i) w[i + (size_t)j * nl] = Y(i, j);
      w[(nl - 1) + (size_t)j * nl] = 1.0;
    }
    for (int i = 0; i < nl - 1; ++i) rhs[i] = 0.0;
    rhs[nl - 1] = 1.0;
    ro_gefa(w, nl, nl, s->ipvt);
    ro_gesl(w, nl, nl, s->ipvt, rhs);
  }
#undef Y
  double total = 0.0;
  for (int i = 0; i < nl; ++i) total += rhs[i];
  double xpopold[4096];
  for (int i = 0; i < nl; ++i) {
    xpopold[i] = fmax(RO_MINPOP, s->xpop[i]);
    s->xpop[i] = fmax(RO_MINPOP, rhs[i] / total);
    if (niter == 0) xpopold[i] = s->xpop[i];
  }
  double tsum = 0.0;
  for (int l = 0; l < nn; ++l) {
    int mu = m->iupp[l] - 1, n = m->ilow[l] - 1;
    double xt = pow(m->xnu[l], 3.0);
    double gm = m->gstat[mu], gn = m->gstat[n];
    if (niter == 0) {
      if (s->xpop[n] <= RO_MINPOP || s->xpop[mu] <= RO_MINPOP) s->tex[l] = s->totalb[l];
      else s->tex[l] = RO_FK * m->xnu[l] / log(s->xpop[n] * gm / (s->xpop[mu] * gn));
    } else {
      double thistex;
      if (s->xpop[n] <= RO_MINPOP || s->xpop[mu] <= RO_MINPOP) thistex = s->tex[l];
      else thistex = RO_FK * m->xnu[l] / log(s->xpop[n] * gm / (s->xpop[mu] * gn));
      if (s->taul[l] > RO_F(0.01)) tsum += fabs((thistex - s->tex[l]) / thistex);
      s->tex[l] = 0.5 * (thistex + s->tex[l]);
      s->taul[l] = cddv * (s->xpop[n] * gm / gn - s->xpop[mu]) / (RO_F(1.0645) * 8.0 * RO_PI * xt / m->aeinst[l]);
    }
  }
  s->tsum = tsum;
  if (niter >= 10) {               /* miniter of radex.inc */
    if (s->nthick == 0) conv = 1;
    else if (tsum / s->nthick < RO_F(1.0e-6)) conv = 1;
  }
  for (int i = 0; i < nl; ++i) s->xpop[i] = RO_F(0.3) * s->xpop[i] + RO_F(0.7) * xpopold[i];
  return conv;
}

/* stop rules */
enum { RO_STOP_PYRADEX = 0, RO_STOP_RADEX = 1 };

/* Radex.run_radex's loop (emcee/pyradex/core.py:896-925).  reuse_last -> first call has niter=1
 * and continues from whatever state holds.  stop_rule RO_STOP_PYRADEX = the loop as written
 * (sum|dx| < abs_tol and iter > miniter; the relative test is NaN-dead);
 * RO_STOP_RADEX = Fortran's own conv flag.  Returns _iter_counter.                             */
/* numpy's pairwise summation (DOUBLE_pairwise_sum), which is what `level_diff.sum()` runs over the
 * 2999-long, zero-padded level_population array (core.py:911-914).                              */
static double ro_np_pairwise(const double *a, long n) {
  if (n < 8) {
    double res = 0.0;
    for (long i = 0; i < n; ++i) res += a[i];
    return res;
  } else if (n <= 128) {
    double r[8];
    long i;
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  } else {
    long n2 = n / 2;
    n2 -= n2 % 8;
    return ro_np_pairwise(a, n2) + ro_np_pairwise(a + n2, n - n2);
  }
}

#define RO_PYRADEX_MAXLEV 2999   /* length of collie.xpop in this build */

int ro_run(ro_state *s, int reuse_last, int stop_rule, int miniter, int maxiter, double abs_tol) {
  int nl = s->mol->nlev;
  int it = reuse_last ? 1 : 0;
  static __thread double last[RO_PYRADEX_MAXLEV], diff[RO_PYRADEX_MAXLEV];
  memset(diff, 0, sizeof(diff));
  memcpy(last, s->xpop, sizeof(double) * nl);
  for (;;) {
    if (it >= maxiter) break;
    int conv = ro_matrix(s, it);
    for (int i = 0; i < nl; ++i) diff[i] = fabs(last[i] - s->xpop[i]);
    double d = ro_np_pairwise(diff, RO_PYRADEX_MAXLEV);
    if (stop_rule == RO_STOP_RADEX) {
      if (conv) break;
    } else if (d < abs_tol && it > miniter) {
      break;
    }
    memcpy(last, s->xpop, sizeof(double) * nl);
    ++it;
  }
  return it;
}

/* source_line_surfbrightness = source_brightness - background_brightness
 * (emcee/pyradex/base_class.py:275-277; core.py:986-1003): no tau/earg guards, and fk/thc are
 * the *astropy* constants of the reference's environment, passed in by the caller.            */
void ro_surface_brightness(const ro_state *s, double fk_epi, double thc_epi, double *out) {
  const ro_mol *m = s->mol;
  for (int l = 0; l < m->nline; ++l) {
    double ftau = exp(-s->taul[l]);
    double xt = m->xnu[l] * m->xnu[l] * m->xnu[l];
    double earg = fk_epi * m->xnu[l] / s->tex[l];
    double bnutex = thc_epi * xt / (exp(earg) - 1.0);
    double toti = s->backi[l] * ftau + bnutex * (1.0 - ftau);
    out[l] = toti - s->backi[l];
  }
}

/* ------------------------------------------------------------------------------------------- */
/* Full forward model for one parameter set, history-free (clean niter=0 start).
 * Mirrors set_params(density={'oH2','pH2'}, column, temperature) + run_radex + epilogue
 * (emcee/emcee_radex.py:120-128).  status bits: 1 T range, 2 N range (the ValueErrors of
 * core.py:734-735,771-772), 4 hit maxiter, 8 non-finite output.                               */
int ro_solve_dens(ro_state *s, double tkin, const double *dens7, double cdmol, double deltav_kms,
                  double tbg, int method, int stop_rule, int miniter, int maxiter, double abs_tol,
                  double fk_epi, double thc_epi, double *surf, int *niter_out) {
  int status = 0;
  if (!(tkin > 0.0 && tkin <= 1e4)) status |= 1;
  if (!(cdmol >= 1e5 && cdmol <= 1e25)) status |= 2;
  if (status) { if (niter_out) *niter_out = 0; return status; }
  ro_set_physics(s, tkin, dens7, 7);
  s->cdmol = cdmol;
  s->deltav = deltav_kms * 1e5;
  s->method = method;
  if (s->tbg != tbg) ro_backrad(s, tbg);
  int it = ro_run(s, 0, stop_rule, miniter, maxiter, abs_tol);
  if (niter_out) *niter_out = it;
  if (it >= maxiter) status |= 4;
  if (surf) {
    ro_surface_brightness(s, fk_epi, thc_epi, surf);
    for (int l = 0; l < s->mol->nline; ++l) if (!isfinite(surf[l])) status |= 8;
  }
  return status;
}

/* density={'oH2': .., 'pH2': ..} as the drivers set it (emcee/emcee_radex.py:122-124): pyradex's density setter
 * folds the two into n(H2) when the molecular file lists H2 itself as a partner, and keeps them apart otherwise
 * (emcee/pyradex/core.py:551-556).                                                                            */
int ro_solve(ro_state *s, double tkin, double n_ph2, double n_oh2, double cdmol, double deltav_kms,
             double tbg, int method, int stop_rule, int miniter, int maxiter, double abs_tol,
             double fk_epi, double thc_epi, double *surf, int *niter_out) {
  double dens[7] = {0.0, n_ph2, n_oh2, 0, 0, 0, 0};
  for (int p = 0; p < s->mol->npart; ++p)
    if (s->mol->part_id[p] == 1) {
      dens[0] = n_ph2 + n_oh2;
      dens[1] = dens[2] = 0.0;
    }
  return ro_solve_dens(s, tkin, dens, cdmol, deltav_kms, tbg, method, stop_rule, miniter, maxiter, abs_tol, fk_epi,
                       thc_epi, surf, niter_out);
}

/* lnlike (emcee/emcee_radex.py:132-167; emcee_radex_2comp.py:169-196): model[] already in Jy km/s. */
double ro_lnlike(const double *model, const double *flux, const double *eflux, int nobs) {
  double chi2 = 0.0, logterm = 0.0;
  const double max_safe = sqrt(DBL_MAX) / 10.0;
  for (int i = 0; i < nobs; ++i)
    if (!isfinite(flux[i]) || !isfinite(model[i])) return -INFINITY;
  for (int i = 0; i < nobs; ++i) {
    double e = fmax(fabs(eflux[i]), 1e-12);
    if (!isfinite(e)) return -INFINITY;
    double r = (flux[i] - model[i]) / e;
    if (!isfinite(r) || fabs(r) > max_safe) return -INFINITY;
    chi2 += r * r;
    logterm += log(e);
  }
  return -0.5 * (chi2 + 2.0 * logterm);
}

/* lnprior, one component (emcee/emcee_radex.py:169-175). bounds[4][2]. */
double ro_lnprior1(const double *p, const double *bounds) {
  for (int i = 0; i < 4; ++i)
    if (p[i] > bounds[2 * i + 1] || p[i] < bounds[2 * i]) return -INFINITY;
  if ((p[2] - p[0] >= 17.5) || (p[2] - p[0] <= 10.0)) return -INFINITY;
  return 0.0;
}

/* lnprior, two components (emcee/emcee_radex_2comp.py:199-234). bounds[8][2]; has_td=0 -> T_d None. */
double ro_lnprior2(const double *p, const double *bounds, int has_td, double t_d) {
  for (int i = 0; i < 8; ++i)
    if (p[i] > bounds[2 * i + 1] || p[i] < bounds[2 * i]) return -INFINITY;
  if (p[5] <= p[1]) return -INFINITY;
  if ((p[2] - p[0]) >= 18.0 || (p[2] - p[0]) <= 9.0 || (p[6] - p[4]) >= 18.0 || (p[6] - p[4]) <= 9.0) return -INFINITY;
  if (p[3] < p[7]) return -INFINITY;
  double logp = 0.0;
  for (int i = 0; i < 8; ++i) {
    if (i == 1 && has_td) {
      double tk = pow(10.0, p[i]);
      if (t_d <= 0) return -INFINITY;
      double sigma = 1.0 * t_d;
      double z = (tk - t_d) / sigma;
      logp += (-0.5 * (z * z) - log(sigma * sqrt(2.0 * M_PI)));
    } else {
      logp += -(bounds[2 * i + 1] - bounds[2 * i]);
    }
  }
  return logp;
}

/* model_lvg + lnprob, one component (emcee/emcee_radex.py:120-130,177-181).
 * p = (log n, log T, log N/dv, log size); flux = surf[Jup-1] * 10^size sr * 1 km/s -> Jy km/s,
 * i.e. x 1e23.  fortho = opr/(1+opr) with opr=3 (emcee/emcee_radex.py:95-96).                 */
double ro_lnprob1(ro_state *s, const double *p, const int *jup, const double *flux, const double *eflux,
                  int nobs, const double *bounds, double tbg, int stop_rule, int miniter, int maxiter,
                  double abs_tol, double fk_epi, double thc_epi) {
  double lp = ro_lnprior1(p, bounds);
  if (!isfinite(lp)) return -INFINITY;
  const double fortho = 3.0 / (1.0 + 3.0);
  double surf[4096], model[64];
  double dens = pow(10.0, p[0]);
  int st = ro_solve(s, pow(10.0, p[1]), (1 - fortho) * dens, fortho * dens, pow(10.0, p[2]), 1.0, tbg, 2,
                    stop_rule, miniter, maxiter, abs_tol, fk_epi, thc_epi, surf, NULL);
  if (st & 3) return -INFINITY;      /* ValueError -> -inf (emcee_radex.py:134-137) */
  for (int i = 0; i < nobs; ++i) model[i] = surf[jup[i] - 1] * pow(10.0, p[3]) * 1e23;
  return lp + ro_lnlike(model, flux, eflux, nobs);
}

/* two components (emcee/emcee_radex_2comp.py:122-147,237-244) */
double ro_lnprob2(ro_state *s, const double *p, const int *jup, const double *flux, const double *eflux,
                  int nobs, const double *bounds, int has_td, double t_d, double tbg, int stop_rule,
                  int miniter, int maxiter, double abs_tol, double fk_epi, double thc_epi) {
  double lp = ro_lnprior2(p, bounds, has_td, t_d);
  if (!isfinite(lp)) return -INFINITY;
  const double fortho = 3.0 / (1.0 + 3.0);
  double surf1[4096], surf2[4096], model[64];
  double d1 = pow(10.0, p[0]), d2 = pow(10.0, p[4]);
  int st = ro_solve(s, pow(10.0, p[1]), (1 - fortho) * d1, fortho * d1, pow(10.0, p[2]), 1.0, tbg, 2,
                    stop_rule, miniter, maxiter, abs_tol, fk_epi, thc_epi, surf1, NULL);
  if (st & 3) return -INFINITY;
  st = ro_solve(s, pow(10.0, p[5]), (1 - fortho) * d2, fortho * d2, pow(10.0, p[6]), 1.0, tbg, 2,
                stop_rule, miniter, maxiter, abs_tol, fk_epi, thc_epi, surf2, NULL);
  if (st & 3) return -INFINITY;
  for (int i = 0; i < nobs; ++i)
    model[i] = surf1[jup[i] - 1] * pow(10.0, p[3]) * 1e23 + surf2[jup[i] - 1] * pow(10.0, p[7]) * 1e23;
  double ll = ro_lnlike(model, flux, eflux, nobs);
  if (!isfinite(ll)) return -INFINITY;
  return lp + ll;
}

/* Batched helpers (used as the timed CPU baseline and by parity tests). One state per call,
 * so callers may run several threads with one state each.                                     */
void ro_solve_batch(ro_state *s, long n, const double *tkin, const double *n_ph2, const double *n_oh2,
                    const double *cdmol, double deltav_kms, double tbg, int method, int stop_rule,
                    int miniter, int maxiter, double abs_tol, double fk_epi, double thc_epi,
                    double *xpop, double *tex, double *tau, double *surf, int *niter, int *status) {
  int nl = s->mol->nlev, nn = s->mol->nline;
  double sb[4096];
  for (long i = 0; i < n; ++i) {
    int it = 0;
    int st = ro_solve(s, tkin[i], n_ph2[i], n_oh2[i], cdmol[i], deltav_kms, tbg, method, stop_rule, miniter,
                      maxiter, abs_tol, fk_epi, thc_epi, sb, &it);
    if (status) status[i] = st;
    if (niter) niter[i] = it;
    if (st & 3) {
      if (xpop) for (int k = 0; k < nl; ++k) xpop[i * nl + k] = NAN;
      if (tex) for (int k = 0; k < nn; ++k) tex[i * nn + k] = NAN;
      if (tau) for (int k = 0; k < nn; ++k) tau[i * nn + k] = NAN;
      if (surf) for (int k = 0; k < nn; ++k) surf[i * nn + k] = NAN;
      continue;
    }
    if (xpop) memcpy(xpop + i * nl, s->xpop, sizeof(double) * nl);
    if (tex) memcpy(tex + i * nn, s->tex, sizeof(double) * nn);
    if (tau) memcpy(tau + i * nn, s->taul, sizeof(double) * nn);
    if (surf) memcpy(surf + i * nn, sb, sizeof(double) * nn);
  }
}

/* the same with one density per LAMDA partner id (dens[i * 7 + id - 1]): any mix of H2, p-H2, o-H2, e, H, He, H+ */
void ro_solve_batch_dens(ro_state *s, long n, const double *tkin, const double *dens7, const double *cdmol,
                         double deltav_kms, double tbg, int method, int stop_rule, int miniter, int maxiter,
                         double abs_tol, double fk_epi, double thc_epi, double *xpop, double *tex, double *tau,
                         double *surf, int *niter, int *status) {
  int nl = s->mol->nlev, nn = s->mol->nline;
  double sb[4096];
  for (long i = 0; i < n; ++i) {
    int it = 0;
    int st = ro_solve_dens(s, tkin[i], dens7 + 7 * i, cdmol[i], deltav_kms, tbg, method, stop_rule, miniter, maxiter,
                           abs_tol, fk_epi, thc_epi, sb, &it);
    if (status) status[i] = st;
    if (niter) niter[i] = it;
    for (int k = 0; k < nl; ++k) if (xpop) xpop[i * nl + k] = (st & 3) ? NAN : s->xpop[k];
    for (int k = 0; k < nn; ++k) {
      if (tex) tex[i * nn + k] = (st & 3) ? NAN : s->tex[k];
      if (tau) tau[i * nn + k] = (st & 3) ? NAN : s->taul[k];
      if (surf) surf[i * nn + k] = (st & 3) ? NAN : sb[k];
    }
  }
}

/* debugging/validation accessor: the assembled (nlev+1)^2 rate matrix of the last ro_matrix call
 * is destroyed only in the reduced copy, so yrate can be compared with the binary's yrate.     */
double *ro_state_yrate(ro_state *s) { return s->yrate; }
