/*
 * radex_b200.h -- C ABI of libradex_b200: the B200-native replacement for the hot path of
 * yangcht/radex_emcee (per-walker RADEX escape-probability solve + SLED likelihood + stretch move).
 *
 * Every entry point is extern "C", takes plain pointers and sizes, returns 0 on success and a
 * negative rb_status code on failure (rb_last_error() holds the message).  There is no CPU
 * fallback: every compute entry point fails with RB_ERR_CUDA if no sm_100 device is usable.
 *
 * "Replaces" citations are relative to the reference tree (/root/reference):
 *   core.py      = emcee/pyradex/core.py
 *   radex.so@X   = symbol at offset X of emcee/pyradex/radex/radex.so (the f2py-wrapped Fortran)
 *   er1 / er2    = emcee/emcee_radex.py / emcee/emcee_radex_2comp.py
 */
#ifndef RADEX_B200_H
#define RADEX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rb_mol rb_mol; /* parsed LAMDA molecular table (host memory)          */
typedef struct rb_ctx rb_ctx; /* one GPU: device-resident SoA tables, stream, scratch */

enum rb_status_code {
  RB_OK = 0,
  RB_ERR_ARG = -1,   /* bad argument                                   */
  RB_ERR_IO = -2,    /* cannot open / parse the LAMDA file             */
  RB_ERR_CUDA = -3,  /* CUDA runtime error or no usable device         */
  RB_ERR_LIMIT = -4  /* molecule larger than the kernels support       */
};

/* per-model status bitmask written by the solve kernels */
enum rb_model_status {
  RB_ST_T_RANGE = 1,   /* T not in (0, 1e4]      -> reference raises ValueError, core.py:734-735 */
  RB_ST_N_RANGE = 2,   /* N not in [1e5, 1e25]   -> reference raises ValueError, core.py:771-772 */
  RB_ST_MAXITER = 4,   /* loop ended at maxiter (reference: silent, core.py:904-907)             */
  RB_ST_NONFINITE = 8  /* a returned brightness is NaN/inf                                       */
};

enum rb_stop_rule {
  RB_STOP_PYRADEX = 0, /* core.py:911-917: sum|dxpop| < abs_tol and iter > miniter (default)     */
  RB_STOP_RADEX = 1    /* Fortran matrix()'s own conv flag (radex.so@0x17f70), dropped by f2py   */
};

enum rb_geometry { RB_GEOM_SPHERE = 1, RB_GEOM_LVG = 2, RB_GEOM_SLAB = 3 }; /* core.py:690-700 */

typedef struct rb_opts {
  int32_t stop_rule;   /* rb_stop_rule                                                    */
  int32_t miniter;     /* 10  (core.py:460-461)                                           */
  int32_t maxiter;     /* 200 (core.py:462-463)                                           */
  int32_t kernel;      /* 0 = default (fastest validated: v2, launches ordered by lead-block size,
                              half-warp engine for small lead blocks),
                          1 = v1 shared-memory pivoted LU,
                          2 = v2 without frozen-top caching (cross-check / A-B),
                          3 = v2 with caching but as a single launch (no ordering),
                          4 = v2 with ordered launches but without the half-warp engine (A-B)  */
  double abs_tol;      /* 1e-16 (core.py:857)                                             */
  double fk_epi;       /* h c / k_B used by the brightness epilogue (astropy, core.py:981-984) */
  double thc_epi;      /* 2 h c     used by the brightness epilogue                        */
} rb_opts;

/* observed SLED of one source (er1:229-240 get_source): up to RB_MAX_OBS CO lines */
#define RB_MAX_OBS 16
typedef struct rb_obs {
  int32_t nobs;
  int32_t jup[RB_MAX_OBS];    /* 1-based upper J == 1-based line index (er1:129)   */
  double flux[RB_MAX_OBS];    /* Jy km/s                                            */
  double eflux[RB_MAX_OBS];   /* Jy km/s                                            */
} rb_obs;

/* ---- library ---------------------------------------------------------------------------- */
const char *rb_last_error(void);
void rb_default_opts(rb_opts *opts);
int rb_device_count(void);

/* ---- LAMDA loader: replaces Fortran readdata's parse, radex.so@0x1cf90 (called from
 * core.py:570,744,887) and the collider discovery of emcee/pyradex/utils.py:53-62 ---------- */
int rb_moldata_load(const char *path, rb_mol **out);
void rb_moldata_free(rb_mol *mol);
int rb_moldata_dims(const rb_mol *mol, int32_t *nlev, int32_t *nline, int32_t *npart);
/* partner_id[npart]: LAMDA ids 1 H2, 2 p-H2, 3 o-H2, 4 e, 5 H, 6 He, 7 H+ */
int rb_moldata_partners(const rb_mol *mol, int32_t *partner_id, int32_t *ncoll, int32_t *ntemp);
int rb_moldata_levels(const rb_mol *mol, double *eterm, double *gstat);
int rb_moldata_lines(const rb_mol *mol, int32_t *iupp, int32_t *ilow, double *aeinst, double *spfreq_ghz,
                     double *eup_k, double *xnu);

/* ---- context: replaces init_radex()/Radex.__init__ (er1:104-117, core.py:209-378) ---------- */
int rb_ctx_create(int device, const rb_mol *mol, rb_ctx **out);
void rb_ctx_destroy(rb_ctx *ctx);
int rb_ctx_sync(rb_ctx *ctx);
/* launch on a caller-owned CUDA stream (cudaStream_t as void*; 0 = CUDA's legacy default stream,
 * which is what PyTorch's default stream is); rb_ctx_reset_stream goes back to the ctx's own. */
int rb_ctx_set_stream(rb_ctx *ctx, void *stream);
int rb_ctx_reset_stream(rb_ctx *ctx);

/* ---- batched solve: replaces set_params + run_radex + tex/tau/level_population +
 * source_line_surfbrightness (core.py:388-438, 856-925, 703-717, 986-1003; base_class.py:275-277)
 * and, inside, readdata's interpolation, backrad, matrix, escprob, lubksb/sgeir
 * (radex.so@0x1cf90, 0x1be30, 0x17f70, 0xa9c0, 0x17cb0).
 *   dens : n x npart, partner order of the file, cm^-3      tkin : K      cdmol : cm^-2
 *   outputs (any may be NULL): xpop n x nlev, tex/tau/surf n x nline
 *     surf = source_line_surfbrightness, erg s^-1 cm^-2 Hz^-1 sr^-1
 * Host-pointer form copies in/out on the ctx stream and synchronises; _dev takes device
 * pointers and is asynchronous on the ctx stream.                                              */
int rb_solve_batch(rb_ctx *ctx, int64_t n, const double *tkin, const double *dens, const double *cdmol,
                   double deltav_kms, double tbg, int geometry, const rb_opts *opts, double *xpop, double *tex,
                   double *tau, double *surf, int32_t *niter, int32_t *status);
int rb_solve_batch_dev(rb_ctx *ctx, int64_t n, const double *tkin, const double *dens, const double *cdmol,
                       double deltav_kms, double tbg, int geometry, const rb_opts *opts, double *xpop,
                       double *tex, double *tau, double *surf, int32_t *niter, int32_t *status);

/* ---- vectorised lnprob: replaces lnprob/lnprior/lnlike/model_lvg, one component (er1:120-181)
 * and two components (er2:122-244).  P is n x 4 / n x 8 (log10 n, T, N/dv, size per component);
 * bounds is 4 x 2 / 8 x 2 (lo, hi); opr fixed at 3 like the drivers (er1:95-96).
 * has_td = 0 reproduces T_d=None.  lnp[n] never holds NaN (emcee rejects NaN): every invalid
 * case maps to -inf exactly where the reference returns -inf.  nsolves (may be NULL) returns
 * the number of RADEX solves actually performed (prior short-circuit, er1:178-180).           */
int rb_lnprob1(rb_ctx *ctx, int64_t n, const double *P, const rb_obs *obs, const double *bounds, double tbg,
               const rb_opts *opts, double *lnp, int64_t *nsolves);
int rb_lnprob1_dev(rb_ctx *ctx, int64_t n, const double *P, const rb_obs *obs, const double *bounds,
                   double tbg, const rb_opts *opts, double *lnp, int64_t *nsolves_dev);
int rb_lnprob2(rb_ctx *ctx, int64_t n, const double *P, const rb_obs *obs, const double *bounds, int has_td,
               double t_d, double tbg, const rb_opts *opts, double *lnp, int64_t *nsolves);
int rb_lnprob2_dev(rb_ctx *ctx, int64_t n, const double *P, const rb_obs *obs, const double *bounds,
                   int has_td, double t_d, double tbg, const rb_opts *opts, double *lnp, int64_t *nsolves_dev);

/* ---- stretch move: replaces emcee.StretchMove.get_proposal + the accept loop of
 * emcee's RedBlueMove (un-vendored; call sites er1:483-499, er2:557-574).
 * All pointers are DEVICE pointers; asynchronous on the ctx stream.
 *   propose: for active walker k (global id gid0 + k*gid_stride) with position s = S[k,:], draw
 *            z = ((a-1)u+1)^2/a and a partner j uniformly from the nc complementary positions
 *            C[nc,ndim]; q = c_j - (c_j - s) z ; logfac[k] = (ndim-1) ln z.
 *   accept : accept iff logfac + lnp_new - lnp_old > ln u'; updates S, lnp_old in place and
 *            adds the number accepted to *naccept.
 * RNG is counter-based (Philox4x32-10) keyed by (seed, step, half, global walker id), so the
 * chain does not depend on how walkers are sharded over GPUs.                                  */
int rb_stretch_propose_dev(rb_ctx *ctx, int64_t ns, int32_t ndim, const double *S, int64_t nc,
                           const double *C, double a, uint64_t seed, uint64_t step, int32_t half,
                           int64_t gid0, int64_t gid_stride, double *Q, double *logfac);
int rb_stretch_accept_dev(rb_ctx *ctx, int64_t ns, int32_t ndim, double *S, double *lnp_old, const double *Q,
                          const double *lnp_new, const double *logfac, uint64_t seed, uint64_t step,
                          int32_t half, int64_t gid0, int64_t gid_stride, int64_t *naccept);

/* counters of the last solve/lnprob call on this ctx (device-measured, read back on request):
 * total matrix iterations summed over models, and kernels launched since ctx creation.        */
int rb_ctx_counters(rb_ctx *ctx, int64_t *total_iters_last, int64_t *launches_total);

/* frozen-top caching statistics of the last solve/lnprob call: stats3[0] matrix iterations that ran
 * on the cached path, [1] captures, [2] invalidations (a frozen line turned optically thick).  */
int rb_ctx_cache_stats(rb_ctx *ctx, int64_t *stats3);

/* measured FP64 FMA throughput of this GPU (TFLOP/s), the roofline denominator bench.py reports */
int rb_fp64_peak(rb_ctx *ctx, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* RADEX_B200_H */
