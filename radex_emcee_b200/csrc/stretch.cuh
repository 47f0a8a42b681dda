// stretch.cuh -- the stretch move over an ensemble that stays in place (second form of the C ABI, radex_b200.h):
// red/blue split as a function of (seed, step, global walker id), sub-ensembles per fitted source, per-walker
// acceptance counters, NaN detection.
//
// What it replaces: emcee's RedBlueMove.propose + StretchMove.get_proposal with the default randomize_split=True
// (un-vendored; call sites emcee/emcee_radex.py:483-499, emcee/emcee_radex_2comp.py:557-574):
//     inds = arange(nwalkers) % 2 ; random.shuffle(inds)        -> a balanced random labelling, new every step
//     for split in (0, 1): s = coords[inds == split], c = coords[inds != split]
//         zz = ((a - 1) u + 1)^2 / a ; rint = randint(len(c)) ; q = c[rint] - (c[rint] - s) zz ; factors = (ndim - 1) ln zz
//         accept where factors + lnp_new - lnp_old > ln u'
// Here the labelling is balanced per BLOCK of consecutive walkers and is drawn from a keyed bijection, so that every
// rank can evaluate it for any walker without communication and the chain does not depend on the rank count.
#pragma once

namespace st2 {

struct SplitDev {
  long long W;        // walkers per source (sub-ensemble)
  int B;              // split block, B | W, even
  int w;              // bits: 2^w >= B
  int randomize;
  unsigned k0, k1;    // Philox key of the split stream
};

// keys of block gb at step `step`: one Philox call; the stream is separated from the proposal / accept streams by
// the key (seed ^ 'SPLTrand'), not by the counter
__device__ __host__ inline void split_key_words(unsigned long long seed, unsigned &k0, unsigned &k1) {
  k0 = (unsigned)seed ^ 0x53504C54u;
  k1 = (unsigned)(seed >> 32) ^ 0x72616E64u;
}

__device__ __forceinline__ void split_keys(const SplitDev &sp, unsigned long long gb, unsigned long long step, uint32_t k[4]) {
  k[0] = (uint32_t)gb;
  k[1] = (uint32_t)(gb >> 32);
  k[2] = (uint32_t)step;
  k[3] = (uint32_t)(step >> 32);
  philox4x32_10(k, sp.k0, sp.k1);
}

// keyed bijection of [0, B): four rounds of (odd multiply + key) mod 2^w and xor-shift, each invertible on w bits;
// values >= B walk the cycle until they are back inside (every cycle of a bijection of [0, 2^w) that starts inside
// [0, B) returns there)
__device__ __forceinline__ unsigned split_perm(unsigned x, const SplitDev &sp, const uint32_t k[4]) {
  const unsigned mask = (sp.w >= 32) ? 0xffffffffu : ((1u << sp.w) - 1u);
  const int s1 = (sp.w + 1) >> 1, s2 = (sp.w >= 3) ? sp.w / 3 : 1;
  do {
    x = (x * 0x9E3779B1u + k[0]) & mask; x ^= x >> s1;
    x = (x * 0x85EBCA6Bu + k[1]) & mask; x ^= x >> s2;
    x = (x * 0xC2B2AE35u + k[2]) & mask; x ^= x >> s1;
    x = (x * 0x27D4EB2Fu + k[3]) & mask; x ^= x >> s2;
  } while (x >= (unsigned)sp.B);
  return x;
}

// local walker index of local slot k of half `half` (rank owns [gid_base, gid_base + nlocal), B | gid_base)
__device__ __forceinline__ long long slot_walker(const SplitDev &sp, unsigned long long step, int half, long long gid_base,
                                                 long long k) {
  const int hb = sp.B >> 1;
  const long long lb = k / hb;
  const int t = (int)(k - lb * hb);
  if (!sp.randomize) return lb * sp.B + 2 * t + half;
  uint32_t key[4];
  split_keys(sp, (unsigned long long)(gid_base / sp.B + lb), step, key);
  return lb * sp.B + split_perm((unsigned)(half * hb + t), sp, key);
}

__device__ __forceinline__ unsigned long long step_of(const unsigned long long *step_ptr, unsigned long long step) {
  return step + (step_ptr ? *step_ptr : 0ULL);
}

__global__ void k_pack(SplitDev sp, const unsigned long long *step_ptr, unsigned long long step_, int half, long long gid_base,
                       long long nhalf, int ndim, const double *X, double *Chalf) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= nhalf) return;
  const long long i = slot_walker(sp, step_of(step_ptr, step_), half, gid_base, k);
  for (int d = 0; d < ndim; ++d) Chalf[k * ndim + d] = X[i * ndim + d];
}

__global__ void k_propose2(SplitDev sp, const unsigned long long *step_ptr, unsigned long long step_, int half,
                           long long gid_base, long long nhalf, int ndim, const double *X, const double *Call, double a,
                           unsigned long long seed, double *Q, double *logfac, int *src_id) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= nhalf) return;
  const unsigned long long step = step_of(step_ptr, step_);
  const long long i = slot_walker(sp, step, half, gid_base, k);
  const unsigned long long gid = (unsigned long long)(gid_base + i);
  uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(step * 2ULL + (unsigned)half),
                   (uint32_t)((step * 2ULL + (unsigned)half) >> 32) & 0x7fffffffu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = u01(c[0], c[1]);
  const double sq = __dadd_rn(__dmul_rn(a - 1.0, u), 1.0);
  const double z = __dmul_rn(sq, sq) / a;
  const long long src = (long long)(gid / (unsigned long long)sp.W), nc = sp.W >> 1;
  long long j = (long long)(u01(c[2], c[3]) * (double)nc);
  if (j >= nc) j = nc - 1;
  j += src * nc;   // the complementary half of the walker's own sub-ensemble, in global slot order
  for (int d = 0; d < ndim; ++d) {
    const double cj = Call[j * ndim + d], s = X[i * ndim + d];
    Q[k * ndim + d] = __dsub_rn(cj, __dmul_rn(cj - s, z));
  }
  logfac[k] = (ndim - 1.0) * log(z);
  if (src_id) src_id[k] = (int)src;
}

__global__ void k_accept2(SplitDev sp, const unsigned long long *step_ptr, unsigned long long step_, int half,
                          long long gid_base, long long nhalf, int ndim, double *X, double *lnp, const double *Q,
                          const double *lnp_new, const double *logfac, unsigned long long seed, long long *naccept,
                          unsigned long long *nan_count, const unsigned long long *nsolves_src,
                          unsigned long long *nsolves_sum, int *accepted = nullptr) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  int isnan_ = 0;
  if (k < nhalf) {
    const unsigned long long step = step_of(step_ptr, step_);
    const long long i = slot_walker(sp, step, half, gid_base, k);
    const unsigned long long gid = (unsigned long long)(gid_base + i);
    uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(step * 2ULL + (unsigned)half),
                     ((uint32_t)((step * 2ULL + (unsigned)half) >> 32) & 0x7fffffffu) | 0x80000000u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double lnu = log(u01(c[0], c[1]));
    const double lnew = lnp_new[k];
    isnan_ = (lnew != lnew);
    const double lnpdiff = logfac[k] + lnew - lnp[i];
    const bool acc = lnpdiff > lnu;   // -inf - -inf = NaN compares false -> rejected, like numpy
    if (acc) {
      for (int d = 0; d < ndim; ++d) X[i * ndim + d] = Q[k * ndim + d];
      lnp[i] = lnew;
      if (naccept) naccept[i] += 1;   // one slot per walker per half-step: no atomics needed
    }
    if (accepted) accepted[k] = acc ? 1 : 0;
  }
  const unsigned nanb = __ballot_sync(0xffffffffu, isnan_);
  if ((threadIdx.x & 31) == 0 && nanb && nan_count) atomicAdd(nan_count, (unsigned long long)__popc(nanb));
  if (k == 0 && nsolves_sum && nsolves_src) atomicAdd(nsolves_sum, *nsolves_src);
}

// ---- speculative second half-step (small ensembles) -------------------------------------------------------------------
// A step of a small ensemble is bound by the latency of ONE solve per half-step (config 1: 50 models on 148 SMs, ~100
// sequential calls of matrix() each).  The second half-step depends on the first only through the partner c_j of every
// proposal, and c_j is one of two known points: where it stood, or the first half-step's proposal for it.  So both
// candidates are proposed before anything is solved, the nhalf + 2 nhalf models go through ONE lnprob launch, and after
// the first half-step's accept/reject the candidate the sequential move would have made is selected.  Same random
// numbers, same arithmetic per candidate: the chain is the sequential one bit for bit; a step costs one solve latency
// and 1.5 times the solves.
// Cold: the walkers of half 0 in slot order BEFORE their move (k_pack); Q0: half 0's proposals, same slot order.
__global__ void k_propose_spec(SplitDev sp, const unsigned long long *step_ptr, unsigned long long step_, long long gid_base,
                               long long nhalf, int ndim, const double *X, const double *Cold, const double *Q0, double a,
                               unsigned long long seed, double *Qa, double *Qb, double *logfac, int *jpart, int *src_a,
                               int *src_b) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= nhalf) return;
  const int half = 1;
  const unsigned long long step = step_of(step_ptr, step_);
  const long long i = slot_walker(sp, step, half, gid_base, k);
  const unsigned long long gid = (unsigned long long)(gid_base + i);
  uint32_t c[4] = {(uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)(step * 2ULL + (unsigned)half),
                   (uint32_t)((step * 2ULL + (unsigned)half) >> 32) & 0x7fffffffu};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const double u = u01(c[0], c[1]);
  const double sq = __dadd_rn(__dmul_rn(a - 1.0, u), 1.0);
  const double z = __dmul_rn(sq, sq) / a;
  const long long src = (long long)(gid / (unsigned long long)sp.W), nc = sp.W >> 1;
  long long j = (long long)(u01(c[2], c[3]) * (double)nc);
  if (j >= nc) j = nc - 1;
  j += src * nc;
  for (int d = 0; d < ndim; ++d) {
    const double s = X[i * ndim + d], co = Cold[j * ndim + d], cn = Q0[j * ndim + d];
    Qa[k * ndim + d] = __dsub_rn(co, __dmul_rn(co - s, z));
    Qb[k * ndim + d] = __dsub_rn(cn, __dmul_rn(cn - s, z));
  }
  logfac[k] = (ndim - 1.0) * log(z);
  jpart[k] = (int)j;
  if (src_a) src_a[k] = src_b[k] = (int)src;
}

// the candidate of slot k that the sequential move would have proposed: its partner moved (Qb) or stayed (Qa)
__global__ void k_select_spec(long long nhalf, int ndim, const int *acc0, const int *jpart, const double *Qa, const double *Qb,
                              const double *La, const double *Lb, double *Qsel, double *Lsel) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= nhalf) return;
  const bool moved = acc0[jpart[k]] != 0;
  for (int d = 0; d < ndim; ++d) Qsel[k * ndim + d] = moved ? Qb[k * ndim + d] : Qa[k * ndim + d];
  Lsel[k] = moved ? Lb[k] : La[k];
}

__global__ void k_step_inc(unsigned long long *step_ptr) { *step_ptr += 1ULL; }

}  // namespace st2
