/*
 * iqs_oracle.c -- CPU restatement of the Intel-QS QubitRegister<ComplexDP> hot path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libiqs_b200.so, libiqs.so) never does.
 *
 * Parity: PINNED.  tests/test_oracle.py checks this file against the reference's own golden
 * vectors and against outputs of the real reference library compiled from
 * /root/reference (oracle/_ref/iqs_ref_driver, fixtures under tests/golden/ made by
 * tests/golden/make_golden.py).
 *
 * Single rank (the reference's IqsMPI=OFF build): M = num_qubits, every position is local.
 * Arithmetic follows the reference bit for bit: complex products are evaluated as
 * (ac - bd) + i(ad + bc) with separately rounded terms (libstdc++ std::complex<double>
 * operator*, GCC x86-64 baseline, no FMA) -- compile with -ffp-contract=off.
 * File:line citations are relative to /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "iqs_program.h"

typedef struct { double re, im; } cx;

static inline cx cmul(cx a, cx b) {
  cx r;
  r.re = a.re * b.re - a.im * b.im;
  r.im = a.re * b.im + a.im * b.re;
  return r;
}
static inline cx cadd(cx a, cx b) { cx r = {a.re + b.re, a.im + b.im}; return r; }
static inline double cnorm(cx a) { return a.re * a.re + a.im * a.im; } /* std::norm */

/* ---- Loop_DN, generic branch: src/highperfkernels.cpp:326-361 -------------------------- */
void oracle_gate1(cx *state, uint64_t gstart, uint64_t gend, unsigned pos, const double m[8]) {
  cx m00 = {m[0], m[1]}, m01 = {m[2], m[3]}, m10 = {m[4], m[5]}, m11 = {m[6], m[7]};
  uint64_t d = 1ull << pos;
  for (uint64_t group = gstart; group < gend; group += 2 * d)
    for (uint64_t ind0 = group; ind0 < group + d; ++ind0) {
      uint64_t i0 = ind0, i1 = ind0 + d; /* indsht0 = 0, indsht1 = 1 << pos (1q.cpp:206) */
      cx in0 = state[i0], in1 = state[i1];
      state[i0] = cadd(cmul(m00, in0), cmul(m01, in1));
      state[i1] = cadd(cmul(m10, in0), cmul(m11, in1));
    }
}

/* ---- Loop_TN: src/highperfkernels.cpp:420-436 ------------------------------------------ */
static void loop_tn(cx *state, uint64_t c11, uint64_t c12, uint64_t c13, uint64_t c21, uint64_t c22,
                    uint64_t c23, uint64_t c31, uint64_t c32, uint64_t ind_shift, const double m[8]) {
  cx m00 = {m[0], m[1]}, m01 = {m[2], m[3]}, m10 = {m[4], m[5]}, m11 = {m[6], m[7]};
  for (uint64_t l1 = c11; l1 < c12; l1 += c13)
    for (uint64_t l2 = l1 + c21; l2 < l1 + c22; l2 += c23)
      for (uint64_t ind0 = l2 + c31; ind0 < l2 + c32; ++ind0) {
        uint64_t ind1 = ind0 + ind_shift;
        cx in0 = state[ind0], in1 = state[ind1];
        state[ind0] = cadd(cmul(m00, in0), cmul(m01, in1));
        state[ind1] = cadd(cmul(m10, in0), cmul(m11, in1));
      }
}

/* ---- controlled gate, both positions local: src/qureg_applyctrl1qubitgate.cpp:312-345 --- */
void oracle_cgate1(cx *state, uint64_t sind, uint64_t eind, unsigned C, unsigned T, const double m[8]) {
  if (C > T)
    loop_tn(state, sind, eind, 1ull << (C + 1), 1ull << C, 1ull << (C + 1), 1ull << (T + 1), 0, 1ull << T,
            1ull << T, m);
  else
    loop_tn(state, sind, eind, 1ull << (T + 1), 0, 1ull << T, 1ull << (C + 1), 1ull << C, 1ull << (C + 1),
            1ull << T, m);
}

/* ---- swap family, both local, pos1 < pos2: src/qureg_applyswap.cpp:203-206 ------------- */
void oracle_swap2x2(cx *state, uint64_t L, unsigned pos1, unsigned pos2, const double m[8]) {
  uint64_t d1 = 1ull << pos1, d2 = 1ull << pos2;
  loop_tn(state, 0, L, 2 * d2, 0, d2, 2 * d1, d1, 2 * d1, d2 - d1, m);
}

/* ---- ApplyDiag, local case: src/qureg_applydiag.cpp:157-174 ----------------------------- */
void oracle_diag2(cx *state, uint64_t L, unsigned pos1, unsigned pos2, const double d[8]) {
  cx d00 = {d[0], d[1]}, d11 = {d[2], d[3]}, d22 = {d[4], d[5]}, d33 = {d[6], d[7]};
  uint64_t delta1 = 1ull << pos1, delta2 = 1ull << pos2;
  uint64_t dmin = delta1 < delta2 ? delta1 : delta2, dmax = delta1 < delta2 ? delta2 : delta1;
  for (uint64_t i = 0; i < L; i += 2 * dmax)
    for (uint64_t j = 0; j < dmax; j += 2 * dmin)
      for (uint64_t k = 0; k < dmin; ++k) {
        cx *a = &state[i + j + k];
        *a = cmul(*a, d00);
        a = &state[i + j + k + delta2];
        *a = cmul(*a, d11);
        a = &state[i + j + k + delta1];
        *a = cmul(*a, d22);
        a = &state[i + j + k + delta1 + delta2];
        *a = cmul(*a, d33);
      }
}

/* ---- Apply2QubitGate with the loop order chosen by POSITION (the reference picks it by
 *      program-qubit order, src/qureg_apply2qubitgate.cpp:37-51, which only differs -- and is
 *      wrong -- for permuted registers) -------------------------------------- */
void oracle_gate2(cx *state, uint64_t n, unsigned ph, unsigned pl, const double m[32]) {
  uint64_t dh = 1ull << ph, dl = 1ull << pl;
  uint64_t big = dh > dl ? dh : dl, small = dh > dl ? dl : dh;
  for (uint64_t i = 0; i < n; i += 2 * big)
    for (uint64_t j = 0; j < big; j += 2 * small)
      for (uint64_t k = 0; k < small; ++k) {
        uint64_t tmp = i + j + k;
        uint64_t idx[4] = {tmp, tmp + dl, tmp + dh, tmp + dh + dl};
        cx in[4] = {state[idx[0]], state[idx[1]], state[idx[2]], state[idx[3]]};
        for (int t = 0; t < 4; ++t) {
          cx acc = {0, 0};
          for (int c = 0; c < 4; ++c) {
            cx mm = {m[2 * (4 * t + c)], m[2 * (4 * t + c) + 1]};
            cx term = cmul(mm, in[c]);
            acc = c == 0 ? term : cadd(acc, term); /* ((m0 uu + m1 ud) + m2 du) + m3 dd, 2q.cpp:60 */
          }
          state[idx[t]] = acc;
        }
      }
}

/* ---- ScaleState: src/highperfkernels.cpp:505-507 ---------------------------------------- */
void oracle_scale(cx *state, uint64_t start, uint64_t end, const double s[2]) {
  cx f = {s[0], s[1]};
  if (f.re == 1.0 && f.im == 0.0) return;
  for (uint64_t i = start; i < end; ++i) state[i] = cmul(state[i], f);
}

/* ---- GetProbability: serial left-to-right sum, src/qureg_measure.cpp:150-155 ------------ */
double oracle_prob1(const cx *state, uint64_t L, unsigned pos) {
  uint64_t delta = 1ull << pos;
  double p = 0.;
  for (uint64_t i = delta; i < L; i += 2 * delta)
    for (uint64_t j = 0; j < delta; ++j) p += cnorm(state[i + j]);
  return p;
}

/* ---- ExpectationValue's reduction: src/qureg_expectval.cpp:173-185 ---------------------- */
double oracle_parity_expect(const cx *state, uint64_t L, uint64_t mask, uint64_t glb_start) {
  double v = 0;
  for (uint64_t i = 0; i < L; ++i) {
    uint64_t x = (glb_start + i) & mask;
    unsigned cnt = 0;
    for (; x; ++cnt) x &= x - 1; /* HammingWeight, :80-86 */
    if (cnt & 1) v -= cnorm(state[i]);
    else v += cnorm(state[i]);
  }
  return v;
}

double oracle_norm2(const cx *state, uint64_t L) { /* src/qureg_utils.cpp:241-245 */
  double s = 0;
  for (uint64_t i = 0; i < L; ++i) s += cnorm(state[i]);
  return s;
}

void oracle_overlap(const cx *state, const cx *psi, uint64_t L, double out[2]) { /* :282-287 */
  double re = 0, im = 0;
  for (uint64_t i = 0; i < L; ++i) {
    cx c = {psi[i].re, -psi[i].im};
    cx o = cmul(c, state[i]);
    re += o.re;
    im += o.im;
  }
  out[0] = re;
  out[1] = im;
}

double oracle_maxabsdiff(const cx *a, const cx *b, uint64_t L, const double s[2]) { /* :54-57 */
  cx f = {s[0], s[1]};
  double mx = -1.0;
  for (uint64_t i = 0; i < L; ++i) {
    cx fb = cmul(f, b[i]);
    double d = hypot(a[i].re - fb.re, a[i].im - fb.im);
    if (d > mx) mx = d;
  }
  return mx;
}

double oracle_l2diff(const cx *a, const cx *b, uint64_t L) { /* :142-146 */
  double s = 0;
  for (uint64_t i = 0; i < L; ++i) {
    cx r = {a[i].re - b[i].re, a[i].im - b[i].im};
    s += cnorm(r);
  }
  return s;
}

/* ---- CollapseQubit: src/qureg_measure.cpp:107-111 --------------------------------------- */
void oracle_collapse(cx *state, uint64_t L, unsigned pos, int value) {
  uint64_t delta = 1ull << pos;
  for (uint64_t i = value ? 0 : delta; i < L; i += 2 * delta)
    for (uint64_t j = 0; j < delta; ++j) state[i + j].re = state[i + j].im = 0.;
}

/* out[0]: any |a|^2 > tol with bit 0, out[1]: with bit 1 (IsClassicalBit, measure.cpp:34-46) */
void oracle_any_above(const cx *state, uint64_t L, unsigned pos, double tol, int out[2]) {
  out[0] = out[1] = 0;
  for (uint64_t i = 0; i < L; ++i)
    if (cnorm(state[i]) > tol) out[(i >> pos) & 1] = 1;
}

/* ---- AmplitudeWiseSum: src/qureg_utils.cpp:199-226 -------------------------------------- */
void oracle_axpy(cx *a, const cx *b, uint64_t L, const double f[2]) {
  cx ff = {f[0], f[1]};
  if (ff.re == 1.0 && ff.im == 0.0)
    for (uint64_t i = 0; i < L; ++i) a[i] = cadd(a[i], b[i]);
  else
    for (uint64_t i = 0; i < L; ++i) a[i] = cadd(a[i], cmul(b[i], ff));
}

/* ---- Permutation index maps: include/permutation.hpp:249-264 ---------------------------- */
static uint64_t data2program(const uint64_t *map, unsigned n, uint64_t v) {
  uint64_t r = 0;
  for (unsigned i = 0; i < n; ++i) r |= ((v >> map[i]) & 1ull) << i;
  return r;
}
static uint64_t program2data(const uint64_t *imap, unsigned n, uint64_t v) {
  uint64_t r = 0;
  for (unsigned i = 0; i < n; ++i) r |= ((v >> imap[i]) & 1ull) << i;
  return r;
}

/* ---- PermuteLocalQubits: src/qureg_permute.cpp:90-100 ----------------------------------- */
void oracle_permute(cx *state, unsigned n, const uint64_t *old_map, const uint64_t *new_imap) {
  uint64_t L = 1ull << n;
  cx *old = (cx *)malloc(L * sizeof(cx));
  memcpy(old, state, L * sizeof(cx));
  for (uint64_t i = 0; i < L; ++i) state[program2data(new_imap, n, data2program(old_map, n, i))] = old[i];
  free(old);
}

/* ============================== program interpreter ==================================== */
typedef struct {
  unsigned n;
  cx *state;
  uint64_t map[64], imap[64]; /* qubit -> position, position -> qubit */
  /* fusion queue (src/qureg_fusion.cpp; window entries: kind, matrix, q1, q2) */
  int fusion;
  unsigned log2llc;
  int nwin, capwin;
  iqs_op *win;
} oreg;

static void mat_h(double m[8]) { /* 1q.cpp:443-451 */
  double f = 1. / sqrt(2.);
  double t[8] = {f, 0, f, 0, f, 0, -f, 0};
  memcpy(m, t, sizeof(t));
}

/* named matrices, same libm calls as the reference front-ends */
static int named_matrix(int kind, const double *p, double m[8]) {
  memset(m, 0, 8 * sizeof(double));
  switch (kind) {
    case OP_H: case OP_CH: mat_h(m); return 1;
    case OP_X: case OP_CX: case OP_SWAP: m[2] = 1; m[4] = 1; return 1;             /* 1q.cpp:336-343 */
    case OP_Y: case OP_CY: m[3] = -1; m[5] = 1; return 1;                          /* :372-379 */
    case OP_Z: case OP_CZ: m[0] = 1; m[6] = -1; return 1;                          /* :408-415 */
    case OP_SQRTX: m[0] = .5; m[1] = .5; m[2] = .5; m[3] = -.5; m[4] = .5; m[5] = -.5; m[6] = .5; m[7] = .5; return 1; /* :354-361 */
    case OP_SQRTY: m[0] = .5; m[1] = .5; m[2] = -.5; m[3] = -.5; m[4] = .5; m[5] = .5; m[6] = .5; m[7] = .5; return 1; /* :390-397 */
    case OP_SQRTZ: case OP_CSQRTZ: m[0] = 1; m[7] = 1; return 1;                   /* :426-433 */
    case OP_T: m[0] = 1; m[6] = cos(M_PI / 4.0); m[7] = sin(M_PI / 4.0); return 1; /* :491-498 */
    case OP_RX: case OP_CRX: { double t = p[0]; m[0] = m[6] = cos(t / 2.); m[3] = m[5] = -sin(t / 2.); return 1; }        /* :284-289 */
    case OP_RY: case OP_CRY: { double t = p[0]; m[0] = m[6] = cos(t / 2.); m[2] = -sin(t / 2.); m[4] = sin(t / 2.); return 1; } /* :298-304 */
    case OP_RZ: case OP_CRZ: { double t = p[0]; m[0] = cos(t / 2.); m[1] = -sin(t / 2.); m[6] = cos(t / 2.); m[7] = sin(t / 2.); return 1; } /* :317-323 */
    case OP_RXY: { double phi = p[0], t = p[1];                                    /* :472-481 */
      m[0] = cos(t / 2.); m[6] = cos(t / 2.);
      m[2] = -sin(t / 2.) * sin(phi); m[3] = -sin(t / 2.) * cos(phi);
      m[4] = sin(t / 2.) * sin(phi);  m[5] = -sin(t / 2.) * cos(phi);
      return 1; }
    case OP_CPHASE: { double t = p[0]; m[0] = 1; m[6] = cos(t); m[7] = sin(t); return 1; } /* ctrl.cpp:606-613 */
    case OP_ISWAP: m[3] = 1; m[5] = 1; return 1;                                   /* swap.cpp:37-42 */
    case OP_SQRTISWAP: { double f = 1. / sqrt(2.); m[0] = f; m[6] = f; m[3] = f; m[5] = f; return 1; } /* :46-53 */
    case OP_4THROOTISWAP: {                                                        /* :64-78 */
      /* a = polar(.5, pi/8), b = polar(.5, 7pi/8); f0 = a - b, f1 = a + b */
      double ar = .5 * cos(M_PI / 8.), ai = .5 * sin(M_PI / 8.);
      double br = .5 * cos(7. * M_PI / 8.), bi = .5 * sin(7. * M_PI / 8.);
      m[0] = ar - br; m[1] = ai - bi; m[2] = ar + br; m[3] = ai + bi;
      m[4] = m[2]; m[5] = m[3]; m[6] = m[0]; m[7] = m[1];
      return 1; }
  }
  return 0;
}

static void helper_gate1(oreg *r, unsigned qubit, const double m[8], uint64_t s, uint64_t e) {
  oracle_gate1(r->state, s, e, (unsigned)r->map[qubit], m); /* 1q.cpp:173-230, P < M */
}
static void helper_cgate1(oreg *r, unsigned cq, unsigned tq, const double m[8], uint64_t s, uint64_t e) {
  unsigned C = (unsigned)r->map[cq], T = (unsigned)r->map[tq];
  uint64_t L = 1ull << r->n;
  /* ctrl.cpp:296-309: sub-block replay with the control above the block */
  if (C > T && C >= r->log2llc && L > (e - s)) {
    if ((s >> C) & 1) oracle_gate1(r->state, s, e, T, m);
    return;
  }
  oracle_cgate1(r->state, s, e, C, T, m);
}

static void flush_fused(oreg *r) { /* fusion.cpp:55-94 */
  uint64_t L = 1ull << r->n;
  uint64_t blocksize = (r->nwin == 1) ? L : (1ull << r->log2llc);
  for (uint64_t l = 0; l < L; l += blocksize)
    for (int k = 0; k < r->nwin; ++k) {
      iqs_op *f = &r->win[k];
      if (f->kind == OP_GATE1) helper_gate1(r, f->q0, f->p, l, l + blocksize);
      else helper_cgate1(r, f->q0, f->q1, f->p, l, l + blocksize);
    }
  r->nwin = 0;
}

static void push_win(oreg *r, int kind, unsigned q0, unsigned q1, const double m[8]) {
  if (r->nwin == r->capwin) {
    r->capwin = r->capwin ? 2 * r->capwin : 64;
    r->win = (iqs_op *)realloc(r->win, sizeof(iqs_op) * r->capwin);
  }
  iqs_op *f = &r->win[r->nwin++];
  f->kind = kind; f->q0 = (int)q0; f->q1 = (int)q1;
  memcpy(f->p, m, 8 * sizeof(double));
}

static void apply1(oreg *r, unsigned q, const double m[8]) { /* 1q.cpp:234-265 */
  unsigned pos = (unsigned)r->map[q];
  if (r->fusion) {
    if (pos < r->log2llc) { push_win(r, OP_GATE1, q, 0, m); return; }
    flush_fused(r);
  }
  helper_gate1(r, q, m, 0, 1ull << r->n);
}
static void applyc(oreg *r, unsigned c, unsigned t, const double m[8]) { /* ctrl.cpp:414-444 */
  if (r->fusion) {
    if ((unsigned)r->map[t] < r->log2llc) { push_win(r, OP_CGATE1, c, t, m); return; }
    flush_fused(r);
  }
  helper_cgate1(r, c, t, m, 0, 1ull << r->n);
}
static void applyswap(oreg *r, unsigned q1, unsigned q2, const double m[8]) { /* swap.cpp:114-209 */
  if (r->fusion) flush_fused(r);
  unsigned p1 = (unsigned)r->map[q1], p2 = (unsigned)r->map[q2];
  if (p1 > p2) { unsigned t = p1; p1 = p2; p2 = t; }
  oracle_swap2x2(r->state, 1ull << r->n, p1, p2, m);
}
static double get_prob(oreg *r, unsigned q) { return oracle_prob1(r->state, 1ull << r->n, (unsigned)r->map[q]); }

static void mat_G(double G[8], double Ginv[8]) { /* expectval.cpp:137-146 */
  double f = 1. / sqrt(2.);
  double g[8] = {f, 0, 0, -f, f, 0, 0, f};
  double gi[8] = {f, 0, f, 0, 0, f, 0, -f};
  memcpy(G, g, sizeof(g));
  memcpy(Ginv, gi, sizeof(gi));
}

static double expect1(oreg *r, unsigned q, int obs) { /* expectval.cpp:18-74 */
  double H[8], G[8], Gi[8], e;
  mat_h(H);
  mat_G(G, Gi);
  if (obs == 1) { apply1(r, q, H); e = 1. - 2. * get_prob(r, q); apply1(r, q, H); }
  else if (obs == 2) { apply1(r, q, G); e = 1. - 2. * get_prob(r, q); apply1(r, q, Gi); }
  else e = 1. - 2. * get_prob(r, q);
  return e;
}

static double expect(oreg *r, int k, const double *qs, const double *obs) { /* expectval.cpp:102-213 */
  if (k == 0) return 1.;
  if (k == 1) return expect1(r, (unsigned)qs[0], (int)obs[0]);
  double H[8], G[8], Gi[8];
  mat_h(H);
  mat_G(G, Gi);
  for (int i = 0; i < k; ++i) {
    if ((int)obs[i] == 1) apply1(r, (unsigned)qs[i], H);
    else if ((int)obs[i] == 2) apply1(r, (unsigned)qs[i], G);
  }
  uint64_t y = 0;
  for (int i = 0; i < k; ++i) y += 1ull << r->map[(unsigned)qs[i]]; /* 64-bit: the int-shift bug (:170) is not reproduced */
  double v = oracle_parity_expect(r->state, 1ull << r->n, y, 0);
  for (int i = 0; i < k; ++i) {
    if ((int)obs[i] == 1) apply1(r, (unsigned)qs[i], H);
    else if ((int)obs[i] == 2) apply1(r, (unsigned)qs[i], Gi);
  }
  return v;
}

/* Entropy() and GoogleStats() (qureg_utils.cpp:305-450): out[0] = entropy in bits, out[1] = average
 * self-information / 2^n, out[2..10] = moments m_k * 2^(n(k-1)) / k!, k = 2..10. */
void oracle_google_stats(const cx *state, uint64_t L, double out[11]) {
  double entropy = 0, avgselfinfo = 0, m[9] = {0};
  for (uint64_t i = 0; i < L; ++i) {
    double pj = state[i].re * state[i].re + state[i].im * state[i].im;
    if (pj != 0.) { double nl = log(pj); entropy -= pj * nl; avgselfinfo -= nl; }
    double pj2 = pj * pj, pj3 = pj2 * pj, pj4 = pj2 * pj2, pj5 = pj3 * pj2, pj6 = pj3 * pj3, pj7 = pj4 * pj3, pj8 = pj4 * pj4,
           pj9 = pj5 * pj4, pj10 = pj5 * pj5;
    m[0] += pj2; m[1] += pj3; m[2] += pj4; m[3] += pj5; m[4] += pj6; m[5] += pj7; m[6] += pj8; m[7] += pj9; m[8] += pj10;
  }
  double two2n = (double)L, factorial = 1.0;
  out[0] = entropy / log(2.0);
  out[1] = avgselfinfo / log(2.0) / two2n;
  for (int i = 0; i < 9; ++i) {
    int k = i + 2;
    factorial *= (double)k;
    out[2 + i] = m[i] * (pow(two2n, (double)(k - 1)) / factorial);
  }
}

/* Run a program on `state` (2^n amplitudes, interleaved). Scalars produced by value-returning
 * ops are appended to scalars[]; returns their count, or -1 on an unknown op.  map_out (n
 * entries) receives the final qubit->position map. */
int oracle_run(unsigned n, double *state, const iqs_op *ops, int nops, double *scalars, int cap, uint64_t *map_out) {
  oreg r;
  memset(&r, 0, sizeof(r));
  r.n = n;
  r.state = (cx *)state;
  for (unsigned i = 0; i < n; ++i) r.map[i] = r.imap[i] = i;
  uint64_t L = 1ull << n;
  int ns = 0;
  double m[8];
  for (int k = 0; k < nops; ++k) {
    const iqs_op *op = &ops[k];
    double out = 0;
    int has_out = 0;
    switch (op->kind) {
      case OP_GATE1: apply1(&r, op->q0, op->p); break;
      case OP_CGATE1: applyc(&r, op->q0, op->q1, op->p); break;
      case OP_SWAPLIKE: applyswap(&r, op->q0, op->q1, op->p); break;
      case OP_DIAG: /* applydiag.cpp:55-174 */
        if (r.fusion) flush_fused(&r);
        oracle_diag2(r.state, L, (unsigned)r.map[op->q0], (unsigned)r.map[op->q1], op->p);
        break;
      case OP_GATE2: oracle_gate2(r.state, L, (unsigned)r.map[op->q0], (unsigned)r.map[op->q1], op->p); break;
      case OP_TOFFOLI: { /* applytoffoli.cpp:23-47 */
        double V[8] = {.5, -.5, .5, .5, .5, .5, .5, -.5}, Vd[8] = {.5, .5, .5, -.5, .5, -.5, .5, .5}, X[8];
        named_matrix(OP_X, NULL, X);
        applyc(&r, op->q0, op->q2, V);
        applyc(&r, op->q1, op->q0, X);
        applyc(&r, op->q0, op->q2, Vd);
        applyc(&r, op->q1, op->q0, X);
        applyc(&r, op->q1, op->q2, V);
        break; }
      case OP_H: case OP_X: case OP_Y: case OP_Z: case OP_SQRTX: case OP_SQRTY: case OP_SQRTZ: case OP_T:
      case OP_RX: case OP_RY: case OP_RZ: case OP_RXY:
        named_matrix(op->kind, op->p, m);
        apply1(&r, op->q0, m);
        break;
      case OP_CH: case OP_CX: case OP_CY: case OP_CZ: case OP_CSQRTZ: case OP_CRX: case OP_CRY: case OP_CRZ: case OP_CPHASE:
        named_matrix(op->kind, op->p, m);
        applyc(&r, op->q0, op->q1, m);
        break;
      case OP_SWAP: case OP_ISWAP: case OP_SQRTISWAP: case OP_4THROOTISWAP:
        named_matrix(op->kind, op->p, m);
        applyswap(&r, op->q0, op->q1, m);
        break;
      case OP_PROB: out = get_prob(&r, op->q0); has_out = 1; break;
      case OP_EXPECT: out = expect(&r, op->q0, op->p, op->p + 16); has_out = 1; break;
      case OP_EXPECT1: out = expect1(&r, op->q0, op->q1); has_out = 1; break;
      case OP_NORM: out = sqrt(oracle_norm2(r.state, L)); has_out = 1; break;
      case OP_ENTROPY: case OP_GOOGLESTATS: { /* qureg_utils.cpp:305-450: serial sums in index order */
        if (r.fusion) flush_fused(&r);
        double st[11];
        oracle_google_stats(r.state, L, st);
        int cnt = op->kind == OP_ENTROPY ? 1 : 11;
        for (int i = 0; i < cnt; ++i) { if (ns < cap) scalars[ns] = st[i]; ++ns; }
        break; }
      case OP_GETAMP: { /* qureg_utils.cpp:84-98 with permutation.hpp program2data_ */
        if (r.fusion) flush_fused(&r);
        uint64_t gi = (uint64_t)op->p[0], di = 0;
        for (unsigned q = 0; q < n; ++q) if ((gi >> q) & 1ull) di |= 1ull << r.map[q];
        if (ns < cap) scalars[ns] = r.state[di].re; ++ns;
        if (ns < cap) scalars[ns] = r.state[di].im; ++ns;
        break; }
      case OP_NORMALIZE: { /* qureg_utils.cpp:162-168, 188-196 */
        double nrm = sqrt(oracle_norm2(r.state, L));
        cx f = {1 / nrm, 0};
        for (uint64_t i = 0; i < L; ++i) r.state[i] = cmul(r.state[i], f);
        break; }
      case OP_COLLAPSE: oracle_collapse(r.state, L, (unsigned)r.map[op->q0], op->q1 != 0); break;
      case OP_PERMUTE: { /* permute.cpp:10-24, 55-104 */
        uint64_t nm[64], nim[64];
        int same = 1;
        for (unsigned i = 0; i < n; ++i) { nm[i] = (uint64_t)op->p[i]; nim[nm[i]] = i; }
        for (unsigned i = 0; i < n; ++i) same = same && nm[i] == r.map[i];
        if (!same) oracle_permute(r.state, n, r.map, nim);
        memcpy(r.map, nm, sizeof(nm));
        memcpy(r.imap, nim, sizeof(nim));
        break; }
      case OP_EMUSWAP: { /* permute.cpp:45-52; permutation.hpp:193-201 */
        uint64_t p1 = r.map[op->q0], p2 = r.map[op->q1];
        r.map[op->q0] = p2; r.map[op->q1] = p1; r.imap[p1] = op->q1; r.imap[p2] = op->q0;
        break; }
      case OP_FUSION_ON: /* fusion.cpp:12-31 */
        if ((unsigned)op->q0 >= n) r.fusion = 0;
        else { r.log2llc = (unsigned)op->q0; r.fusion = 1; }
        break;
      case OP_FUSION_OFF: if (r.nwin) flush_fused(&r); r.fusion = 0; break;
      case OP_SPEC_ON: case OP_SPEC_OFF: case OP_SPEC2_ON: case OP_SPEC2_OFF: break; /* same values (SURVEY top table) */
      default: free(r.win); return -1;
    }
    if (has_out) { if (ns < cap) scalars[ns] = out; ++ns; }
  }
  if (r.nwin) flush_fused(&r); /* convenience: leave no pending gates at the end of a program */
  if (map_out) for (unsigned i = 0; i < n; ++i) map_out[i] = r.map[i];
  free(r.win);
  return ns;
}
