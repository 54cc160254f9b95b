/*
 * TEST INFRASTRUCTURE ONLY -- NOT PRODUCT CODE.  See grid_oracle.h.
 *
 * Serial plain-C restatement of the arithmetic defined by CP2K's REF grid
 * backend.  Every function cites the reference file:line it restates
 * (paths relative to /root/reference/src/grid).  The restatement is written
 * for clarity, not speed:
 *   - density-matrix transforms are expressed through a small operator algebra
 *     (derivative / position operators acting on Cartesian Gaussians) instead
 *     of one hand-written routine per grid_func value;
 *   - the re-centring of the polynomial prefactors is a polynomial
 *     convolution;
 *   - the triclinic ("general") path evaluates the Gaussian directly instead
 *     of through three 2-D product tables;
 *   - points are visited one at a time (no symmetric pairing, no SIMD).
 * Loop bounds (discretised radius, ceil(-1e-8 - ...) rounding, quadratic
 * x-bounds, border masks) follow the reference exactly because they decide
 * WHICH grid points receive a contribution.
 */
#include "grid_oracle.h"

#include <assert.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static oracle_counters g_cnt;

void grid_oracle_reset_counters(void) { memset(&g_cnt, 0, sizeof(g_cnt)); }
void grid_oracle_get_counters(oracle_counters *out) { *out = g_cnt; }

/* ---------------------------------------------------------------------------
 * Index helpers, cf. common/grid_common.h:56-163.
 * ------------------------------------------------------------------------- */
static inline int ncoset(const int l) {
  return (l < 0) ? 0 : ((l + 1) * (l + 2) * (l + 3)) / 6;
}
static inline int coset(const int lx, const int ly, const int lz) {
  const int l = lx + ly + lz;
  return ncoset(l - 1) + ((l - lx) * (l - lx + 1)) / 2 + lz;
}
static inline int pmod(const int a, const int m) { return ((a % m) + m) % m; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

typedef struct {
  int l[3];
} orb;
static inline int oidx(const orb a) { return coset(a.l[0], a.l[1], a.l[2]); }
static inline orb oup(const int i, orb a) {
  a.l[i] += 1;
  return a;
}
static inline orb odn(const int i, orb a) { /* clamps at zero */
  a.l[i] = imax(0, a.l[i] - 1);
  return a;
}

/* ---------------------------------------------------------------------------
 * Operator algebra for the density transforms, restating
 * common/grid_prepare_pab.h:33-303 (dispatch :315-430, l-range :447-517).
 *
 * A primitive Cartesian Gaussian g(l) with exponent z obeys
 *    d/dx_i g(l)  =  l_i g(l - e_i)  -  2 z g(l + e_i)
 *    x_j g(l)     =  g(l + e_j)                       (x about its own centre)
 * Each grid_func is a short sum of (operator on a) x (operator on b).
 * ------------------------------------------------------------------------- */
enum opk { OP_ID, OP_D, OP_N, OP_DD, OP_RD, OP_R, OP_CORE };
typedef struct {
  enum opk k;
  int i, j;
} opr;
typedef struct {
  orb o;
  double c;
} oterm;

static int op_expand(const opr op, const orb l, const double z, oterm out[4]) {
  const int i = op.i, j = op.j;
  switch (op.k) {
  case OP_ID:
    out[0] = (oterm){l, 1.0};
    return 1;
  case OP_D: /* grid derivative */
    out[0] = (oterm){odn(i, l), (double)l.l[i]};
    out[1] = (oterm){oup(i, l), -2.0 * z};
    return 2;
  case OP_N: /* derivative wrt. the nuclear position: opposite sign */
    out[0] = (oterm){odn(i, l), -(double)l.l[i]};
    out[1] = (oterm){oup(i, l), +2.0 * z};
    return 2;
  case OP_DD:
    if (i != j) { /* mixed second derivative */
      out[0] = (oterm){odn(i, odn(j, l)), (double)(l.l[i] * l.l[j])};
      out[1] = (oterm){oup(i, odn(j, l)), -2.0 * z * l.l[j]};
      out[2] = (oterm){odn(i, oup(j, l)), -2.0 * z * l.l[i]};
      out[3] = (oterm){oup(i, oup(j, l)), 4.0 * z * z};
      return 4;
    } else { /* pure second derivative */
      out[0] = (oterm){odn(i, odn(i, l)), (double)(l.l[i] * (l.l[i] - 1))};
      out[1] = (oterm){l, -2.0 * z * (2 * l.l[i] + 1)};
      out[2] = (oterm){oup(i, oup(i, l)), 4.0 * z * z};
      return 3;
    }
  case OP_RD: /* x_j * d/dx_i ; the reference raises j first, then lowers i */
    out[0] = (oterm){odn(i, oup(j, l)), (double)l.l[i]};
    out[1] = (oterm){oup(i, oup(j, l)), -2.0 * z};
    return 2;
  case OP_R:
    out[0] = (oterm){oup(j, l), 1.0};
    return 1;
  case OP_CORE:
    out[0] = (oterm){oup(i, l), 2.0 * z};
    return 1;
  }
  return 0;
}

typedef struct {
  double c;
  opr a, b;
} fterm;

static int func_terms(const int func, fterm t[3], int ldiff[4]) {
  /* ldiff = {la_max_diff, la_min_diff, lb_max_diff, lb_min_diff} */
  const opr ID = {OP_ID, 0, 0};
  int n = 0;
  ldiff[0] = +1, ldiff[1] = -1, ldiff[2] = +1, ldiff[3] = -1;
  if (func == 100) { /* AB */
    ldiff[0] = ldiff[1] = ldiff[2] = ldiff[3] = 0;
    t[n++] = (fterm){1.0, ID, ID};
  } else if (func == 200) { /* DADB: 0.5 * grad a . grad b */
    for (int i = 0; i < 3; i++)
      t[n++] = (fterm){0.5, {OP_D, i, 0}, {OP_D, i, 0}};
  } else if (301 <= func && func <= 303) { /* a db - da b */
    const int i = func - 301;
    t[n++] = (fterm){+1.0, ID, {OP_D, i, 0}};
    t[n++] = (fterm){-1.0, {OP_D, i, 0}, ID};
  } else if (411 <= func && func <= 433) { /* a r_j d_i b - d_i a r_j b */
    const int i = (func - 400) / 10 - 1, j = (func - 400) % 10 - 1;
    if (i < 0 || i > 2 || j < 0 || j > 2)
      return -1;
    ldiff[2] = +2;
    t[n++] = (fterm){+1.0, ID, {OP_RD, i, j}};
    t[n++] = (fterm){-1.0, {OP_D, i, 0}, {OP_R, 0, j}};
  } else if (501 <= func && func <= 503) { /* a db + da b */
    const int i = func - 501;
    t[n++] = (fterm){1.0, ID, {OP_D, i, 0}};
    t[n++] = (fterm){1.0, {OP_D, i, 0}, ID};
  } else if (601 <= func && func <= 603) { /* d_i a d_i b */
    const int i = func - 601;
    t[n++] = (fterm){1.0, {OP_D, i, 0}, {OP_D, i, 0}};
  } else if (701 <= func && func <= 703) { /* d_i d_j a * d_i d_j b */
    const int i = func - 701, j = (i + 1) % 3;
    ldiff[0] = +2, ldiff[1] = -2, ldiff[2] = +2, ldiff[3] = -2;
    t[n++] = (fterm){1.0, {OP_DD, i, j}, {OP_DD, i, j}};
  } else if (801 <= func && func <= 803) { /* d_i^2 a * d_i^2 b */
    const int i = func - 801;
    ldiff[0] = +2, ldiff[1] = -2, ldiff[2] = +2, ldiff[3] = -2;
    t[n++] = (fterm){1.0, {OP_DD, i, i}, {OP_DD, i, i}};
  } else if (901 <= func && func <= 903) {
    t[n++] = (fterm){1.0, {OP_N, func - 901, 0}, ID};
  } else if (904 <= func && func <= 906) {
    t[n++] = (fterm){1.0, ID, {OP_N, func - 904, 0}};
  } else if (1001 <= func && func <= 1003) {
    t[n++] = (fterm){1.0, {OP_CORE, func - 1001, 0}, ID};
  } else {
    return -1;
  }
  return n;
}

/* cab[idx(b')][idx(a')] += coefficient * pab[o2+idx(b)][o1+idx(a)],
 * cf. ref/grid_ref_prepare_pab.c:54-82. */
static void prepare_cab(const int func, const int o1, const int o2,
                        const int la_max, const int la_min, const int lb_max,
                        const int lb_min, const double zeta, const double zetb,
                        const int n1, const double *pab, const int n1_cab,
                        double *cab) {
  fterm ft[3];
  int ldiff[4];
  const int nft = func_terms(func, ft, ldiff);
  assert(nft > 0);
  for (int la = la_min; la <= la_max; la++)
    for (int ax = 0; ax <= la; ax++)
      for (int ay = 0; ay <= la - ax; ay++) {
        const orb a = {{ax, ay, la - ax - ay}};
        for (int lb = lb_min; lb <= lb_max; lb++)
          for (int bx = 0; bx <= lb; bx++)
            for (int by = 0; by <= lb - bx; by++) {
              const orb b = {{bx, by, lb - bx - by}};
              const double p = pab[(o2 + oidx(b)) * n1 + o1 + oidx(a)];
              for (int t = 0; t < nft; t++) {
                oterm ta[4], tb[4];
                const int na = op_expand(ft[t].a, a, zeta, ta);
                const int nb = op_expand(ft[t].b, b, zetb, tb);
                for (int ia = 0; ia < na; ia++)
                  for (int ib = 0; ib < nb; ib++)
                    cab[oidx(tb[ib].o) * n1_cab + oidx(ta[ia].o)] +=
                        ft[t].c * ta[ia].c * tb[ib].c * p;
              }
            }
      }
}

/* ---------------------------------------------------------------------------
 * Re-centring (x-a)^la (x-b)^lb = sum_k alpha_k (x-p)^k, restating
 * ref/grid_ref_collint.h:827-911 as a polynomial convolution.
 * ------------------------------------------------------------------------- */
static double binom(const int n, const int k) {
  double r = 1.0;
  for (int i = 1; i <= k; i++)
    r = r * (double)(n - k + i) / (double)i;
  return r;
}

/* alpha[d][la][lb][k], leading dims (la_max+1),(lb_max+1),(lp+1) */
static void make_alpha(const int la_max, const int lb_max, const double ra[3],
                       const double rb[3], const double rp[3], double *alpha) {
  const int lp = la_max + lb_max;
  memset(alpha, 0, sizeof(double) * 3 * (la_max + 1) * (lb_max + 1) * (lp + 1));
  for (int d = 0; d < 3; d++) {
    const double pa = rp[d] - ra[d], pb = rp[d] - rb[d];
    for (int la = 0; la <= la_max; la++)
      for (int lb = 0; lb <= lb_max; lb++) {
        double *al =
            &alpha[((d * (la_max + 1) + la) * (lb_max + 1) + lb) * (lp + 1)];
        for (int s = 0; s <= la; s++) {   /* power of (x-p) taken from a */
          for (int t = 0; t <= lb; t++) { /* power of (x-p) taken from b */
            al[s + t] += binom(la, s) * pow(pa, la - s) * binom(lb, t) *
                         pow(pb, lb - t);
          }
        }
      }
  }
}

/* dir=+1: cxyz += T cab ;  dir=-1: cab += T^t cxyz.  cxyz is dense
 * [lz][ly][lx] with edge lp+1 like the reference. */
static void cab_cxyz(const int dir, const int la_max, const int la_min,
                     const int lb_max, const int lb_min,
                     const double prefactor, const double ra[3],
                     const double rb[3], const double rp[3], double *cab,
                     double *cxyz) {
  const int lp = la_max + lb_max, lp1 = lp + 1;
  double *alpha =
      malloc(sizeof(double) * 3 * (la_max + 1) * (lb_max + 1) * lp1);
  make_alpha(la_max, lb_max, ra, rb, rp, alpha);
#define AL(d, la, lb, k)                                                       \
  alpha[(((d) * (la_max + 1) + (la)) * (lb_max + 1) + (lb)) * lp1 + (k)]
  for (int la = la_min; la <= la_max; la++)
    for (int ax = 0; ax <= la; ax++)
      for (int ay = 0; ay <= la - ax; ay++) {
        const int az = la - ax - ay;
        for (int lb = lb_min; lb <= lb_max; lb++)
          for (int bx = 0; bx <= lb; bx++)
            for (int by = 0; by <= lb - bx; by++) {
              const int bz = lb - bx - by;
              const int ic =
                  coset(bx, by, bz) * ncoset(la_max) + coset(ax, ay, az);
              for (int kz = 0; kz <= az + bz; kz++)
                for (int ky = 0; ky <= ay + by; ky++)
                  for (int kx = 0; kx <= ax + bx; kx++) {
                    const double w = AL(0, ax, bx, kx) * AL(1, ay, by, ky) *
                                     AL(2, az, bz, kz) * prefactor;
                    const int ix = (kz * lp1 + ky) * lp1 + kx;
                    if (dir > 0)
                      cxyz[ix] += cab[ic] * w;
                    else
                      cab[ic] += cxyz[ix] * w;
                  }
            }
      }
#undef AL
  free(alpha);
}

/* ---------------------------------------------------------------------------
 * Orthorhombic cell, restating ref/grid_ref_collint.h:206-327 (geometry,
 * tables, index map) and :30-200 (the separable sweep z -> y -> x).
 * ------------------------------------------------------------------------- */
/* distance (in grid points) of cube index k from the centre pair {0,1},
 * cf. the (2*k1-1)/2 rule at ref/grid_ref_collint.h:139-146, 46-48 */
static inline int pair_dist(const int k) { return (k <= 0) ? -k : k - 1; }

static void ortho_map(const int dir, const int lp, const double zetp,
                      const double *dh, const double *dh_inv,
                      const double rp[3], const int npts_global[3],
                      const int npts_local[3], const int shift_local[3],
                      const double radius, double *cxyz, double *grid) {
  const int lp1 = lp + 1;
  const double h[3] = {dh[0], dh[4], dh[8]};
  const double hinv[3] = {dh_inv[0], dh_inv[4], dh_inv[8]};

  /* centre cell and offset of rp inside it (:222-235) */
  int center[3];
  double roff[3];
  for (int i = 0; i < 3; i++) {
    double s = 0.0;
    for (int j = 0; j < 3; j++)
      s += dh_inv[j * 3 + i] * rp[j];
    center[i] = (int)floor(s);
    roff[i] = rp[i] - ((double)center[i]) * h[i];
  }

  /* discretised radius (:237-239) and cube bounds (:242-254) */
  const double drmin = fmin(h[0], fmin(h[1], h[2]));
  const double R = drmin * fmax(1.0, ceil(radius / drmin));
  int lb[3], ub[3], cmax = 0;
  for (int i = 0; i < 3; i++) {
    lb[i] = (int)ceil(-1e-8 - R * hinv[i]);
    ub[i] = 1 - lb[i];
    cmax = imax(cmax, ub[i]);
    if (npts_global[i] != npts_local[i]) { /* non-periodic: must fit */
      const int off =
          pmod(center[i] + lb[i] - shift_local[i], npts_global[i]) - lb[i];
      assert(off + ub[i] < npts_local[i]);
      assert(off + lb[i] >= 0);
    }
  }
  const int w = 2 * cmax + 1;

  /* 1-D tables  pol[d][l][g] = (x_g - xp)^l exp(-zetp (x_g - xp)^2)  built by
   * stepping outwards from the centre pair with the product rule
   *   E(x+d) = E(x) * q(x) * e1,  q(x+d) = q(x) * e1^2      (:259-293)     */
  double *pol = calloc((size_t)3 * lp1 * w, sizeof(double));
  int *map = malloc(sizeof(int) * 3 * w);
#define POL(d, l, g) pol[((d) * lp1 + (l)) * w + (g) + cmax]
  for (int d = 0; d < 3; d++) {
    const double dr = h[d], ro = roff[d];
    const double e1 = exp(-zetp * dr * dr), e2 = e1 * e1;
    /* downwards: start from the point g=+1 and walk to g=0,-1,...,lb */
    double E = exp(-zetp * (dr - ro) * (dr - ro));
    double q = exp(-2.0 * zetp * (dr - ro) * (-dr));
    for (int g = 0; g >= lb[d]; g--) {
      const double x = g * dr - ro;
      E *= q * e1;
      q *= e2;
      double v = E;
      for (int l = 0; l <= lp; l++, v *= x)
        POL(d, l, g) = v;
    }
    /* upwards: start from the point g=0 and walk to g=1,2,...,ub */
    E = exp(-zetp * ro * ro);
    q = exp(-2.0 * zetp * (-ro) * dr);
    for (int g = 1; g <= ub[d]; g++) {
      const double x = g * dr - ro;
      E *= q * e1;
      q *= e2;
      double v = E;
      for (int l = 0; l <= lp; l++, v *= x)
        POL(d, l, g) = v;
    }
    for (int g = -cmax; g <= cmax; g++)
      map[d * w + g + cmax] =
          pmod(center[d] + g - shift_local[d], npts_global[d]); /* :297-305 */
  }

  double *cxy = malloc(sizeof(double) * lp1 * lp1);
  double *cx = malloc(sizeof(double) * lp1);
  const size_t sy = npts_local[0], sz = (size_t)npts_local[0] * npts_local[1];

  for (int k = lb[2]; k <= ub[2]; k++) { /* kstart == lb[2] (:308) */
    const double kr = pair_dist(k) * h[2];
    const double krem = R * R - kr * kr; /* :144-147 */
    const int jstart = (int)ceil(-1e-8 - sqrt(fmax(0.0, krem)) * hinv[1]);
    const size_t kg = (size_t)map[2 * w + k + cmax];
    g_cnt.nplanes += 1;

    memset(cxy, 0, sizeof(double) * lp1 * lp1);
    if (dir > 0) /* cxyz -> cxy (:173-200) */
      for (int lz = 0; lz <= lp; lz++)
        for (int ly = 0; ly <= lp - lz; ly++)
          for (int lx = 0; lx <= lp - lz - ly; lx++)
            cxy[ly * lp1 + lx] +=
                cxyz[(lz * lp1 + ly) * lp1 + lx] * POL(2, lz, k);

    for (int j = jstart; j <= 1 - jstart; j++) {
      const double jr = pair_dist(j) * h[1];
      const double jrem = krem - jr * jr; /* :46-50 */
      const int istart = (int)ceil(-1e-8 - sqrt(fmax(0.0, jrem)) * hinv[0]);
      const size_t jg = (size_t)map[1 * w + j + cmax];
      g_cnt.nrows += 1;

      memset(cx, 0, sizeof(double) * lp1);
      if (dir > 0) /* cxy -> cx (:99-126) */
        for (int ly = 0; ly <= lp; ly++)
          for (int lx = 0; lx <= lp - ly; lx++)
            cx[lx] += cxy[ly * lp1 + lx] * POL(1, ly, j);

      for (int i = istart; i <= 1 - istart; i++) { /* :52-92 */
        const size_t ig = (size_t)map[0 * w + i + cmax];
        double *g = &grid[kg * sz + jg * sy + ig];
        if (dir > 0) {
          double v = 0.0;
          for (int lx = 0; lx <= lp; lx++)
            v += cx[lx] * POL(0, lx, i);
          *g += v;
        } else {
          const double v = *g;
          for (int lx = 0; lx <= lp; lx++)
            cx[lx] += v * POL(0, lx, i);
        }
      }
      g_cnt.npts += 2 - 2 * istart;

      if (dir < 0)
        for (int ly = 0; ly <= lp; ly++)
          for (int lx = 0; lx <= lp - ly; lx++)
            cxy[ly * lp1 + lx] += cx[lx] * POL(1, ly, j);
    }

    if (dir < 0)
      for (int lz = 0; lz <= lp; lz++)
        for (int ly = 0; ly <= lp - lz; ly++)
          for (int lx = 0; lx <= lp - lz - ly; lx++)
            cxyz[(lz * lp1 + ly) * lp1 + lx] +=
                cxy[ly * lp1 + lx] * POL(2, lz, k);
  }
#undef POL
  free(cx);
  free(cxy);
  free(map);
  free(pol);
}

/* ---------------------------------------------------------------------------
 * General (triclinic, or border-masked) cell, restating
 * ref/grid_ref_collint.h:582-691 (bounds, masks), :416-483 (quadratic row
 * bounds), :333-389 (inner loop) and :697-762 (Cartesian -> lattice
 * polynomial basis, done here by polynomial multiplication).
 * ------------------------------------------------------------------------- */
/* c = a * b for dense trivariate polynomials [k][j][i] truncated at degree lp */
static void poly3_mul(const int lp, const double *a, const double *b,
                      double *c) {
  const int n = lp + 1;
  memset(c, 0, sizeof(double) * n * n * n);
  for (int ak = 0; ak < n; ak++)
    for (int aj = 0; aj < n - ak; aj++)
      for (int ai = 0; ai < n - ak - aj; ai++) {
        const double av = a[(ak * n + aj) * n + ai];
        if (av == 0.0)
          continue;
        for (int bk = 0; ak + bk < n; bk++)
          for (int bj = 0; ak + bk + aj + bj < n; bj++)
            for (int bi = 0; ak + bk + aj + bj + ai + bi < n; bi++)
              c[((ak + bk) * n + aj + bj) * n + ai + bi] +=
                  av * b[(bk * n + bj) * n + bi];
      }
}

static void cxyz_cijk(const int dir, const int lp, const double *dh,
                      double *cxyz, double *cijk) {
  const int n = lp + 1, n3 = n * n * n;
  /* pw[c][p] = (linear form for Cartesian component c)^p as a polynomial in
   * the lattice offsets (di,dj,dk):  x_c = di*dh[0][c]+dj*dh[1][c]+dk*dh[2][c]
   */
  double *pw = calloc((size_t)3 * n * n3, sizeof(double));
  double *lin = calloc(n3, sizeof(double)), *t1 = malloc(sizeof(double) * n3),
         *t2 = malloc(sizeof(double) * n3);
  for (int c = 0; c < 3; c++) {
    memset(lin, 0, sizeof(double) * n3);
    if (lp >= 1) {
      lin[(0 * n + 0) * n + 1] = dh[0 * 3 + c];
      lin[(0 * n + 1) * n + 0] = dh[1 * 3 + c];
      lin[(1 * n + 0) * n + 0] = dh[2 * 3 + c];
    }
    pw[(c * n + 0) * n3 + 0] = 1.0;
    for (int p = 1; p <= lp; p++)
      poly3_mul(lp, &pw[(c * n + p - 1) * n3], lin, &pw[(c * n + p) * n3]);
  }
  for (int lz = 0; lz <= lp; lz++)
    for (int ly = 0; ly <= lp - lz; ly++)
      for (int lx = 0; lx <= lp - lz - ly; lx++) {
        poly3_mul(lp, &pw[(0 * n + lx) * n3], &pw[(1 * n + ly) * n3], t1);
        poly3_mul(lp, t1, &pw[(2 * n + lz) * n3], t2);
        const int ic = (lz * n + ly) * n + lx;
        for (int q = 0; q < n3; q++) {
          if (dir > 0)
            cijk[q] += cxyz[ic] * t2[q];
          else
            cxyz[ic] += cijk[q] * t2[q];
        }
      }
  free(t2);
  free(t1);
  free(lin);
  free(pw);
}

static void general_map(const int dir, const int border_mask, const int lp,
                        const double zetp, const double *dh,
                        const double *dh_inv, const double rp[3],
                        const int npts_global[3], const int npts_local[3],
                        const int shift_local[3], const int border_width[3],
                        const double radius, double *cxyz, double *grid) {
  const int n = lp + 1;
  double *cijk = calloc((size_t)n * n * n, sizeof(double));
  if (dir > 0)
    cxyz_cijk(+1, lp, dh, cxyz, cijk);

  /* admissible local index window, shrunk by the border mask (:591-609) */
  int bnd[3][2];
  for (int d = 0; d < 3; d++) {
    bnd[d][0] = 0;
    bnd[d][1] = npts_local[d] - 1;
    if (border_mask & (1 << (2 * d)))
      bnd[d][0] += border_width[d];
    if (border_mask & (1 << (2 * d + 1)))
      bnd[d][1] -= border_width[d];
  }

  /* centre in lattice coordinates (:611-618) */
  double gp[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      gp[i] += dh_inv[j * 3 + i] * rp[j];

  /* index box from the 27 probes of the bounding cube (:624-640) */
  int imin3[3] = {INT_MAX, INT_MAX, INT_MAX};
  int imax3[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (int a = -1; a <= 1; a++)
    for (int b = -1; b <= 1; b++)
      for (int c = -1; c <= 1; c++) {
        const double x = rp[0] + a * radius, y = rp[1] + b * radius,
                     z = rp[2] + c * radius;
        for (int d = 0; d < 3; d++) {
          const double s =
              dh_inv[0 * 3 + d] * x + dh_inv[1 * 3 + d] * y + dh_inv[2 * 3 + d] * z;
          imin3[d] = imin(imin3[d], (int)floor(s));
          imax3[d] = imax(imax3[d], (int)ceil(s));
        }
      }

  double *cij = malloc(sizeof(double) * n * n), *ci = malloc(sizeof(double) * n);
  const size_t sy = npts_local[0], sz = (size_t)npts_local[0] * npts_local[1];

  for (int k = imin3[2]; k <= imax3[2]; k++) {
    const int kg = pmod(k - shift_local[2], npts_global[2]);
    if (kg < bnd[2][0] || bnd[2][1] < kg)
      continue;
    const double dk = k - gp[2];
    g_cnt.nplanes += 1;

    memset(cij, 0, sizeof(double) * n * n);
    if (dir > 0) { /* :489-508 */
      double dkp = 1.0;
      for (int kl = 0; kl <= lp; kl++, dkp *= dk)
        for (int jl = 0; jl <= lp - kl; jl++)
          for (int il = 0; il <= lp - kl - jl; il++)
            cij[jl * n + il] += cijk[(kl * n + jl) * n + il] * dkp;
    }

    for (int j = imin3[1]; j <= imax3[1]; j++) {
      const int jg = pmod(j - shift_local[1], npts_global[1]);
      if (jg < bnd[1][0] || bnd[1][1] < jg)
        continue;
      const double dj = j - gp[1];

      /* |i*dh[0] + v|^2 = radius^2 solved for i (:450-463) */
      double qa = 0.0, qb = 0.0, qc = 0.0;
      for (int c = 0; c < 3; c++) {
        const double v = (0 - gp[0]) * dh[0 * 3 + c] +
                         (j - gp[1]) * dh[1 * 3 + c] +
                         (k - gp[2]) * dh[2 * 3 + c];
        qa += dh[0 * 3 + c] * dh[0 * 3 + c];
        qb += 2.0 * v * dh[0 * 3 + c];
        qc += v * v;
      }
      const double disc = qb * qb - 4.0 * qa * (qc - radius * radius);
      if (!(0.0 < disc))
        continue;
      const double sq = sqrt(disc), inv2a = 1.0 / (2.0 * qa);
      const int ismin = (int)ceil((-qb - sq) * inv2a);
      const int ismax = (int)floor((-qb + sq) * inv2a);
      g_cnt.nrows += 1;

      memset(ci, 0, sizeof(double) * n);
      if (dir > 0) { /* :395-410 */
        double djp = 1.0;
        for (int jl = 0; jl <= lp; jl++, djp *= dj)
          for (int il = 0; il <= lp - jl; il++)
            ci[il] += cij[jl * n + il] * djp;
      }

      for (int i = ismin; i <= ismax; i++) { /* :344-388 */
        const int ig = pmod(i - shift_local[0], npts_global[0]);
        if (ig < bnd[0][0] || bnd[0][1] < ig)
          continue;
        const double di = i - gp[0];
        double r2 = 0.0;
        for (int c = 0; c < 3; c++) {
          const double rc =
              di * dh[0 * 3 + c] + dj * dh[1 * 3 + c] + dk * dh[2 * 3 + c];
          r2 += rc * rc;
        }
        double dip = exp(-zetp * r2);
        double *g = &grid[(size_t)kg * sz + (size_t)jg * sy + ig];
        g_cnt.npts += 1;
        if (dir > 0) {
          double v = 0.0;
          for (int il = 0; il <= lp; il++, dip *= di)
            v += ci[il] * dip;
          *g += v;
        } else {
          const double v = *g;
          for (int il = 0; il <= lp; il++, dip *= di)
            ci[il] += v * dip;
        }
      }

      if (dir < 0) {
        double djp = 1.0;
        for (int jl = 0; jl <= lp; jl++, djp *= dj)
          for (int il = 0; il <= lp - jl; il++)
            cij[jl * n + il] += ci[il] * djp;
      }
    }

    if (dir < 0) {
      double dkp = 1.0;
      for (int kl = 0; kl <= lp; kl++, dkp *= dk)
        for (int jl = 0; jl <= lp - kl; jl++)
          for (int il = 0; il <= lp - kl - jl; il++)
            cijk[(kl * n + jl) * n + il] += cij[jl * n + il] * dkp;
    }
  }
  if (dir < 0)
    cxyz_cijk(-1, lp, dh, cxyz, cijk);
  free(ci);
  free(cij);
  free(cijk);
}

/* ---------------------------------------------------------------------------
 * One Gaussian product <-> grid, restating ref/grid_ref_collint.h:917-965
 * (radius screen, product centre, prefactor) and :800-821 (path selection).
 * Returns false when the product is skipped.
 * ------------------------------------------------------------------------- */
static bool cab_grid(const int dir, const bool ortho, const int border_mask,
                     const int la_max, const int la_min, const int lb_max,
                     const int lb_min, const double zeta, const double zetb,
                     const double rscale, const double *dh,
                     const double *dh_inv, const double ra[3],
                     const double rab[3], const int npts_global[3],
                     const int npts_local[3], const int shift_local[3],
                     const int border_width[3], const double radius,
                     double *cab, double *grid) {
  double dh_max = 0.0;
  for (int i = 0; i < 9; i++)
    dh_max = fmax(dh_max, fabs(dh[i]));
  if (2.0 * radius < dh_max)
    return false;

  const double zetp = zeta + zetb, f = zetb / zetp;
  const double rab2 = rab[0] * rab[0] + rab[1] * rab[1] + rab[2] * rab[2];
  const double prefactor = rscale * exp(-zeta * f * rab2);
  double rp[3], rb[3];
  for (int i = 0; i < 3; i++) {
    rp[i] = ra[i] + f * rab[i];
    rb[i] = ra[i] + rab[i];
  }
  const int lp = la_max + lb_max, n = lp + 1;
  double *cxyz = calloc((size_t)n * n * n, sizeof(double));
  const double before_pts = g_cnt.npts, before_rows = g_cnt.nrows,
               before_planes = g_cnt.nplanes;

  if (dir > 0)
    cab_cxyz(+1, la_max, la_min, lb_max, lb_min, prefactor, ra, rb, rp, cab,
             cxyz);
  const bool use_ortho = ortho && border_mask == 0; /* :810 */
  if (use_ortho)
    ortho_map(dir, lp, zetp, dh, dh_inv, rp, npts_global, npts_local,
              shift_local, radius, cxyz, grid);
  else
    general_map(dir, border_mask, lp, zetp, dh, dh_inv, rp, npts_global,
                npts_local, shift_local, border_width, radius, cxyz, grid);
  if (dir < 0)
    cab_cxyz(-1, la_max, la_min, lb_max, lb_min, prefactor, ra, rb, rp, cab,
             cxyz);
  free(cxyz);

  /* model flop count of the REF loop nest */
  const double pts = g_cnt.npts - before_pts, rows = g_cnt.nrows - before_rows,
               planes = g_cnt.nplanes - before_planes;
  const double t2 = 0.5 * (lp + 1) * (lp + 2);
  if (use_ortho)
    g_cnt.flops += pts * (2.0 * (lp + 1) + (dir > 0 ? 1 : 0)) +
                   rows * 2.0 * t2 + planes * 2.0 * ncoset(lp);
  else
    g_cnt.flops += pts * (3.0 * (lp + 1) + (dir > 0 ? 4 : 3)) +
                   rows * (2.0 * t2 + (lp + 1) + 40.0) +
                   planes * (2.0 * ncoset(lp) + (lp + 1));
  g_cnt.ntasks += 1;
  return true;
}

void grid_oracle_collocate_pgf_product(
    bool ortho, int border_mask, int func, int la_max, int la_min, int lb_max,
    int lb_min, double zeta, double zetb, double rscale, const double *dh,
    const double *dh_inv, const double *ra, const double *rab,
    const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, double radius, int o1, int o2, int n1, int n2,
    const double *pab, double *grid) {
  (void)n2;
  fterm ft[3];
  int ld[4];
  if (func_terms(func, ft, ld) < 0) {
    fprintf(stderr, "grid_oracle: unknown grid_func %d\n", func);
    abort();
  }
  /* ref/grid_ref_collocate.c:33-44 */
  const int la_max_c = la_max + ld[0], la_min_c = imax(la_min + ld[1], 0);
  const int lb_max_c = lb_max + ld[2], lb_min_c = imax(lb_min + ld[3], 0);
  const int n1c = ncoset(la_max_c), n2c = ncoset(lb_max_c);
  double *cab = calloc((size_t)n1c * n2c, sizeof(double));
  prepare_cab(func, o1, o2, la_max, la_min, lb_max, lb_min, zeta, zetb, n1, pab,
              n1c, cab);
  cab_grid(+1, ortho, border_mask, la_max_c, la_min_c, lb_max_c, lb_min_c, zeta,
           zetb, rscale, dh, dh_inv, ra, rab, npts_global, npts_local,
           shift_local, border_width, radius, cab, grid);
  free(cab);
}

/* ---------------------------------------------------------------------------
 * Matrix elements, forces and virial from the integrated cab, restating
 * common/grid_process_vab.h:29-205 (l-range growth :222-251).
 * ------------------------------------------------------------------------- */
typedef struct {
  const double *cab;
  int m1;
  double zeta, zetb;
  const double *rab;
} pctx;
static inline double C(const pctx *p, const orb a, const orb b) {
  return p->cab[oidx(b) * p->m1 + oidx(a)];
}
/* what: 0 = hab, 1 = force on a (i), 2 = force on b (i), 3 = virial a (i,j),
 *       4 = virial b (i,j); all for compute_tau = false */
static double plain(const pctx *p, const int what, const int i, const int j,
                    const orb a, const orb b) {
  const double za = p->zeta, zb = p->zetb;
  const double *rab = p->rab;
  switch (what) {
  case 0:
    return C(p, a, b);
  case 1:
    return 2.0 * za * C(p, oup(i, a), b) - a.l[i] * C(p, odn(i, a), b);
  case 2:
    return 2.0 * zb * (C(p, oup(i, a), b) - rab[i] * C(p, a, b)) -
           b.l[i] * C(p, a, odn(i, b));
  case 3:
    return 2.0 * za * C(p, oup(i, oup(j, a)), b) -
           a.l[j] * C(p, oup(i, odn(j, a)), b);
  case 4:
    return 2.0 * zb *
               (C(p, oup(i, oup(j, a)), b) - C(p, oup(i, a), b) * rab[j] -
                C(p, oup(j, a), b) * rab[i] + C(p, a, b) * rab[j] * rab[i]) -
           b.l[j] * C(p, a, oup(i, odn(j, b)));
  }
  return 0.0;
}
/* kinetic-energy-density variant: 0.5 * grad a . grad b applied to `plain` */
static double elem(const pctx *p, const bool tau, const int what, const int i,
                   const int j, const orb a, const orb b) {
  if (!tau)
    return plain(p, what, i, j, a, b);
  double s = 0.0;
  for (int k = 0; k < 3; k++) {
    s += 0.5 * a.l[k] * b.l[k] * plain(p, what, i, j, odn(k, a), odn(k, b));
    s -= p->zeta * b.l[k] * plain(p, what, i, j, oup(k, a), odn(k, b));
    s -= a.l[k] * p->zetb * plain(p, what, i, j, odn(k, a), oup(k, b));
    s += 2.0 * p->zeta * p->zetb * plain(p, what, i, j, oup(k, a), oup(k, b));
  }
  return s;
}

/* ref/grid_ref_integrate.c:44-152 (without the hdab/hadb/a_hdab outputs,
 * which the batched API never requests). */
void grid_oracle_integrate_pgf_product(
    bool ortho, bool compute_tau, int border_mask, int la_max, int la_min,
    int lb_max, int lb_min, double zeta, double zetb, const double *dh,
    const double *dh_inv, const double *ra, const double *rab,
    const int *npts_global, const int *npts_local, const int *shift_local,
    const int *border_width, double radius, int o1, int o2, int n1, int n2,
    const double *grid, double *hab, const double *pab, double *forces,
    double *virials) {
  (void)n2;
  const bool do_f = (forces != NULL), do_v = (virials != NULL);
  assert(!do_v || do_f);
  int d_amax = 0, d_amin = 0, d_bmax = 0, d_bmin = 0;
  if (do_f || do_v)
    d_amax += 1, d_amin -= 1, d_bmin -= 1;
  if (do_v)
    d_amax += 1, d_bmax += 1;
  if (compute_tau)
    d_amax += 1, d_bmax += 1, d_amin -= 1, d_bmin -= 1;
  const int la_max_l = la_max + d_amax, lb_max_l = lb_max + d_bmax;
  const int la_min_l = imax(0, la_min + d_amin),
            lb_min_l = imax(0, lb_min + d_bmin);
  const int m1 = ncoset(la_max_l), m2 = ncoset(lb_max_l);
  double *cab = calloc((size_t)m1 * m2, sizeof(double));

  cab_grid(-1, ortho, border_mask, la_max_l, la_min_l, lb_max_l, lb_min_l, zeta,
           zetb, 1.0, dh, dh_inv, ra, rab, npts_global, npts_local, shift_local,
           border_width, radius, cab, (double *)grid);

  const pctx p = {cab, m1, zeta, zetb, rab};
  for (int la = la_min; la <= la_max; la++)
    for (int ax = 0; ax <= la; ax++)
      for (int ay = 0; ay <= la - ax; ay++) {
        const orb a = {{ax, ay, la - ax - ay}};
        for (int lb = lb_min; lb <= lb_max; lb++)
          for (int bx = 0; bx <= lb; bx++)
            for (int by = 0; by <= lb - bx; by++) {
              const orb b = {{bx, by, lb - bx - by}};
              const int ix = (o2 + oidx(b)) * n1 + o1 + oidx(a);
              hab[ix] += elem(&p, compute_tau, 0, 0, 0, a, b);
              if (do_f) {
                const double pv = pab[ix];
                for (int i = 0; i < 3; i++) {
                  forces[0 * 3 + i] += pv * elem(&p, compute_tau, 1, i, 0, a, b);
                  forces[1 * 3 + i] += pv * elem(&p, compute_tau, 2, i, 0, a, b);
                }
              }
              if (do_v) {
                const double pv = pab[ix];
                for (int i = 0; i < 3; i++)
                  for (int j = 0; j < 3; j++) {
                    virials[(0 * 3 + i) * 3 + j] +=
                        pv * elem(&p, compute_tau, 3, i, j, a, b);
                    virials[(1 * 3 + i) * 3 + j] +=
                        pv * elem(&p, compute_tau, 4, i, j, a, b);
                  }
              }
            }
      }
  free(cab);
}

/* ---------------------------------------------------------------------------
 * Task lists, restating ref/grid_ref_task_list.c.  Tasks are processed in the
 * order given (no sorting is needed for a serial accumulation).
 * ------------------------------------------------------------------------- */
typedef struct {
  int level, iatom, jatom, iset, jset, ipgf, jpgf, border_mask, block_num;
  double radius, rab[3];
} otask;
typedef struct {
  int npts_global[3], npts_local[3], shift_local[3], border_width[3];
  double dh[9], dh_inv[9];
} olayout;
typedef struct {
  bool ortho;
  int ntasks, nlevels, natoms, nkinds, nblocks;
  int *block_offsets, *atom_kinds;
  double *atom_positions;
  const oracle_basis_set **basis_sets;
  otask *tasks;
  olayout *layouts;
  int maxco;
} olist;

void grid_oracle_create_task_list(
    bool ortho, int ntasks, int nlevels, int natoms, int nkinds, int nblocks,
    const int *block_offsets, const double *atom_positions,
    const int *atom_kinds, const oracle_basis_set **basis_sets,
    const int *level_list, const int *iatom_list, const int *jatom_list,
    const int *iset_list, const int *jset_list, const int *ipgf_list,
    const int *jpgf_list, const int *border_mask_list,
    const int *block_num_list, const double *radius_list,
    const double *rab_list, const int *npts_global, const int *npts_local,
    const int *shift_local, const int *border_width, const double *dh,
    const double *dh_inv, void **out) {
  if (*out != NULL)
    grid_oracle_free_task_list(*out);
  olist *t = calloc(1, sizeof(olist));
  t->ortho = ortho;
  t->ntasks = ntasks, t->nlevels = nlevels, t->natoms = natoms;
  t->nkinds = nkinds, t->nblocks = nblocks;
  t->block_offsets = malloc(sizeof(int) * imax(nblocks, 1));
  memcpy(t->block_offsets, block_offsets, sizeof(int) * nblocks);
  t->atom_kinds = malloc(sizeof(int) * imax(natoms, 1));
  memcpy(t->atom_kinds, atom_kinds, sizeof(int) * natoms);
  t->atom_positions = malloc(sizeof(double) * 3 * imax(natoms, 1));
  memcpy(t->atom_positions, atom_positions, sizeof(double) * 3 * natoms);
  t->basis_sets = malloc(sizeof(void *) * imax(nkinds, 1));
  memcpy(t->basis_sets, basis_sets, sizeof(void *) * nkinds);
  t->tasks = malloc(sizeof(otask) * imax(ntasks, 1));
  for (int i = 0; i < ntasks; i++) {
    otask *k = &t->tasks[i];
    k->level = level_list[i], k->iatom = iatom_list[i];
    k->jatom = jatom_list[i], k->iset = iset_list[i], k->jset = jset_list[i];
    k->ipgf = ipgf_list[i], k->jpgf = jpgf_list[i];
    k->border_mask = border_mask_list[i], k->block_num = block_num_list[i];
    k->radius = radius_list[i];
    memcpy(k->rab, &rab_list[3 * i], sizeof(double) * 3);
  }
  t->layouts = malloc(sizeof(olayout) * imax(nlevels, 1));
  for (int l = 0; l < nlevels; l++) {
    olayout *y = &t->layouts[l];
    memcpy(y->npts_global, &npts_global[3 * l], sizeof(int) * 3);
    memcpy(y->npts_local, &npts_local[3 * l], sizeof(int) * 3);
    memcpy(y->shift_local, &shift_local[3 * l], sizeof(int) * 3);
    memcpy(y->border_width, &border_width[3 * l], sizeof(int) * 3);
    memcpy(y->dh, &dh[9 * l], sizeof(double) * 9);
    memcpy(y->dh_inv, &dh_inv[9 * l], sizeof(double) * 9);
  }
  t->maxco = 1;
  for (int i = 0; i < nkinds; i++)
    t->maxco = imax(t->maxco, basis_sets[i]->maxco);
  *out = t;
}

void grid_oracle_free_task_list(void *p) {
  if (p == NULL)
    return;
  olist *t = p;
  free(t->block_offsets);
  free(t->atom_kinds);
  free(t->atom_positions);
  free(t->basis_sets);
  free(t->tasks);
  free(t->layouts);
  free(t);
}

/* Everything needed to address one task's Cartesian sub-block. */
typedef struct {
  const oracle_basis_set *ib, *jb;
  int iatom, jatom, iset, jset, ncoseta, ncosetb, ncoa, ncob;
  int nsgf_seta, nsgf_setb, nsgfa, nsgfb, sgfa, sgfb;
  double zeta, zetb;
  bool transpose;
} tinfo;

static tinfo task_info(const olist *t, const otask *k) {
  tinfo s;
  s.iatom = k->iatom - 1, s.jatom = k->jatom - 1;
  s.iset = k->iset - 1, s.jset = k->jset - 1;
  s.ib = t->basis_sets[t->atom_kinds[s.iatom] - 1];
  s.jb = t->basis_sets[t->atom_kinds[s.jatom] - 1];
  s.zeta = s.ib->zet[s.iset * s.ib->maxpgf + k->ipgf - 1];
  s.zetb = s.jb->zet[s.jset * s.jb->maxpgf + k->jpgf - 1];
  s.ncoseta = ncoset(s.ib->lmax[s.iset]);
  s.ncosetb = ncoset(s.jb->lmax[s.jset]);
  s.ncoa = s.ib->npgf[s.iset] * s.ncoseta;
  s.ncob = s.jb->npgf[s.jset] * s.ncosetb;
  s.nsgf_seta = s.ib->nsgf_set[s.iset], s.nsgf_setb = s.jb->nsgf_set[s.jset];
  s.nsgfa = s.ib->nsgf, s.nsgfb = s.jb->nsgf;
  s.sgfa = s.ib->first_sgf[s.iset] - 1, s.sgfb = s.jb->first_sgf[s.jset] - 1;
  s.transpose = (s.iatom <= s.jatom);
  return s;
}

/* element (sgf_a, sgf_b) of the spherical block, honouring the storage rule of
 * ref/grid_ref_task_list.c:257-267 */
static inline size_t blk_index(const tinfo *s, const int ia, const int jb) {
  return s->transpose ? (size_t)(s->sgfb + jb) * s->nsgfa + s->sgfa + ia
                      : (size_t)(s->sgfa + ia) * s->nsgfb + s->sgfb + jb;
}

/* pab[ncob][ncoa] = sphi_b^T * block * sphi_a  (ref/grid_ref_task_list.c:237-271) */
static void decontract(const tinfo *s, const double *block, double *pab) {
  for (int jc = 0; jc < s->ncob; jc++)
    for (int ic = 0; ic < s->ncoa; ic++) {
      double acc = 0.0;
      for (int jb = 0; jb < s->nsgf_setb; jb++) {
        double w = 0.0;
        for (int ia = 0; ia < s->nsgf_seta; ia++)
          w += block[blk_index(s, ia, jb)] *
               s->ib->sphi[(s->sgfa + ia) * s->ib->maxco + ic];
        acc += s->jb->sphi[(s->sgfb + jb) * s->jb->maxco + jc] * w;
      }
      pab[jc * s->ncoa + ic] = acc;
    }
}

/* block += sphi_a * hab * sphi_b^T  (ref/grid_ref_task_list.c:462-499) */
static void contract(const tinfo *s, const double *hab, double *block) {
  for (int jb = 0; jb < s->nsgf_setb; jb++)
    for (int ia = 0; ia < s->nsgf_seta; ia++) {
      double acc = 0.0;
      for (int jc = 0; jc < s->ncob; jc++) {
        double w = 0.0;
        for (int ic = 0; ic < s->ncoa; ic++)
          w += hab[jc * s->ncoa + ic] *
               s->ib->sphi[(s->sgfa + ia) * s->ib->maxco + ic];
        acc += s->jb->sphi[(s->sgfb + jb) * s->jb->maxco + jc] * w;
      }
      block[blk_index(s, ia, jb)] += acc;
    }
}

/* ref/grid_ref_task_list.c:277-456; the grids are overwritten (:418-424). */
void grid_oracle_collocate_task_list(const void *p, int func, int nlevels,
                                     const oracle_buffer *pab_blocks,
                                     oracle_buffer **grids) {
  const olist *t = p;
  assert(t->nlevels == nlevels);
  for (int l = 0; l < nlevels; l++) {
    const olayout *y = &t->layouts[l];
    memset(grids[l]->host_buffer, 0,
           sizeof(double) * y->npts_local[0] * y->npts_local[1] *
               y->npts_local[2]);
  }
  double *pab = malloc(sizeof(double) * t->maxco * t->maxco);
  for (int it = 0; it < t->ntasks; it++) {
    const otask *k = &t->tasks[it];
    const tinfo s = task_info(t, k);
    const olayout *y = &t->layouts[k->level - 1];
    decontract(&s, &pab_blocks->host_buffer[t->block_offsets[k->block_num - 1]],
               pab);
    grid_oracle_collocate_pgf_product(
        t->ortho, k->border_mask, func, s.ib->lmax[s.iset], s.ib->lmin[s.iset],
        s.jb->lmax[s.jset], s.jb->lmin[s.jset], s.zeta, s.zetb,
        (s.iatom == s.jatom) ? 1.0 : 2.0, y->dh, y->dh_inv,
        &t->atom_positions[3 * s.iatom], k->rab, y->npts_global, y->npts_local,
        y->shift_local, y->border_width, k->radius, (k->ipgf - 1) * s.ncoseta,
        (k->jpgf - 1) * s.ncosetb, s.ncoa, s.ncob, pab,
        grids[k->level - 1]->host_buffer);
  }
  free(pab);
}

/* ref/grid_ref_task_list.c:505-691; outputs are overwritten (:672-678),
 * per-pair forces are scaled by 1 (same atom) or 2 (:625-641). */
void grid_oracle_integrate_task_list(const void *p, bool compute_tau,
                                     int natoms, int nlevels,
                                     const oracle_buffer *pab_blocks,
                                     const oracle_buffer **grids,
                                     oracle_buffer *hab_blocks, double *forces,
                                     double *virial) {
  const olist *t = p;
  assert(t->nlevels == nlevels && t->natoms == natoms);
  memset(hab_blocks->host_buffer, 0, hab_blocks->size);
  if (forces != NULL)
    memset(forces, 0, sizeof(double) * 3 * natoms);
  if (virial != NULL)
    memset(virial, 0, sizeof(double) * 9);
  const bool need_pab = (forces != NULL || virial != NULL);
  double *pab = malloc(sizeof(double) * t->maxco * t->maxco);
  double *hab = malloc(sizeof(double) * t->maxco * t->maxco);
  for (int it = 0; it < t->ntasks; it++) {
    const otask *k = &t->tasks[it];
    const tinfo s = task_info(t, k);
    const olayout *y = &t->layouts[k->level - 1];
    const int off = t->block_offsets[k->block_num - 1];
    if (need_pab)
      decontract(&s, &pab_blocks->host_buffer[off], pab);
    memset(hab, 0, sizeof(double) * s.ncoa * s.ncob);
    double f[6] = {0}, v[18] = {0};
    grid_oracle_integrate_pgf_product(
        t->ortho, compute_tau, k->border_mask, s.ib->lmax[s.iset],
        s.ib->lmin[s.iset], s.jb->lmax[s.jset], s.jb->lmin[s.jset], s.zeta,
        s.zetb, y->dh, y->dh_inv, &t->atom_positions[3 * s.iatom], k->rab,
        y->npts_global, y->npts_local, y->shift_local, y->border_width,
        k->radius, (k->ipgf - 1) * s.ncoseta, (k->jpgf - 1) * s.ncosetb, s.ncoa,
        s.ncob, grids[k->level - 1]->host_buffer, hab, need_pab ? pab : NULL,
        (forces != NULL) ? f : NULL, (virial != NULL) ? v : NULL);
    contract(&s, hab, &hab_blocks->host_buffer[off]);
    const double scale = (s.iatom == s.jatom) ? 1.0 : 2.0;
    if (forces != NULL)
      for (int i = 0; i < 3; i++) {
        forces[3 * s.iatom + i] += scale * f[i];
        forces[3 * s.jatom + i] += scale * f[3 + i];
      }
    if (virial != NULL)
      for (int i = 0; i < 9; i++)
        virial[i] += scale * (v[i] + v[9 + i]);
  }
  free(hab);
  free(pab);
}
