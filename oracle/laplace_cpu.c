/*
 * CPU restatement of the reference's vectorised vmult -- ORACLE / CPU BASELINE,
 * test infrastructure only (see oracle/__init__.py; parity unpinned by the
 * reference, pinned instead against oracle/operators.py O1/O2 in tests/).
 *
 * Follows Test::vmult in mode "CG (SC)" (/root/reference/benchmark_01.h:579-617):
 * for each batch of 8 cells (deal.II: VectorizedArray<double>, 8 lanes with
 * AVX-512, benchmark_01.h:149)
 *   read_dof_values           gather + apply_hanging_node_constraints(false)  (:622)
 *   evaluate(gradients)       basis change to Gauss collocation + collocation derivative,
 *                             even-odd sum factorisation                       (:603)
 *   submit_gradient(get_gradient(q), q)   Cartesian: g_d *= w_q h              (:605-606)
 *   integrate(gradients)                                                       (:608)
 *   distribute_local_to_global   apply_hanging_node_constraints(true) + scatter-add (:652)
 * The hanging-node interpolation works lane by lane on the compressed mask
 * (the "index" strategy, HN_TYPE 0 of README.md:27).  SIMD across the cells of
 * a batch is written with GCC vector extensions (one 512-bit vector = the 8
 * lanes of a VectorizedArray).
 *
 * Throughput semantics of benchmark_01 (benchmark_01.h:536-573): every
 * thread applies the operator to its own copy of the vectors.
 */
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define VL 8
#define MAXN 9
typedef double v8d __attribute__((vector_size(64), aligned(64)));

typedef struct
{
  int n;
  const double *S, *Dc, *W0, *qw; /* n*n, n*n, n*n, n */
  /* even-odd halves: [matrix: S, S^T, Dc, Dc^T][E | O][(n+1)/2 * (n+1)/2] */
  double eo[4][2][25];
} shape_t;

/* E[i][j] = (A[i][j] + A[i][n-1-j]) / 2 (middle column: A[i][m]), O[i][j] = (A[i][j] - A[i][n-1-j]) / 2 */
static void make_eo(int n, const double *A, int transpose, double *E, double *O)
{
  const int h = n / 2, he = (n + 1) / 2;
  for (int i = 0; i < he; ++i)
    for (int j = 0; j < he; ++j)
      {
        const double a = transpose ? A[j * n + i] : A[i * n + j];
        const double b = transpose ? A[(n - 1 - j) * n + i] : A[i * n + (n - 1 - j)];
        if (j < h)
          {
            E[i * he + j] = 0.5 * (a + b);
            O[i * he + j] = 0.5 * (a - b);
          }
        else
          {
            E[i * he + j] = a;
            O[i * he + j] = 0;
          }
      }
}

/* y = A x for a matrix with A[n-1-i][n-1-j] = sign * A[i][j], even-odd form */
static inline __attribute__((always_inline)) void eo_apply(const int n, const double *E, const double *O, const int sign,
                                                           const v8d *x, v8d *y)
{
  const int h = n / 2, he = (n + 1) / 2;
  v8d xs[5], xd[5];
  for (int j = 0; j < h; ++j)
    {
      xs[j] = x[j] + x[n - 1 - j];
      xd[j] = x[j] - x[n - 1 - j];
    }
  if (n & 1) xs[h] = x[h];
  for (int i = 0; i < h; ++i)
    {
      v8d e = E[i * he] * xs[0], o = O[i * he] * xd[0];
      for (int j = 1; j < he; ++j) e += E[i * he + j] * xs[j];
      for (int j = 1; j < h; ++j) o += O[i * he + j] * xd[j];
      y[i]         = e + o;
      y[n - 1 - i] = sign > 0 ? e - o : o - e;
    }
  if (n & 1)
    {
      if (sign > 0)
        {
          v8d e = E[h * he] * xs[0];
          for (int j = 1; j < he; ++j) e += E[h * he + j] * xs[j];
          y[h] = e;
        }
      else
        {
          v8d o = O[h * he] * xd[0];
          for (int j = 1; j < h; ++j) o += O[h * he + j] * xd[j];
          y[h] = o;
        }
    }
}

static inline __attribute__((always_inline)) int line_base(const int n, const int dir, const int o0, const int o1)
{
  return dir == 0 ? n * (o0 + n * o1) : dir == 1 ? o0 + n * n * o1 : o0 + n * o1;
}

/* in-place 1D sweep along dir with matrix `which` (0: S, 1: S^T) */
static inline __attribute__((always_inline)) void sweep(const int n, const shape_t *sh, const int which, const int dir, v8d *data)
{
  const int stride = dir == 0 ? 1 : dir == 1 ? n : n * n;
  for (int o1 = 0; o1 < n; ++o1)
    for (int o0 = 0; o0 < n; ++o0)
      {
        const int base = line_base(n, dir, o0, o1);
        v8d in[MAXN], out[MAXN];
        for (int j = 0; j < n; ++j) in[j] = data[base + j * stride];
        eo_apply(n, sh->eo[which][0], sh->eo[which][1], +1, in, out);
        for (int i = 0; i < n; ++i) data[base + i * stride] = out[i];
      }
}

/* gradient part along dir: r += Dc^T ( w_q h (Dc u) ) */
static inline __attribute__((always_inline)) void grad_dir(const int n, const shape_t *sh, const int dir, const v8d *u, v8d *r,
                                                           const v8d hv)
{
  const int stride = dir == 0 ? 1 : dir == 1 ? n : n * n;
  for (int o1 = 0; o1 < n; ++o1)
    for (int o0 = 0; o0 < n; ++o0)
      {
        const int base = line_base(n, dir, o0, o1);
        const v8d wt   = hv * (sh->qw[o0] * sh->qw[o1]);
        v8d in[MAXN], g[MAXN], out[MAXN];
        for (int j = 0; j < n; ++j) in[j] = u[base + j * stride];
        eo_apply(n, sh->eo[2][0], sh->eo[2][1], -1, in, g);      /* get_gradient */
        for (int q = 0; q < n; ++q) g[q] *= wt * sh->qw[q];       /* submit_gradient: JxW J^-1 J^-T = w_q h */
        eo_apply(n, sh->eo[3][0], sh->eo[3][1], -1, g, out);     /* integrate */
        for (int i = 0; i < n; ++i) r[base + i * stride] += out[i];
      }
}

/* hanging-node interpolation of one lane (compressed mask), three directional passes */
static inline __attribute__((always_inline)) void hn_lane(const int n, const double *W0, const int transpose, const unsigned mask,
                                                          v8d *vdata, const int lane)
{
  double *data = (double *)vdata + lane; /* scalar view of this lane: element i sits at data[i * VL] */
  const int k          = n - 1;
  const unsigned v     = mask >> 5;
  const unsigned face  = (mask & 8u) ? v : 0u;
  const unsigned edge  = (mask & 16u) ? v : 0u;
  const unsigned child = (~mask) & 7u; /* child bit = 1 - subcell bit */
  for (int d = 0; d < 3; ++d)
    {
      const int t0 = d == 0 ? 1 : 0, t1 = d == 2 ? 1 : 2;
      const int f0 = (face >> t0) & 1u, f1 = (face >> t1) & 1u, ed = (edge >> d) & 1u;
      if (!(f0 || f1 || ed)) continue;
      const int c0 = (int)((child >> t0) & 1u) * k, c1 = (int)((child >> t1) & 1u) * k;
      const int stride = d == 0 ? 1 : d == 1 ? n : n * n;
      const int upper  = (child >> d) & 1u;
      for (int b = 0; b < n; ++b)
        for (int a = 0; a < n; ++a)
          {
            const int on0 = a == c0, on1 = b == c1;
            if (!((f0 && on0) || (f1 && on1) || (ed && on0 && on1))) continue;
            const int base = line_base(n, d, a, b);
            double in[MAXN], out[MAXN];
            for (int i = 0; i < n; ++i) in[i] = data[(base + (upper ? k - i : i) * stride) * VL];
            for (int i = 0; i < n; ++i)
              {
                double s = 0;
                for (int j = 0; j < n; ++j) s += (transpose ? W0[j * n + i] : W0[i * n + j]) * in[j];
                out[i] = s;
              }
            for (int i = 0; i < n; ++i) data[(base + (upper ? k - i : i) * stride) * VL] = out[i];
          }
    }
}

static inline __attribute__((always_inline)) void vmult_n(const int n, const shape_t *sh, const long n_cells,
                                                          const uint32_t *idx, const uint8_t *masks, const double *h,
                                                          const double *src, double *dst, const int apply_constraints)
{
  const int n3 = n * n * n;
  v8d u[MAXN * MAXN * MAXN], r[MAXN * MAXN * MAXN];
  for (long c0 = 0; c0 < n_cells; c0 += VL)
    {
      const int nl = (n_cells - c0) < VL ? (int)(n_cells - c0) : VL;
      v8d hv;
      int any_hn = 0;
      const uint32_t *ip[VL];
      for (int v = 0; v < VL; ++v)
        {
          const long c = c0 + (v < nl ? v : 0);
          hv[v]        = v < nl ? h[c] : 0.0;
          ip[v]        = idx + c * n3;
          if (v < nl && masks[c]) any_hn = 1;
        }
      for (int i = 0; i < n3; ++i)
        {
          v8d t;
          for (int v = 0; v < VL; ++v) t[v] = src[ip[v][i]];
          u[i] = t;
        }
      if (apply_constraints && any_hn)
        for (int v = 0; v < nl; ++v)
          if (masks[c0 + v]) hn_lane(n, sh->W0, 0, masks[c0 + v], u, v);
      sweep(n, sh, 0, 0, u);
      sweep(n, sh, 0, 1, u);
      sweep(n, sh, 0, 2, u);
      memset(r, 0, sizeof(v8d) * n3);
      grad_dir(n, sh, 0, u, r, hv);
      grad_dir(n, sh, 1, u, r, hv);
      grad_dir(n, sh, 2, u, r, hv);
      sweep(n, sh, 1, 2, r);
      sweep(n, sh, 1, 1, r);
      sweep(n, sh, 1, 0, r);
      if (apply_constraints && any_hn)
        for (int v = 0; v < nl; ++v)
          if (masks[c0 + v]) hn_lane(n, sh->W0, 1, masks[c0 + v], r, v);
      for (int i = 0; i < n3; ++i)
        for (int v = 0; v < nl; ++v) dst[ip[v][i]] += r[i][v];
    }
}

#define DEFINE_VMULT(N)                                                                                                         \
  static __attribute__((noinline)) void vmult_##N(const shape_t *sh, long n_cells, const uint32_t *idx, const uint8_t *masks,    \
                                                  const double *h, const double *src, double *dst, int ac)                      \
  {                                                                                                                             \
    vmult_n(N, sh, n_cells, idx, masks, h, src, dst, ac);                                                                       \
  }
DEFINE_VMULT(2)
DEFINE_VMULT(3)
DEFINE_VMULT(4)
DEFINE_VMULT(5)
DEFINE_VMULT(6)
DEFINE_VMULT(7)
DEFINE_VMULT(8)
DEFINE_VMULT(9)

/* dst += A src, one thread */
int oracle_vmult(int degree, long n_cells, const uint32_t *idx, const uint8_t *masks, const double *h, const double *S,
                 const double *Dc, const double *W0, const double *qw, const double *src, double *dst, int apply_constraints)
{
  if (degree < 1 || degree > 8) return 1;
  shape_t sh;
  sh.n  = degree + 1;
  sh.S  = S;
  sh.Dc = Dc;
  sh.W0 = W0;
  sh.qw = qw;
  make_eo(sh.n, S, 0, sh.eo[0][0], sh.eo[0][1]);
  make_eo(sh.n, S, 1, sh.eo[1][0], sh.eo[1][1]);
  make_eo(sh.n, Dc, 0, sh.eo[2][0], sh.eo[2][1]);
  make_eo(sh.n, Dc, 1, sh.eo[3][0], sh.eo[3][1]);
  switch (degree)
    {
#define CASE(K, N)                                                      \
  case K:                                                               \
    vmult_##N(&sh, n_cells, idx, masks, h, src, dst, apply_constraints); \
    break;
      CASE(1, 2) CASE(2, 3) CASE(3, 4) CASE(4, 5) CASE(5, 6) CASE(6, 7) CASE(7, 8) CASE(8, 9)
#undef CASE
    }
  return 0;
}

/* benchmark_01 timing loop: n_threads replicas, each applying the operator
 * n_rep times to its own vectors (src == 1.0, benchmark_01.h:510-511).  Returns
 * the mean time of one vmult over repetitions and threads (benchmark_01.h:571-572). */
double oracle_benchmark(int degree, long n_cells, long n_dofs, const uint32_t *idx, const uint8_t *masks, const double *h,
                        const double *S, const double *Dc, const double *W0, const double *qw, int apply_constraints,
                        int n_rep, int n_threads)
{
  double total = 0;
#pragma omp parallel num_threads(n_threads) reduction(+ : total)
  {
    double *src = (double *)malloc(sizeof(double) * n_dofs), *dst = (double *)calloc(n_dofs, sizeof(double));
    for (long i = 0; i < n_dofs; ++i) src[i] = 1.0;
    for (int w = 0; w < 2; ++w) /* warm-up: page faults, thread start-up */
      oracle_vmult(degree, n_cells, idx, masks, h, S, Dc, W0, qw, src, dst, apply_constraints);
    double mine = 0;
    for (int rep = 0; rep < n_rep; ++rep)
      {
#pragma omp barrier
        const double t0 = omp_get_wtime();
        oracle_vmult(degree, n_cells, idx, masks, h, S, Dc, W0, qw, src, dst, apply_constraints);
        mine += omp_get_wtime() - t0;
      }
    total += mine / n_rep;
    free(src);
    free(dst);
  }
  return total / n_threads;
}
