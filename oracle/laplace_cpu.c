/*
 * CPU restatement of the reference's vectorised vmult -- ORACLE / CPU BASELINE,
 * test infrastructure only (see oracle/__init__.py; parity unpinned by the
 * reference, pinned instead against oracle/operators.py O1/O2 in tests/).
 *
 * Follows Test::vmult in mode "CG (SC)" (/root/reference/benchmark_01.h:579-617):
 * for each batch of VL cells (deal.II: VectorizedArray<double>, 8 lanes with
 * AVX-512, benchmark_01.h:149)
 *   read_dof_values           gather + apply_hanging_node_constraints(false)  (:622)
 *   evaluate(gradients)       basis change to Gauss collocation + collocation derivative (:603)
 *   submit_gradient(get_gradient(q), q)   Cartesian: g_d *= w_q h              (:605-606)
 *   integrate(gradients)                                                       (:608)
 *   distribute_local_to_global   apply_hanging_node_constraints(true) + scatter-add (:652)
 * The hanging-node interpolation works lane by lane on the compressed mask
 * (the "index" strategy, HN_TYPE 0 of README.md:27).
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

typedef struct
{
  int n;
  const double *S, *Dc, *W0, *qw; /* n*n, n*n, n*n, n */
} shape_t;

/* in-place 1D contraction along direction dir of data[n^3][VL]:
 * out[i] = sum_j M[i][j] in[j]   (transpose: M[j][i]) */
static inline __attribute__((always_inline)) void sweep(const int n, const double *M, const int transpose, const int dir,
                                                        double *data)
{
  const int stride = dir == 0 ? 1 : dir == 1 ? n : n * n;
  for (int o1 = 0; o1 < n; ++o1)
    for (int o0 = 0; o0 < n; ++o0)
      {
        const int base = dir == 0 ? n * (o0 + n * o1) : dir == 1 ? o0 + n * n * o1 : o0 + n * o1;
        double in[MAXN][VL];
        for (int j = 0; j < n; ++j)
          for (int v = 0; v < VL; ++v) in[j][v] = data[(base + j * stride) * VL + v];
        for (int i = 0; i < n; ++i)
          {
            double acc[VL];
            for (int v = 0; v < VL; ++v) acc[v] = 0;
            for (int j = 0; j < n; ++j)
              {
                const double m = transpose ? M[j * n + i] : M[i * n + j];
                for (int v = 0; v < VL; ++v) acc[v] += m * in[j][v];
              }
            for (int v = 0; v < VL; ++v) data[(base + i * stride) * VL + v] = acc[v];
          }
      }
}

/* gradient part: r += Dc^T ( w * (Dc u) ) along dir, reading u, accumulating into r */
static inline __attribute__((always_inline)) void grad_dir(const int n, const shape_t *sh, const int dir, const double *u,
                                                           double *r, const double *hv)
{
  const int stride = dir == 0 ? 1 : dir == 1 ? n : n * n;
  for (int o1 = 0; o1 < n; ++o1)
    for (int o0 = 0; o0 < n; ++o0)
      {
        const int base  = dir == 0 ? n * (o0 + n * o1) : dir == 1 ? o0 + n * n * o1 : o0 + n * o1;
        const double wt = sh->qw[o0] * sh->qw[o1];
        double g[MAXN][VL];
        for (int q = 0; q < n; ++q)
          {
            for (int v = 0; v < VL; ++v) g[q][v] = 0;
            for (int j = 0; j < n; ++j)
              {
                const double d = sh->Dc[q * n + j];
                for (int v = 0; v < VL; ++v) g[q][v] += d * u[(base + j * stride) * VL + v];
              }
            const double w = wt * sh->qw[q];
            for (int v = 0; v < VL; ++v) g[q][v] *= w * hv[v]; /* submit_gradient: JxW J^-1 J^-T = w_q h */
          }
        for (int i = 0; i < n; ++i)
          for (int q = 0; q < n; ++q)
            {
              const double d = sh->Dc[q * n + i];
              for (int v = 0; v < VL; ++v) r[(base + i * stride) * VL + v] += d * g[q][v];
            }
      }
}

/* hanging-node interpolation of one lane (compressed mask), three directional passes */
static void hn_lane(const int n, const double *W0, const int transpose, const unsigned mask, double *data, const int lane)
{
  const int k          = n - 1;
  const unsigned v     = mask >> 5;
  const unsigned face  = (mask & 8u) ? v : 0u;
  const unsigned edge  = (mask & 16u) ? v : 0u;
  const unsigned child = (~mask) & 7u; /* child bit = 1 - subcell bit */
  for (int d = 0; d < 3; ++d)
    {
      const int t0 = d == 0 ? 1 : 0, t1 = d == 2 ? 1 : 2;
      const int stride = d == 0 ? 1 : d == 1 ? n : n * n;
      const int upper  = (child >> d) & 1u;
      for (int b = 0; b < n; ++b)
        for (int a = 0; a < n; ++a)
          {
            const int on0 = a == (int)((child >> t0) & 1u) * k, on1 = b == (int)((child >> t1) & 1u) * k;
            const int sel = (((face >> t0) & 1u) && on0) || (((face >> t1) & 1u) && on1) || (((edge >> d) & 1u) && on0 && on1);
            if (!sel) continue;
            const int base = d == 0 ? n * (a + n * b) : d == 1 ? a + n * n * b : a + n * b;
            double in[MAXN], out[MAXN];
            for (int i = 0; i < n; ++i) in[i] = data[(base + (upper ? k - i : i) * stride) * VL + lane];
            for (int i = 0; i < n; ++i)
              {
                double s = 0;
                for (int j = 0; j < n; ++j) s += (transpose ? W0[j * n + i] : W0[i * n + j]) * in[j];
                out[i] = s;
              }
            for (int i = 0; i < n; ++i) data[(base + (upper ? k - i : i) * stride) * VL + lane] = out[i];
          }
    }
}

static inline __attribute__((always_inline)) void vmult_n(const int n, const shape_t *sh, const long n_cells,
                                                          const uint32_t *idx, const uint8_t *masks, const double *h,
                                                          const double *src, double *dst, const int apply_constraints)
{
  const int n3 = n * n * n;
  double u[MAXN * MAXN * MAXN * VL] __attribute__((aligned(64)));
  double r[MAXN * MAXN * MAXN * VL] __attribute__((aligned(64)));
  for (long c0 = 0; c0 < n_cells; c0 += VL)
    {
      const int nl = (n_cells - c0) < VL ? (int)(n_cells - c0) : VL;
      double hv[VL];
      int any_hn = 0;
      for (int v = 0; v < VL; ++v)
        {
          hv[v] = v < nl ? h[c0 + v] : 0.0;
          if (v < nl && masks[c0 + v]) any_hn = 1;
        }
      for (int i = 0; i < n3; ++i)
        for (int v = 0; v < VL; ++v) u[i * VL + v] = v < nl ? src[idx[(c0 + v) * n3 + i]] : 0.0;
      if (apply_constraints && any_hn)
        for (int v = 0; v < nl; ++v)
          if (masks[c0 + v]) hn_lane(n, sh->W0, 0, masks[c0 + v], u, v);
      sweep(n, sh->S, 0, 0, u);
      sweep(n, sh->S, 0, 1, u);
      sweep(n, sh->S, 0, 2, u);
      memset(r, 0, sizeof(double) * n3 * VL);
      grad_dir(n, sh, 0, u, r, hv);
      grad_dir(n, sh, 1, u, r, hv);
      grad_dir(n, sh, 2, u, r, hv);
      sweep(n, sh->S, 1, 2, r);
      sweep(n, sh->S, 1, 1, r);
      sweep(n, sh->S, 1, 0, r);
      if (apply_constraints && any_hn)
        for (int v = 0; v < nl; ++v)
          if (masks[c0 + v]) hn_lane(n, sh->W0, 1, masks[c0 + v], r, v);
      for (int i = 0; i < n3; ++i)
        for (int v = 0; v < nl; ++v) dst[idx[(c0 + v) * n3 + i]] += r[i * VL + v];
    }
}

/* dst += A src, one thread */
int oracle_vmult(int degree, long n_cells, const uint32_t *idx, const uint8_t *masks, const double *h, const double *S,
                 const double *Dc, const double *W0, const double *qw, const double *src, double *dst, int apply_constraints)
{
  shape_t sh = {degree + 1, S, Dc, W0, qw};
  switch (degree)
    {
#define CASE(K)                                                                         \
  case K:                                                                               \
    vmult_n(K + 1, &sh, n_cells, idx, masks, h, src, dst, apply_constraints);         \
    break;
      CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
      default:
        return 1;
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
    oracle_vmult(degree, n_cells, idx, masks, h, S, Dc, W0, qw, src, dst, apply_constraints); /* warm-up */
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
