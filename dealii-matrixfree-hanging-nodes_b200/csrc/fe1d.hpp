// 1D shape data of FE_Q(k) with QGauss(k+1) on [0,1]: what deal.II's ShapeInfo
// hands to the reference's evaluators (FE_Q / QGauss built at
// benchmark_01.h:244-245, benchmark_03.h:434-435).  Computed in long double,
// rounded once.
#pragma once
#include <cmath>
#include <vector>

namespace mfhn
{
struct Shape1D
{
  int n = 0;                       // k+1
  std::vector<double> nodes;       // Gauss-Lobatto support points
  std::vector<double> qpts, qw;    // Gauss points / weights
  std::vector<double> S;           // S[q*n+i]  = l_i(qpts[q])
  std::vector<double> G;           // G[q*n+i]  = l_i'(qpts[q])
  std::vector<double> Dc;          // Dc[q*n+p] = collocation derivative at Gauss points
  std::vector<double> W[2];        // W[s][i*n+j] = l_j((x_i+s)/2)
  std::vector<double> M, K;        // 1D mass / stiffness in the nodal basis (n x n)
};

namespace detail
{
using ld = long double;

inline void legendre(int n, ld x, ld &p, ld &dp)
{
  ld p0 = 1, p1 = x;
  if (n == 0) { p = 1; dp = 0; return; }
  for (int m = 2; m <= n; ++m)
    {
      ld t = ((2 * m - 1) * x * p1 - (m - 1) * p0) / m;
      p0 = p1;
      p1 = t;
    }
  p  = p1;
  dp = n * (x * p1 - p0) / (x * x - 1);
}

inline void lagrange(const std::vector<ld> &nodes, ld x, std::vector<ld> &v, std::vector<ld> &d)
{
  const int n = nodes.size();
  v.assign(n, 0);
  d.assign(n, 0);
  for (int j = 0; j < n; ++j)
    {
      ld denom = 1;
      for (int m = 0; m < n; ++m)
        if (m != j) denom *= nodes[j] - nodes[m];
      ld val = 1;
      for (int m = 0; m < n; ++m)
        if (m != j) val *= x - nodes[m];
      ld der = 0;
      for (int m = 0; m < n; ++m)
        if (m != j)
          {
            ld t = 1;
            for (int r = 0; r < n; ++r)
              if (r != j && r != m) t *= x - nodes[r];
            der += t;
          }
      v[j] = val / denom;
      d[j] = der / denom;
    }
}
} // namespace detail

inline Shape1D make_shape(int degree)
{
  using detail::ld;
  const ld pi = acosl(-1.0L);
  const int n = degree + 1;
  std::vector<ld> gp(n), gw(n), gl(n);
  for (int i = 0; i < n; ++i)
    {
      ld x = -cosl(pi * (i + 0.75L) / (n + 0.5L));
      for (int it = 0; it < 100; ++it)
        {
          ld p, dp;
          detail::legendre(n, x, p, dp);
          ld dx = p / dp;
          x -= dx;
          if (fabsl(dx) < 1e-19L) break;
        }
      ld p, dp;
      detail::legendre(n, x, p, dp);
      gp[i] = x;
      gw[i] = 2 / ((1 - x * x) * dp * dp);
    }
  gl[0] = -1;
  gl[n - 1] = 1;
  const int m = n - 1;
  for (int j = 1; j < m; ++j)
    {
      ld x = -cosl(pi * j / m);
      for (int it = 0; it < 100; ++it)
        {
          ld p, dp;
          detail::legendre(m, x, p, dp);
          ld d2p = (2 * x * dp - m * (m + 1) * p) / (1 - x * x);
          ld dx  = dp / d2p;
          x -= dx;
          if (fabsl(dx) < 1e-19L) break;
        }
      gl[j] = x;
    }
  // map to [0,1] and symmetrise
  std::vector<ld> q(n), w(n), x(n);
  for (int i = 0; i < n; ++i)
    {
      q[i] = (gp[i] + 1) / 2;
      w[i] = gw[i] / 2;
      x[i] = (gl[i] + 1) / 2;
    }
  for (int i = 0; i < n; ++i)
    {
      ld a = (q[i] + (1 - q[n - 1 - i])) / 2;
      ld b = (w[i] + w[n - 1 - i]) / 2;
      ld c = (x[i] + (1 - x[n - 1 - i])) / 2;
      gp[i] = a;
      gw[i] = b;
      gl[i] = c;
    }
  q = gp;
  w = gw;
  x = gl;

  Shape1D s;
  s.n = n;
  s.nodes.resize(n);
  s.qpts.resize(n);
  s.qw.resize(n);
  s.S.resize(n * n);
  s.G.resize(n * n);
  s.Dc.resize(n * n);
  s.W[0].resize(n * n);
  s.W[1].resize(n * n);
  s.M.resize(n * n);
  s.K.resize(n * n);
  std::vector<ld> v, d, Sl(n * n), Gl(n * n);
  for (int i = 0; i < n; ++i)
    {
      s.nodes[i] = (double)x[i];
      s.qpts[i]  = (double)q[i];
      s.qw[i]    = (double)w[i];
    }
  for (int qq = 0; qq < n; ++qq)
    {
      detail::lagrange(x, q[qq], v, d);
      for (int i = 0; i < n; ++i)
        {
          Sl[qq * n + i] = v[i];
          Gl[qq * n + i] = d[i];
          s.S[qq * n + i] = (double)v[i];
          s.G[qq * n + i] = (double)d[i];
        }
      detail::lagrange(q, q[qq], v, d);
      for (int p = 0; p < n; ++p) s.Dc[qq * n + p] = (double)d[p];
    }
  for (int sub = 0; sub < 2; ++sub)
    for (int i = 0; i < n; ++i)
      {
        detail::lagrange(x, (x[i] + sub) / 2, v, d);
        for (int j = 0; j < n; ++j) s.W[sub][i * n + j] = (double)v[j];
      }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      {
        ld mm = 0, kk = 0;
        for (int qq = 0; qq < n; ++qq)
          {
            mm += w[qq] * Sl[qq * n + i] * Sl[qq * n + j];
            kk += w[qq] * Gl[qq * n + i] * Gl[qq * n + j];
          }
        s.M[i * n + j] = (double)mm;
        s.K[i * n + j] = (double)kk;
      }
  return s;
}
} // namespace mfhn
