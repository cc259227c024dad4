// 1D shape data of FE_Q(k) with QGauss(k+1) on [0,1]: what deal.II's ShapeInfo
// hands to the reference's evaluators (FE_Q / QGauss built at
// benchmark_01.h:244-245, benchmark_03.h:434-435).  Computed in long double,
// rounded once.
#pragma once
#include <cmath>
#include <utility>
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
  // Simultaneous diagonalisation of the pencil (K, M): T = V^-1 with V^T M V = I, V^T K V = diag(lambda), i.e.
  // M = T^T T and K = T^T diag(lambda) T.  M and K are persymmetric, so every row of T is symmetric or
  // antisymmetric under i -> n-1-i; the (n+1)/2 symmetric rows come first, each group by ascending eigenvalue.
  std::vector<double> T;           // T[i*n+j]
  std::vector<double> lambda;      // lambda[i], ordered like the rows of T
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

// Generalised symmetric eigenproblem K v = lambda M v of the 1D stiffness / mass pair, solved in long
// double inside the symmetric and the antisymmetric subspace separately (both matrices are persymmetric, so
// the two subspaces decouple): Cholesky M = L L^T, cyclic Jacobi on L^-1 K L^-T, T = Q^T L^T.
inline void diagonalise(int n, const std::vector<ld> &S, const std::vector<ld> &G, const std::vector<ld> &w, std::vector<double> &Tout,
                        std::vector<double> &lam)
{
  std::vector<ld> M(n * n), K(n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      {
        ld mm = 0, kk = 0;
        for (int q = 0; q < n; ++q)
          {
            mm += w[q] * S[q * n + i] * S[q * n + j];
            kk += w[q] * G[q * n + i] * G[q * n + j];
          }
        M[i * n + j] = mm;
        K[i * n + j] = kk;
      }
  Tout.assign(n * n, 0.0);
  lam.assign(n, 0.0);
  const int h = n / 2, he = (n + 1) / 2;
  const ld r2 = sqrtl(0.5L);
  int row0 = 0;
  for (int parity = 0; parity < 2; ++parity)
    {
      const int m = parity == 0 ? he : h; // dimension of the subspace
      if (m == 0) continue;
      // orthonormal basis B[n x m]: (e_j +- e_{n-1-j}) / sqrt 2, and e_h for the middle entry of the symmetric part
      std::vector<ld> B(n * m, 0);
      for (int j = 0; j < m; ++j)
        {
          if (parity == 0 && (n % 2) && j == h)
            B[h * m + j] = 1;
          else
            {
              B[j * m + j]           = r2;
              B[(n - 1 - j) * m + j] = parity == 0 ? r2 : -r2;
            }
        }
      auto project = [&](const std::vector<ld> &A) {
        std::vector<ld> AB(n * m, 0), R(m * m, 0);
        for (int i = 0; i < n; ++i)
          for (int j = 0; j < m; ++j)
            for (int k = 0; k < n; ++k) AB[i * m + j] += A[i * n + k] * B[k * m + j];
        for (int i = 0; i < m; ++i)
          for (int j = 0; j < m; ++j)
            for (int k = 0; k < n; ++k) R[i * m + j] += B[k * m + i] * AB[k * m + j];
        return R;
      };
      std::vector<ld> Ms = project(M), Ks = project(K);
      // Cholesky Ms = L L^T
      std::vector<ld> L(m * m, 0);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j <= i; ++j)
          {
            ld s = Ms[i * m + j];
            for (int k = 0; k < j; ++k) s -= L[i * m + k] * L[j * m + k];
            L[i * m + j] = i == j ? sqrtl(s) : s / L[j * m + j];
          }
      // Li = L^-1 (lower triangular)
      std::vector<ld> Li(m * m, 0);
      for (int c = 0; c < m; ++c)
        for (int i = c; i < m; ++i)
          {
            ld s = i == c ? 1 : 0;
            for (int k = c; k < i; ++k) s -= L[i * m + k] * Li[k * m + c];
            Li[i * m + c] = s / L[i * m + i];
          }
      // A = Li Ks Li^T
      std::vector<ld> tmp(m * m, 0), A(m * m, 0);
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j)
          for (int k = 0; k < m; ++k) tmp[i * m + j] += Li[i * m + k] * Ks[k * m + j];
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j)
          for (int k = 0; k < m; ++k) A[i * m + j] += tmp[i * m + k] * Li[j * m + k];
      for (int i = 0; i < m; ++i)
        for (int j = 0; j < i; ++j) A[i * m + j] = A[j * m + i] = (A[i * m + j] + A[j * m + i]) / 2;
      // cyclic Jacobi: A = Q diag Q^T
      std::vector<ld> Q(m * m, 0);
      for (int i = 0; i < m; ++i) Q[i * m + i] = 1;
      for (int sweep = 0; sweep < 100; ++sweep)
        {
          ld off = 0, dia = 0;
          for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) (i == j ? dia : off) += A[i * m + j] * A[i * m + j];
          if (off <= 1e-40L * dia || off == 0) break;
          for (int p = 0; p < m; ++p)
            for (int q = p + 1; q < m; ++q)
              {
                if (A[p * m + q] == 0) continue;
                const ld theta = (A[q * m + q] - A[p * m + p]) / (2 * A[p * m + q]);
                const ld t     = (theta >= 0 ? 1 : -1) / (fabsl(theta) + sqrtl(theta * theta + 1));
                const ld c = 1 / sqrtl(t * t + 1), sn = t * c;
                for (int k = 0; k < m; ++k)
                  {
                    const ld akp = A[k * m + p], akq = A[k * m + q];
                    A[k * m + p] = c * akp - sn * akq;
                    A[k * m + q] = sn * akp + c * akq;
                  }
                for (int k = 0; k < m; ++k)
                  {
                    const ld apk = A[p * m + k], aqk = A[q * m + k];
                    A[p * m + k] = c * apk - sn * aqk;
                    A[q * m + k] = sn * apk + c * aqk;
                  }
                for (int k = 0; k < m; ++k)
                  {
                    const ld qkp = Q[k * m + p], qkq = Q[k * m + q];
                    Q[k * m + p] = c * qkp - sn * qkq;
                    Q[k * m + q] = sn * qkp + c * qkq;
                  }
              }
        }
      // ascending eigenvalues
      std::vector<int> order(m);
      for (int i = 0; i < m; ++i) order[i] = i;
      for (int i = 0; i < m; ++i)
        for (int j = i + 1; j < m; ++j)
          if (A[order[j] * m + order[j]] < A[order[i] * m + order[i]]) std::swap(order[i], order[j]);
      // rows of T in the subspace: Q^T L^T; back to nodal coordinates: (Q^T L^T) B^T
      for (int r = 0; r < m; ++r)
        {
          const int e = order[r];
          ld ev = A[e * m + e];
          if (fabsl(ev) < 1e-12L) ev = 0; // the constant mode (all other eigenvalues are > 9)
          lam[row0 + r] = (double)ev;
          for (int j = 0; j < n; ++j)
            {
              ld s = 0;
              for (int a = 0; a < m; ++a)
                {
                  ld qlt = 0; // (Q^T L^T)[e][a] = sum_b Q[b][e] L[a][b]
                  for (int b = 0; b < m; ++b) qlt += Q[b * m + e] * L[a * m + b];
                  s += qlt * B[j * m + a];
                }
              Tout[(row0 + r) * n + j] = (double)s;
            }
        }
      row0 += m;
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
  detail::diagonalise(n, Sl, Gl, w, s.T, s.lambda);
  return s;
}
} // namespace mfhn
