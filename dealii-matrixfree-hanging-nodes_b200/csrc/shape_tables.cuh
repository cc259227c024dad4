// 1D shape tables in __constant__ memory.  Every access in the kernels uses
// compile-time indices (fully unrolled loops over the template degree), so the
// entries become constant-bank operands of the FMA instructions and cost
// neither registers nor load instructions.
//
// The tables are `static`: every translation unit that defines kernels owns a
// private copy and uploads it once per device (ensure_shape_tables) before its
// first launch.  This keeps the kernel families in separately compiled objects
// (parallel builds) without relocatable device code.
#pragma once
#include "fe1d.hpp"

#include <cuda_runtime.h>

#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

namespace mfhn
{
enum TableId
{
  T_S  = 0, // S[q*n+i]   nodal -> Gauss values
  T_DC = 1, // Dc[q*n+p]  collocation derivative
  T_W0 = 2, // W0[i*n+j]  subface interpolation, lower half
  T_M  = 3, // 1D mass (nodal basis)
  T_K  = 4, // 1D stiffness (nodal basis)
  N_FULL = 5,
  // even-odd halves of the persymmetric matrices M and K ((n+1)/2 x (n+1)/2)
  T_ME = 0,
  T_MO = 1,
  T_KE = 2,
  T_KO = 3,
  // simultaneous diagonalisation M = T^T T, K = T^T diag(lambda) T (fe1d.hpp): the rows of T are symmetric
  // (first (n+1)/2 rows, table T_TE holds their left halves incl. the middle column) or antisymmetric
  // (remaining n/2 rows, T_TO holds their left halves); row stride (n+1)/2 in both tables
  T_TE = 4,
  T_TO = 5,
  N_EO = 6
};
constexpr int MAX_N  = 9;
constexpr int MAX_HE = 5;

template <typename Number>
struct ShapeTables
{
  Number full[8][N_FULL][MAX_N * MAX_N];
  Number qw[8][MAX_N];
  Number eo[8][N_EO][MAX_HE * MAX_HE];
  Number lam[8][MAX_N + 1]; // eigenvalues in the row order of T
};

static __constant__ ShapeTables<double> c_shape_d; // 36.2 KB
static __constant__ ShapeTables<float> c_shape_f;  // 18.1 KB

template <typename Number>
struct Shape;
template <>
struct Shape<double>
{
  template <int n, int T>
  static __device__ __forceinline__ double get(int idx)
  {
    return c_shape_d.full[n - 2][T][idx];
  }
  template <int n>
  static __device__ __forceinline__ double qw(int i)
  {
    return c_shape_d.qw[n - 2][i];
  }
  template <int n, int T>
  static __device__ __forceinline__ double eo(int idx)
  {
    return c_shape_d.eo[n - 2][T][idx];
  }
  template <int n>
  static __device__ __forceinline__ double lam(int i)
  {
    return c_shape_d.lam[n - 2][i];
  }
};
template <>
struct Shape<float>
{
  template <int n, int T>
  static __device__ __forceinline__ float get(int idx)
  {
    return c_shape_f.full[n - 2][T][idx];
  }
  template <int n>
  static __device__ __forceinline__ float qw(int i)
  {
    return c_shape_f.qw[n - 2][i];
  }
  template <int n, int T>
  static __device__ __forceinline__ float eo(int idx)
  {
    return c_shape_f.eo[n - 2][T][idx];
  }
  template <int n>
  static __device__ __forceinline__ float lam(int i)
  {
    return c_shape_f.lam[n - 2][i];
  }
};

// Host tables (computed once per process, shared by all translation units through the inline function).
inline const ShapeTables<double> &host_shape_tables()
{
  static const ShapeTables<double> tables = [] {
    ShapeTables<double> hd;
    std::memset(&hd, 0, sizeof(hd));
    for (int k = 1; k <= 8; ++k)
      {
        const Shape1D s = make_shape(k);
        const int n = k + 1, h = n / 2, he = (n + 1) / 2;
        auto &t = hd.full[k - 1];
        for (int i = 0; i < n * n; ++i)
          {
            t[T_S][i]  = s.S[i];
            t[T_DC][i] = s.Dc[i];
            t[T_W0][i] = s.W[0][i];
            t[T_M][i]  = s.M[i];
            t[T_K][i]  = s.K[i];
          }
        for (int i = 0; i < n; ++i)
          {
            hd.qw[k - 1][i]  = s.qw[i];
            hd.lam[k - 1][i] = s.lambda[i];
          }
        // even-odd halves of the persymmetric M and K: E = (A[i][j] + A[i][n-1-j]) / 2 (middle
        // column: A[i][m]), O = (A[i][j] - A[i][n-1-j]) / 2
        for (int which = 0; which < 2; ++which)
          {
            const std::vector<double> &A = which == 0 ? s.M : s.K;
            double *E = hd.eo[k - 1][which == 0 ? T_ME : T_KE], *O = hd.eo[k - 1][which == 0 ? T_MO : T_KO];
            for (int i = 0; i < he; ++i)
              for (int j = 0; j < he; ++j)
                {
                  if (j < h)
                    {
                      E[i * he + j] = 0.5 * (A[i * n + j] + A[i * n + (n - 1 - j)]);
                      O[i * he + j] = 0.5 * (A[i * n + j] - A[i * n + (n - 1 - j)]);
                    }
                  else
                    {
                      E[i * he + j] = A[i * n + j];
                      O[i * he + j] = 0;
                    }
                }
          }
        for (int i = 0; i < he; ++i)
          for (int j = 0; j < he; ++j) hd.eo[k - 1][T_TE][i * he + j] = s.T[i * n + j];
        for (int i = 0; i < h; ++i)
          for (int j = 0; j < h; ++j) hd.eo[k - 1][T_TO][i * he + j] = s.T[(he + i) * n + j];
      }
    return hd;
  }();
  return tables;
}

// Upload this translation unit's copy of the tables to `device` (once).
static inline void ensure_shape_tables(int device)
{
  static std::mutex mutex;
  static bool uploaded[64] = {};
  std::lock_guard<std::mutex> lock(mutex);
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
  if (uploaded[device]) return;
  const ShapeTables<double> &hd = host_shape_tables();
  static ShapeTables<float> hf;
  {
    const double *ps = reinterpret_cast<const double *>(&hd);
    float *pf        = reinterpret_cast<float *>(&hf);
    for (size_t i = 0; i < sizeof(hd) / sizeof(double); ++i) pf[i] = (float)ps[i];
  }
  cudaError_t e = cudaMemcpyToSymbol(c_shape_d, &hd, sizeof(hd));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_shape_f, &hf, sizeof(hf));
  if (e != cudaSuccess) throw std::runtime_error(std::string("shape table upload: ") + cudaGetErrorString(e));
  uploaded[device] = true;
}
} // namespace mfhn
