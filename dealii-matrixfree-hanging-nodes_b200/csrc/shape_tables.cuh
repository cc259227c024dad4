// 1D shape tables in __constant__ memory.  Every access in the kernels uses
// compile-time indices (fully unrolled loops over the template degree), so the
// entries become constant-bank operands of the FMA instructions and cost
// neither registers nor load instructions.
#pragma once

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
  N_EO = 4
};
constexpr int MAX_N  = 9;
constexpr int MAX_HE = 5;

template <typename Number>
struct ShapeTables
{
  Number full[8][N_FULL][MAX_N * MAX_N];
  Number qw[8][MAX_N];
  Number eo[8][N_EO][MAX_HE * MAX_HE];
};

// defined here: this header is included by exactly one translation unit (op.cu)
__constant__ ShapeTables<double> c_shape_d; // 31.7 KB
__constant__ ShapeTables<float> c_shape_f;  // 15.9 KB

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
};
} // namespace mfhn
