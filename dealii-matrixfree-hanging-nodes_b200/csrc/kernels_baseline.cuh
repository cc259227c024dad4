// "baseline-B": restatement of the design of deal.II's CUDAWrappers::MatrixFree
// path that the reference's cuda/benchmark_03.cu runs (benchmark_03.h:280-357;
// deal.II headers cuda_matrix_free.templates.h, cuda_fe_evaluation.h,
// cuda_hanging_nodes_internal.h, cuda_tensor_product_kernels.h -- not available
// here, restated from SURVEY.md 2b D7-D10 / 2d).  It is a REPORTED BASELINE
// measured on the same box, not the product path:
//   * one thread per DoF, (k+1)^3 threads per cell, cells_per_block = 8 / 2 / 1 for k = 1 / 2 / >= 3;
//   * static shared memory values[] + gradients[3][];
//   * local_to_global[cell * padding_length + i], padding_length = 2^ceil(3 log2(k+1));
//   * per-quadrature-point geometry streamed from global memory: full 3x3
//     inv_jacobian[(d*3+e) * n_cells * pad + cell * pad + q] and JxW[cell * pad + q];
//   * dense (non even-odd) sum factorisation with the shape matrices in
//     __constant__ memory and a block-wide barrier after every sweep;
//   * resolve_hanging_nodes: three direction passes, every thread decides from the
//     mask whether its DoF is constrained, two barriers per pass;
//   * scatter with atomicAdd (no colouring, the reference's default).
#pragma once
#include "kernels_generic.cuh"

#include <cuda_runtime.h>

namespace mfhn
{
struct BaselineParams
{
  const uint32_t *local_to_global; // [n_cells * pad]
  const void *inv_jacobian;        // Number[9 * n_cells * pad]
  const void *JxW;                 // Number[n_cells * pad]
  const uint8_t *masks;
  const void *src;
  void *dst;
  long long n_cells, cell_begin, cell_end;
  int pad, apply_constraints;
};

template <int n>
struct BaselineCfg
{
  static constexpr int n3  = n * n * n;
  static constexpr int cpb = n == 2 ? 8 : n == 3 ? 2 : 1; // cells_per_block_shmem of deal.II in 3D
  static constexpr int pad = n3 <= 8 ? 8 : n3 <= 32 ? 32 : n3 <= 64 ? 64 : n3 <= 128 ? 128 : n3 <= 256 ? 256 : n3 <= 512 ? 512 : 1024;
};

// one 1D sweep: out(x) = sum_k M[x][k] in(k) along `dir`, in place with two barriers (as deal.II's `apply<..., in_place>`)
template <int n, int T, bool transpose, bool add, typename Number>
__device__ __forceinline__ void baseline_sweep(const Number *in, Number *out, const int dir, const int x, const int y, const int z)
{
  const int stride = dir == 0 ? 1 : dir == 1 ? n : n * n;
  const int q      = dir == 0 ? x : dir == 1 ? y : z;
  const int base   = x + n * (y + n * z) - q * stride;
  Number t         = Number(0);
#pragma unroll
  for (int k = 0; k < n; ++k)
    {
      // runtime row index q: the shape matrix is read from constant memory like deal.II's global_shape_values
      const Number m = transpose ? Shape<Number>::template get<n, T>(k * n + q) : Shape<Number>::template get<n, T>(q * n + k);
      t += m * in[base + k * stride];
    }
  if (in == out) __syncthreads();
  if (add)
    out[base + q * stride] += t;
  else
    out[base + q * stride] = t;
  __syncthreads();
}

template <int n, bool transpose, typename Number>
__device__ __forceinline__ void baseline_resolve_hanging_nodes(Number *values, const unsigned mask, const int x, const int y, const int z)
{
  constexpr int k = n - 1;
  unsigned face, edge, cb;
  decode_mask(mask, face, edge, cb);
  const int a[3] = {x, y, z};
  for (int d = 0; d < 3; ++d)
    {
      const int t0 = (d == 0) ? 1 : 0, t1 = (d == 2) ? 1 : 2;
      const bool on0 = a[t0] == (int)((cb >> t0) & 1u) * k, on1 = a[t1] == (int)((cb >> t1) & 1u) * k;
      const bool constrained = mask != 0u && ((((face >> t0) & 1u) && on0) || (((face >> t1) & 1u) && on1) || (((edge >> d) & 1u) && on0 && on1));
      const bool upper  = (cb >> d) & 1u;
      const int stride  = d == 0 ? 1 : d == 1 ? n : n * n;
      const int base    = x + n * (y + n * z) - a[d] * stride;
      Number t          = Number(0);
      if (constrained)
        {
          const int i = upper ? k - a[d] : a[d];
          for (int j = 0; j < n; ++j)
            {
              const Number w = transpose ? Shape<Number>::template get<n, T_W0>(j * n + i) : Shape<Number>::template get<n, T_W0>(i * n + j);
              t += w * values[base + (upper ? k - j : j) * stride];
            }
        }
      __syncthreads();
      if (constrained) values[base + a[d] * stride] = t;
      __syncthreads();
    }
}

template <int n, typename Number>
__global__ void __launch_bounds__(BaselineCfg<n>::n3 *BaselineCfg<n>::cpb) baseline_kernel(const BaselineParams p)
{
  using Cfg = BaselineCfg<n>;
  constexpr int n3 = Cfg::n3;
  __shared__ Number s_values[Cfg::cpb * n3];
  __shared__ Number s_grad[3][Cfg::cpb * n3];
  const int cib = threadIdx.x / n3, i = threadIdx.x % n3;
  const int x = i % n, y = (i / n) % n, z = i / (n * n);
  const long long cell = p.cell_begin + (long long)blockIdx.x * Cfg::cpb + cib;
  const bool valid     = cell < p.cell_end;
  Number *values = s_values + cib * n3;
  Number *gx = s_grad[0] + cib * n3, *gy = s_grad[1] + cib * n3, *gz = s_grad[2] + cib * n3;
  const Number *src = static_cast<const Number *>(p.src);
  Number *dst       = static_cast<Number *>(p.dst);
  const long long slot = (valid ? cell : p.cell_begin) * p.pad + i;
  const uint32_t g     = p.local_to_global[slot];
  const unsigned mask  = (valid && p.apply_constraints) ? p.masks[cell] : 0u;

  values[i] = valid ? __ldg(src + g) : Number(0); // read_dof_values
  __syncthreads();
  if (p.apply_constraints) baseline_resolve_hanging_nodes<n, false>(values, mask, x, y, z);
  // evaluate(gradients): values to the quadrature points, then collocation gradients
  baseline_sweep<n, T_S, false, false>(values, values, 0, x, y, z);
  baseline_sweep<n, T_S, false, false>(values, values, 1, x, y, z);
  baseline_sweep<n, T_S, false, false>(values, values, 2, x, y, z);
  baseline_sweep<n, T_DC, false, false>(values, gx, 0, x, y, z);
  baseline_sweep<n, T_DC, false, false>(values, gy, 1, x, y, z);
  baseline_sweep<n, T_DC, false, false>(values, gz, 2, x, y, z);
  {
    // apply_for_each_quad_point: submit_gradient(get_gradient()) with the 3x3 inverse Jacobian and JxW of this q-point
    const Number *ij      = static_cast<const Number *>(p.inv_jacobian);
    const long long plane = p.n_cells * (long long)p.pad;
    const long long q     = (valid ? cell : p.cell_begin) * p.pad + i;
    Number J[3][3];
#pragma unroll
    for (int d = 0; d < 3; ++d)
#pragma unroll
      for (int e = 0; e < 3; ++e) J[d][e] = __ldg(ij + (d * 3 + e) * plane + q);
    const Number jxw = __ldg(static_cast<const Number *>(p.JxW) + q);
    const Number gr[3] = {gx[i], gy[i], gz[i]};
    Number gp[3], go[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) gp[d] = J[0][d] * gr[0] + J[1][d] * gr[1] + J[2][d] * gr[2]; // get_gradient: J^-T grad_ref
#pragma unroll
    for (int d = 0; d < 3; ++d) go[d] = (J[d][0] * gp[0] + J[d][1] * gp[1] + J[d][2] * gp[2]) * jxw; // submit_gradient
    gx[i] = go[0];
    gy[i] = go[1];
    gz[i] = go[2];
    __syncthreads();
  }
  // integrate(gradients)
  baseline_sweep<n, T_DC, true, false>(gx, values, 0, x, y, z);
  baseline_sweep<n, T_DC, true, true>(gy, values, 1, x, y, z);
  baseline_sweep<n, T_DC, true, true>(gz, values, 2, x, y, z);
  baseline_sweep<n, T_S, true, false>(values, values, 2, x, y, z);
  baseline_sweep<n, T_S, true, false>(values, values, 1, x, y, z);
  baseline_sweep<n, T_S, true, false>(values, values, 0, x, y, z);
  if (p.apply_constraints) baseline_resolve_hanging_nodes<n, true>(values, mask, x, y, z);
  if (valid) atomicAdd(dst + g, values[i]); // distribute_local_to_global
}

// fills the padded deal.II-style arrays on the device (Cartesian cells: J^-1 = I / h, JxW = w_q h^3)
template <int n, typename Number>
__global__ void baseline_setup_kernel(uint32_t *l2g, Number *inv_jac, Number *jxw, const uint32_t *idx, const Number *h, long long n_cells, int pad)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_cells * pad) return;
  const long long c = t / pad;
  const int i       = (int)(t % pad);
  constexpr int n3  = n * n * n;
  const long long plane = n_cells * (long long)pad;
  if (i < n3)
    {
      l2g[t]           = idx[c * n3 + i];
      const Number hh  = h[c];
      const int x = i % n, y = (i / n) % n, z = i / (n * n);
      jxw[t] = Shape<Number>::template qw<n>(x) * Shape<Number>::template qw<n>(y) * Shape<Number>::template qw<n>(z) * hh * hh * hh;
      for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) inv_jac[(d * 3 + e) * plane + t] = d == e ? Number(1) / hh : Number(0);
    }
  else
    {
      l2g[t] = 0;
      jxw[t] = Number(0);
      for (int d = 0; d < 9; ++d) inv_jac[d * plane + t] = Number(0);
    }
}
} // namespace mfhn
