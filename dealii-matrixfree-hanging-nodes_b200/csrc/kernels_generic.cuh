// Generic fused cell kernel: all degrees 1..8, double / float, Cartesian or
// affine geometry.  One thread owns one 1D line of a cell per sweep; cell data
// lives in shared memory between sweeps.  This is the general path (and the
// fallback for degrees the register-tiled kernel does not cover); the fast
// Cartesian path is kernels_plane.cuh.
//
// Fuses the reference's LaplaceOperatorLocal::operator() (benchmark_03.h:297-313):
// read_dof_values -> hanging-node interpolation -> evaluate(gradients) ->
// submit_gradient(get_gradient) -> integrate(gradients) -> interpolation^T ->
// distribute_local_to_global (atomic add).
#pragma once
#include "layouts.hpp"
#include "shape_tables.cuh"

#include <cstdint>

namespace mfhn
{
template <int n, int T, bool transpose, typename Number>
__device__ __forceinline__ void mat_vec(const Number (&in)[n], Number (&out)[n])
{
#pragma unroll
  for (int i = 0; i < n; ++i)
    {
      Number s = Number(0);
#pragma unroll
      for (int j = 0; j < n; ++j)
        s += Shape<Number>::template get<n, T>(transpose ? j * n + i : i * n + j) * in[j];
      out[i] = s;
    }
}

__device__ __forceinline__ void decode_mask(unsigned m, unsigned &face, unsigned &edge, unsigned &childbits)
{
  // compressed_constraint_kind -> face / edge bits; child bit = 1 - subcell bit
  const unsigned v = m >> 5;
  face             = (m & 8u) ? v : 0u;
  edge             = (m & 16u) ? v : 0u;
  childbits        = (~m) & 7u;
}

template <int n, typename Number>
__device__ __forceinline__ void load_line(const Number *base, int stride, Number (&v)[n])
{
#pragma unroll
  for (int i = 0; i < n; ++i) v[i] = base[i * stride];
}
template <int n, typename Number>
__device__ __forceinline__ void store_line(Number *base, int stride, const Number (&v)[n])
{
#pragma unroll
  for (int i = 0; i < n; ++i) base[i * stride] = v[i];
}

// One directional pass of the hanging-node interpolation for the line owned
// by thread (a,b) = (coordinate in the lower, higher transversal direction).
template <int n, bool transpose, typename Number>
__device__ __forceinline__ void hn_pass_line(Number *line, int stride, int d, int a, int b, unsigned face,
                                             unsigned edge, unsigned childbits)
{
  constexpr int k = n - 1;
  const int t0 = (d == 0) ? 1 : 0, t1 = (d == 2) ? 1 : 2;
  const bool on0 = a == (int)((childbits >> t0) & 1u) * k;
  const bool on1 = b == (int)((childbits >> t1) & 1u) * k;
  const bool sel = (((face >> t0) & 1u) && on0) || (((face >> t1) & 1u) && on1) || (((edge >> d) & 1u) && on0 && on1);
  if (!sel) return;
  const bool upper = (childbits >> d) & 1u; // upper child: W_1[i][j] = W_0[k-i][k-j]
  Number v[n], w[n];
#pragma unroll
  for (int i = 0; i < n; ++i) v[i] = line[(upper ? k - i : i) * stride];
  mat_vec<n, T_W0, transpose>(v, w);
#pragma unroll
  for (int i = 0; i < n; ++i) line[(upper ? k - i : i) * stride] = w[i];
}

template <int n>
struct GenericCfg
{
  static constexpr int nx  = n | 1;                 // padded row length (odd: conflict-free x sweeps)
  static constexpr int cs  = n * n * nx;            // array stride per cell
  static constexpr int tpc = n * n;                 // threads per cell
  // n^2 <= 32: the threads of a cell sit in ONE warp (32 / n^2 cells per warp, the other lanes leave at once), so the
  // barriers between the sweeps are __syncwarp over those lanes instead of block barriers over 10 unrelated cells
  static constexpr int wc  = tpc <= 32 ? 32 / tpc : 0; // cells per warp (0: block mode)
  static constexpr unsigned lanes = wc ? (wc * tpc == 32 ? 0xffffffffu : (1u << (wc * tpc)) - 1u) : 0u;
  static constexpr int cpb = wc ? 8 * wc : ((256 / tpc) > 0 ? (256 / tpc) : 1);
  static constexpr int threads = wc ? 256 : cpb * tpc;
};

template <int VARIANT>
__host__ __device__ constexpr int generic_n_arrays()
{
  return (VARIANT == GV_QPOINT_METRIC || VARIANT == GV_QPOINT_GENERAL) ? 4 : 2;
}

// DIAG = true computes the diagonal of the operator instead of applying it: for every local
// DoF j the cell operator (with the hanging-node interpolation and its transpose) is applied to
// the unit vector e_j and entry j of the result is added to dst -- (k+1)^3 cell applications per
// cell, a one-time setup cost of a point-Jacobi preconditioner.
template <int n, typename Number, int VARIANT, bool DIAG = false>
__global__ void __launch_bounds__(GenericCfg<n>::threads) generic_cell_kernel(const CellLoopParams p)
{
  using Cfg        = GenericCfg<n>;
  constexpr int nx = Cfg::nx, cs = Cfg::cs, k = n - 1;
  constexpr int NA = generic_n_arrays<VARIANT>();
  constexpr bool ROWS = VARIANT == GV_QPOINT_ROWS; // general-purpose constraint algorithm around the Cartesian q-point sequence
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Number *smem = reinterpret_cast<Number *>(smem_raw);

  const int tid  = threadIdx.x;
  if (Cfg::wc && (tid & 31) >= Cfg::wc * Cfg::tpc) return; // lanes beyond the warp's cells
  const int cib  = Cfg::wc ? (tid >> 5) * Cfg::wc + (tid & 31) / Cfg::tpc : tid / Cfg::tpc;
  const int l    = Cfg::wc ? (tid & 31) % Cfg::tpc : tid % Cfg::tpc;
  auto cell_sync = [&]() {
    if (Cfg::wc)
      __syncwarp(Cfg::lanes);
    else
      __syncthreads();
  };
  const int a    = l % n;
  const int b    = l / n;
  const long long cell = p.cell_begin + (long long)blockIdx.x * Cfg::cpb + cib;
  const bool valid     = cell < p.cell_end;
  Number *A0 = smem + (size_t)cib * NA * cs; // U (then R / P)
  Number *A1 = A0 + cs;

  const Number *src = static_cast<const Number *>(p.src);
  Number *dst       = static_cast<Number *>(p.dst);

  // line bases for the sweeps along x, y, z of thread (a,b)
  const int bx = nx * (a + n * b), by = a + nx * n * b, bz = a + nx * b;
  constexpr int sx = 1, sy = nx, sz = nx * n;

  uint32_t gidx[n];
  unsigned mask = 0;
  if (valid)
    {
      const uint32_t *ip = p.idx + cell * (long long)(n * n * n) + l;
#pragma unroll
      for (int z = 0; z < n; ++z) gidx[z] = ip[z * n * n];
      if (p.apply_constraints && !ROWS) mask = p.masks[cell];
    }
  // ROWS: constrained cells gather / scatter through the weighted rows of their mask's interpolation matrix
  const int row_kind = (ROWS && valid && p.apply_constraints) ? p.row_kind[cell] : 0;
  const uint32_t *cell_idx = p.idx + cell * (long long)(n * n * n);
  const int32_t *row_ptr   = ROWS ? p.row_ptr + (long long)(row_kind > 0 ? row_kind - 1 : 0) * (n * n * n + 1) : nullptr;
#pragma unroll 1
  for (int unit = 0; unit < (DIAG ? n * n * n : 1); ++unit)
  {
  if (DIAG)
    {
      cell_sync();
#pragma unroll
      for (int z = 0; z < n; ++z) A0[bz + z * sz] = (l + n * n * z == unit) ? Number(1) : Number(0);
    }
  else if (valid)
    {
      if (ROWS && row_kind > 0)
        {
          const Number *rv = static_cast<const Number *>(p.row_val);
#pragma unroll 1
          for (int z = 0; z < n; ++z)
            {
              const int i = l + n * n * z;
              Number s    = Number(0);
              for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) s += rv[e] * src[cell_idx[p.row_col[e]]];
              A0[bz + z * sz] = s;
            }
        }
      else
        {
#pragma unroll
          for (int z = 0; z < n; ++z) A0[bz + z * sz] = src[gidx[z]];
        }
    }
  const bool any_hn = Cfg::wc ? __any_sync(Cfg::lanes, mask != 0) : __syncthreads_or(mask != 0);
  unsigned face = 0, edge = 0, childbits = 0;
  decode_mask(mask, face, edge, childbits);
  if (any_hn)
    {
      if (mask) hn_pass_line<n, false>(A0 + bx, sx, 0, a, b, face, edge, childbits);
      cell_sync();
      if (mask) hn_pass_line<n, false>(A0 + by, sy, 1, a, b, face, edge, childbits);
      cell_sync();
      if (mask) hn_pass_line<n, false>(A0 + bz, sz, 2, a, b, face, edge, childbits);
      cell_sync();
    }

  Number u[n], v[n], w[n];
  if (VARIANT == GV_SEPARABLE)
    {
      const Number h = valid ? static_cast<const Number *>(p.geom)[cell] : Number(0);
      // x: p = M u, q = K u
      load_line<n>(A0 + bx, sx, u);
      mat_vec<n, T_M, false>(u, v);
      mat_vec<n, T_K, false>(u, w);
      store_line<n>(A0 + bx, sx, v);
      store_line<n>(A1 + bx, sx, w);
      cell_sync();
      // y: a = M p, b = M q + K p
      load_line<n>(A0 + by, sy, u);
      load_line<n>(A1 + by, sy, v);
      mat_vec<n, T_M, false>(v, w); // M q
      mat_vec<n, T_K, false>(u, v); // K p
#pragma unroll
      for (int i = 0; i < n; ++i) w[i] += v[i];
      mat_vec<n, T_M, false>(u, v); // M p
      store_line<n>(A0 + by, sy, v);
      store_line<n>(A1 + by, sy, w);
      cell_sync();
      // z: r = h (M b + K a)
      load_line<n>(A0 + bz, sz, u);
      load_line<n>(A1 + bz, sz, v);
      mat_vec<n, T_M, false>(v, w);
      mat_vec<n, T_K, false>(u, v);
#pragma unroll
      for (int i = 0; i < n; ++i) w[i] = h * (w[i] + v[i]);
      store_line<n>(A0 + bz, sz, w);
      cell_sync();
    }
  else
    {
      // evaluate: basis change to Gauss collocation, three sweeps in place
      load_line<n>(A0 + bx, sx, u);
      mat_vec<n, T_S, false>(u, v);
      store_line<n>(A0 + bx, sx, v);
      cell_sync();
      load_line<n>(A0 + by, sy, u);
      mat_vec<n, T_S, false>(u, v);
      store_line<n>(A0 + by, sy, v);
      cell_sync();
      load_line<n>(A0 + bz, sz, u);
      mat_vec<n, T_S, false>(u, v);
      store_line<n>(A0 + bz, sz, v);
      cell_sync();
      if (VARIANT == GV_QPOINT_CARTESIAN || ROWS)
        {
          // gradient, q-point factor w_q * h (Cartesian: J^-1 J^-T detJ = h), integrate; direction by direction
          const Number h  = valid ? static_cast<const Number *>(p.geom)[cell] : Number(0);
          Number wq[n];
#pragma unroll
          for (int i = 0; i < n; ++i) wq[i] = Shape<Number>::template qw<n>(i);
          Number wA = Number(0), wB = Number(0);
#pragma unroll
          for (int i = 0; i < n; ++i)
            {
              if (i == a) wA = wq[i];
              if (i == b) wB = wq[i];
            }
          const Number fac = h * wA * wB;
          // x
          load_line<n>(A0 + bx, sx, u);
          mat_vec<n, T_DC, false>(u, v);
#pragma unroll
          for (int i = 0; i < n; ++i) v[i] *= wq[i] * fac;
          mat_vec<n, T_DC, true>(v, w);
          store_line<n>(A1 + bx, sx, w);
          cell_sync();
          // y
          load_line<n>(A0 + by, sy, u);
          mat_vec<n, T_DC, false>(u, v);
#pragma unroll
          for (int i = 0; i < n; ++i) v[i] *= wq[i] * fac;
          mat_vec<n, T_DC, true>(v, w);
          load_line<n>(A1 + by, sy, v);
#pragma unroll
          for (int i = 0; i < n; ++i) w[i] += v[i];
          store_line<n>(A1 + by, sy, w);
          cell_sync();
          // z
          load_line<n>(A0 + bz, sz, u);
          mat_vec<n, T_DC, false>(u, v);
#pragma unroll
          for (int i = 0; i < n; ++i) v[i] *= wq[i] * fac;
          mat_vec<n, T_DC, true>(v, w);
          load_line<n>(A1 + bz, sz, v);
#pragma unroll
          for (int i = 0; i < n; ++i) w[i] += v[i];
          store_line<n>(A0 + bz, sz, w); // result back into A0
          cell_sync();
        }
      else
        {
          Number *GX = A0 + cs, *GY = A0 + 2 * cs, *GZ = A0 + 3 * cs;
          Number G[6] = {0, 0, 0, 0, 0, 0};
          if (valid)
            {
#pragma unroll
              for (int i = 0; i < 6; ++i)
                if (VARIANT == GV_QPOINT_METRIC) G[i] = static_cast<const Number *>(p.geom)[cell * 6 + i];
            }
          load_line<n>(A0 + bx, sx, u);
          mat_vec<n, T_DC, false>(u, v);
          store_line<n>(GX + bx, sx, v);
          load_line<n>(A0 + by, sy, u);
          mat_vec<n, T_DC, false>(u, v);
          store_line<n>(GY + by, sy, v);
          load_line<n>(A0 + bz, sz, u);
          mat_vec<n, T_DC, false>(u, v);
          store_line<n>(GZ + bz, sz, v);
          cell_sync();
          // q-point operation on the z line of thread (x=a, y=b): g <- w_q * G g
          {
            Number wq[n];
#pragma unroll
            for (int i = 0; i < n; ++i) wq[i] = Shape<Number>::template qw<n>(i);
            Number wA = Number(0), wB = Number(0);
#pragma unroll
            for (int i = 0; i < n; ++i)
              {
                if (i == a) wA = wq[i];
                if (i == b) wB = wq[i];
              }
#pragma unroll
            for (int z = 0; z < n; ++z)
              {
                const Number gx = GX[bz + z * sz], gy = GY[bz + z * sz], gz = GZ[bz + z * sz];
                Number ww = wA * wB * wq[z];
                if (VARIANT == GV_QPOINT_GENERAL)
                  {
                    // submit_gradient with this point's own coefficient (JxW included): coalesced over (x,y)
                    ww = Number(1);
                    if (valid)
                      {
                        const Number *gq = static_cast<const Number *>(p.geom) + (cell * 6) * (long long)(n * n * n) + (l + n * n * z);
#pragma unroll
                        for (int i = 0; i < 6; ++i) G[i] = __ldg(gq + i * (n * n * n));
                      }
                  }
                GX[bz + z * sz] = ww * (G[0] * gx + G[1] * gy + G[2] * gz);
                GY[bz + z * sz] = ww * (G[1] * gx + G[3] * gy + G[4] * gz);
                GZ[bz + z * sz] = ww * (G[2] * gx + G[4] * gy + G[5] * gz);
              }
          }
          cell_sync();
          load_line<n>(GX + bx, sx, u);
          mat_vec<n, T_DC, true>(u, v);
          store_line<n>(A0 + bx, sx, v);
          cell_sync();
          load_line<n>(GY + by, sy, u);
          mat_vec<n, T_DC, true>(u, v);
          load_line<n>(A0 + by, sy, w);
#pragma unroll
          for (int i = 0; i < n; ++i) v[i] += w[i];
          store_line<n>(A0 + by, sy, v);
          cell_sync();
          load_line<n>(GZ + bz, sz, u);
          mat_vec<n, T_DC, true>(u, v);
          load_line<n>(A0 + bz, sz, w);
#pragma unroll
          for (int i = 0; i < n; ++i) v[i] += w[i];
          store_line<n>(A0 + bz, sz, v);
          cell_sync();
        }
      // integrate: S^T in z, y, x
      load_line<n>(A0 + bz, sz, u);
      mat_vec<n, T_S, true>(u, v);
      store_line<n>(A0 + bz, sz, v);
      cell_sync();
      load_line<n>(A0 + by, sy, u);
      mat_vec<n, T_S, true>(u, v);
      store_line<n>(A0 + by, sy, v);
      cell_sync();
      load_line<n>(A0 + bx, sx, u);
      mat_vec<n, T_S, true>(u, v);
      store_line<n>(A0 + bx, sx, v);
      cell_sync();
    }

  if (any_hn)
    {
      if (mask) hn_pass_line<n, true>(A0 + bx, sx, 0, a, b, face, edge, childbits);
      cell_sync();
      if (mask) hn_pass_line<n, true>(A0 + by, sy, 1, a, b, face, edge, childbits);
      cell_sync();
      if (mask) hn_pass_line<n, true>(A0 + bz, sz, 2, a, b, face, edge, childbits);
      cell_sync();
    }
  if (valid)
    {
      if (ROWS && row_kind > 0)
        {
          // distribute_local_to_global with constraints: every entry spreads over its row (transposed weights)
          const Number *rv = static_cast<const Number *>(p.row_val);
#pragma unroll 1
          for (int z = 0; z < n; ++z)
            {
              const int i    = l + n * n * z;
              const Number r = A0[bz + z * sz];
              for (int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) atomicAdd(dst + cell_idx[p.row_col[e]], rv[e] * r);
            }
        }
      else
        {
#pragma unroll
          for (int z = 0; z < n; ++z)
            if (!DIAG || l + n * n * z == unit) atomicAdd(dst + gidx[z], A0[bz + z * sz]);
        }
    }
  } // unit vectors
}
} // namespace mfhn
