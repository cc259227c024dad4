// Conjugate gradients with a point-Jacobi preconditioner around the operator (BASELINE.json config 5; the reference
// contains no solver, SURVEY.md 0.5 -- this is the step either side of vmult in a real solve).
//
// Chronopoulos / Gear form of preconditioned CG: ONE batched reduction per iteration (three scalars in one
// all-reduce), one vmult, and two fused vector kernels -- no other pass over the vectors:
//
//   u = D^-1 r,  w = A u,  (gamma, delta, rho) = (r.u, w.u, r.r)          [cg_dots_kernel + one all-reduce]
//   beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)   [cg_scalars_kernel, on the device]
//   p = u + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s,  u = D^-1 r    [cg_update_kernel: one pass]
//
// alpha, beta and the residual history stay on the device; the host reads the residual every `check_every`
// iterations only.  Vector entries [0, n_owned) take part (LinearAlgebra::distributed::Vector semantics: the ghost
// section belongs to the exchange inside vmult).
#include "layouts.hpp"

#include <cuda_runtime.h>

#include <cmath>
#include <stdexcept>
#include <string>

namespace mfhn
{
namespace
{
struct CgScalars
{
  double dots[3];   // gamma = r.u, delta = w.u, rho = r.r of the current iteration (accumulated, then reduced over the ranks)
  double gamma_old; // r.u of the previous iteration
  double alpha, beta;
  int iteration;
};

template <typename Number>
__global__ void __launch_bounds__(256) cg_dots_kernel(const Number *__restrict__ r, const Number *__restrict__ u, const Number *__restrict__ w,
                                                       const long long n, CgScalars *sc)
{
  double a = 0, b = 0, c = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
      const double ri = (double)r[i], ui = (double)u[i], wi = (double)w[i];
      a += ri * ui;
      b += wi * ui;
      c += ri * ri;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
  __shared__ double sh[3][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    {
      sh[0][warp] = a;
      sh[1][warp] = b;
      sh[2][warp] = c;
    }
  __syncthreads();
  if (threadIdx.x < 3)
    {
      double s = 0;
      for (int i = 0; i < 8; ++i) s += sh[threadIdx.x][i];
      atomicAdd(&sc->dots[threadIdx.x], s);
    }
}

// one thread: the scalars of the iteration from the (globally reduced) dot products
__global__ void cg_scalars_kernel(CgScalars *sc, double *history)
{
  const double gamma = sc->dots[0], delta = sc->dots[1], rho = sc->dots[2];
  const int it = sc->iteration;
  double alpha, beta;
  if (it == 0)
    {
      beta  = 0.0;
      alpha = delta != 0.0 ? gamma / delta : 0.0;
    }
  else
    {
      beta               = sc->gamma_old != 0.0 ? gamma / sc->gamma_old : 0.0;
      const double denom = delta - beta * gamma / sc->alpha;
      alpha              = denom != 0.0 ? gamma / denom : 0.0;
    }
  sc->alpha     = alpha;
  sc->beta      = beta;
  sc->gamma_old = gamma;
  if (history) history[it] = sqrt(rho > 0.0 ? rho : 0.0);
  sc->iteration = it + 1;
  sc->dots[0] = sc->dots[1] = sc->dots[2] = 0.0;
}

template <typename Number>
__global__ void __launch_bounds__(256) cg_update_kernel(Number *__restrict__ p, Number *__restrict__ s, Number *__restrict__ x, Number *__restrict__ r,
                                                         Number *__restrict__ u, const Number *__restrict__ w, const Number *__restrict__ inv_diag,
                                                         const long long n, const CgScalars *sc)
{
  const Number alpha = (Number)sc->alpha, beta = (Number)sc->beta;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
      const Number pi = u[i] + beta * p[i];
      const Number si = w[i] + beta * s[i];
      const Number ri = r[i] - alpha * si;
      p[i]            = pi;
      s[i]            = si;
      x[i] += alpha * pi;
      r[i] = ri;
      u[i] = inv_diag ? inv_diag[i] * ri : ri;
    }
}

// r = b - r (r holds A x), u = D^-1 r
template <typename Number>
__global__ void __launch_bounds__(256) cg_residual_kernel(Number *__restrict__ r, Number *__restrict__ u, const Number *__restrict__ b,
                                                           const Number *__restrict__ inv_diag, const long long n)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
      const Number ri = b[i] - r[i];
      r[i]            = ri;
      u[i]            = inv_diag ? inv_diag[i] * ri : ri;
    }
}

// inverse of the diagonal, 0 where the diagonal is 0 (hanging entries of the vector)
template <typename Number>
__global__ void __launch_bounds__(256) invert_diagonal_kernel(Number *d, const long long n)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = d[i] != Number(0) ? Number(1) / d[i] : Number(0);
}

void check_launch(const char *what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
unsigned grid_for(const long long n)
{
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long blocks = (n + 255) / 256;
  return (unsigned)std::max<long long>(1, std::min<long long>(blocks, (long long)sms * 8)); // grid-stride, a multiple of the SM count
}
} // namespace

size_t cg_scalars_bytes() { return sizeof(CgScalars); }

void run_cg_dots(int number, const void *r, const void *u, const void *w, long long n, void *scalars, cudaStream_t stream)
{
  if (number == 0)
    cg_dots_kernel<double><<<grid_for(n), 256, 0, stream>>>((const double *)r, (const double *)u, (const double *)w, n, (CgScalars *)scalars);
  else
    cg_dots_kernel<float><<<grid_for(n), 256, 0, stream>>>((const float *)r, (const float *)u, (const float *)w, n, (CgScalars *)scalars);
  check_launch("cg dots");
}
void run_cg_scalars(void *scalars, double *history, cudaStream_t stream)
{
  cg_scalars_kernel<<<1, 1, 0, stream>>>((CgScalars *)scalars, history);
  check_launch("cg scalars");
}
void run_cg_update(int number, void *p, void *s, void *x, void *r, void *u, const void *w, const void *inv_diag, long long n, const void *scalars,
                   cudaStream_t stream)
{
  if (number == 0)
    cg_update_kernel<double><<<grid_for(n), 256, 0, stream>>>((double *)p, (double *)s, (double *)x, (double *)r, (double *)u, (const double *)w,
                                                              (const double *)inv_diag, n, (const CgScalars *)scalars);
  else
    cg_update_kernel<float><<<grid_for(n), 256, 0, stream>>>((float *)p, (float *)s, (float *)x, (float *)r, (float *)u, (const float *)w,
                                                             (const float *)inv_diag, n, (const CgScalars *)scalars);
  check_launch("cg update");
}
void run_cg_residual(int number, void *r, void *u, const void *b, const void *inv_diag, long long n, cudaStream_t stream)
{
  if (number == 0)
    cg_residual_kernel<double><<<grid_for(n), 256, 0, stream>>>((double *)r, (double *)u, (const double *)b, (const double *)inv_diag, n);
  else
    cg_residual_kernel<float><<<grid_for(n), 256, 0, stream>>>((float *)r, (float *)u, (const float *)b, (const float *)inv_diag, n);
  check_launch("cg residual");
}
void run_invert_diagonal(int number, void *d, long long n, cudaStream_t stream)
{
  if (number == 0)
    invert_diagonal_kernel<double><<<grid_for(n), 256, 0, stream>>>((double *)d, n);
  else
    invert_diagonal_kernel<float><<<grid_for(n), 256, 0, stream>>>((float *)d, n);
  check_launch("invert diagonal");
}
} // namespace mfhn
