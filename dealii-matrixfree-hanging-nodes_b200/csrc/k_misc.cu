// Auxiliary kernels: baseline-B (kernels_baseline.cuh), the stand-alone hanging-node interpolation
// (benchmark_00_likwid.cc:56-59), vector pack / unpack (update_ghost_values / compress of
// LinearAlgebra::distributed::Vector, benchmark_03.h:323-324) and the FMA micro-benchmark.
#include "kernels_baseline.cuh"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace mfhn
{
void run_generic_f64(int degree, int variant, bool diag, const CellLoopParams &p, int device, cudaStream_t stream);
void run_generic_f32(int degree, int variant, bool diag, const CellLoopParams &p, int device, cudaStream_t stream);
void run_generic(int degree, int number, int variant, bool diag, const CellLoopParams &p, int device, cudaStream_t stream)
{
  if (number == 0)
    run_generic_f64(degree, variant, diag, p, device, stream);
  else
    run_generic_f32(degree, variant, diag, p, device, stream);
}

namespace
{
void check_launch(const char *what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

template <int n, typename Number>
void launch_baseline(BaselineArrays &a, const uint32_t *d_idx, const void *d_h, long long n_cells, const CellLoopParams &p, cudaStream_t stream)
{
  using Cfg = BaselineCfg<n>;
  if (!a.l2g)
    {
      const size_t slots = (size_t)std::max<long long>(n_cells, 1) * Cfg::pad;
      if (cudaMalloc(&a.l2g, slots * sizeof(uint32_t)) != cudaSuccess || cudaMalloc(&a.invjac, slots * 9 * sizeof(Number)) != cudaSuccess ||
          cudaMalloc(&a.jxw, slots * sizeof(Number)) != cudaSuccess)
        throw std::runtime_error("baseline arrays: out of device memory");
      if (n_cells > 0)
        baseline_setup_kernel<n, Number><<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(a.l2g, (Number *)a.invjac, (Number *)a.jxw, d_idx,
                                                                                             (const Number *)d_h, n_cells, Cfg::pad);
      check_launch("baseline setup");
    }
  BaselineParams b;
  b.local_to_global   = a.l2g;
  b.inv_jacobian      = a.invjac;
  b.JxW               = a.jxw;
  b.masks             = p.masks;
  b.src               = p.src;
  b.dst               = p.dst;
  b.n_cells           = n_cells;
  b.cell_begin        = p.cell_begin;
  b.cell_end          = p.cell_end;
  b.pad               = Cfg::pad;
  b.apply_constraints = p.apply_constraints;
  const long long nc  = p.cell_end - p.cell_begin;
  if (nc <= 0) return;
  baseline_kernel<n, Number><<<(unsigned)((nc + Cfg::cpb - 1) / Cfg::cpb), Cfg::n3 * Cfg::cpb, 0, stream>>>(b);
  check_launch("baseline kernel");
}
template <typename Number>
void launch_baseline_number(int degree, BaselineArrays &a, const uint32_t *d_idx, const void *d_h, long long n_cells, const CellLoopParams &p,
                            cudaStream_t stream)
{
  switch (degree)
    {
      case 1: return launch_baseline<2, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 2: return launch_baseline<3, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 3: return launch_baseline<4, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 4: return launch_baseline<5, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 5: return launch_baseline<6, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 6: return launch_baseline<7, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 7: return launch_baseline<8, Number>(a, d_idx, d_h, n_cells, p, stream);
      case 8: return launch_baseline<9, Number>(a, d_idx, d_h, n_cells, p, stream);
      default: throw std::runtime_error("unsupported degree");
    }
}

// the three directional passes of the interpolation (or its transpose) on the n^3 values of one cell in shared memory;
// thread l = a + n b handles one line per direction
template <int n, typename Number>
__device__ __forceinline__ void hn_passes_block(Number *s, const int l, const unsigned mask, const bool transpose)
{
  const int a = l % n, b = l / n;
  unsigned face, edge, cb;
  decode_mask(mask, face, edge, cb);
  for (int d = 0; d < 3; ++d)
    {
      Number *line   = s + (d == 0 ? n * (a + n * b) : d == 1 ? a + n * n * b : a + n * b);
      const int strd = d == 0 ? 1 : d == 1 ? n : n * n;
      if (mask)
        {
          if (transpose)
            hn_pass_line<n, true>(line, strd, d, a, b, face, edge, cb);
          else
            hn_pass_line<n, false>(line, strd, d, a, b, face, edge, cb);
        }
      __syncthreads();
    }
}

template <int n, typename Number>
__global__ void hn_only_kernel(Number *values, const uint8_t *masks, long long n_cells, int transpose)
{
  // FEEvaluationHangingNodesFactory::apply on cell-local values (benchmark_00_likwid.cc:56-59)
  __shared__ Number s[n * n * n];
  const long long cell = blockIdx.x;
  if (masks[cell] == 0) return; // unconstrained cell: nothing to interpolate (block-uniform)
  const int l = threadIdx.x;
  Number *g = values + cell * (n * n * n);
  for (int z = 0; z < n; ++z) s[l + n * n * z] = g[l + n * n * z];
  const unsigned mask = masks[cell];
  __syncthreads();
  hn_passes_block<n>(s, l, mask, transpose != 0);
  for (int z = 0; z < n; ++z) g[l + n * n * z] = s[l + n * n * z];
}

// "DG (C)" of the reference's stage decomposition (benchmark_01.cc:189-199, benchmark_01.h:617-677 with VectorType1):
// every cell owns private DoFs, no quadrature-point work -- gather_plain [+ interpolation], [interpolation^T +]
// scatter_plain, dst += values.  Two kernels: a flat streaming pass over the entries of the unconstrained cells and one
// block per constrained cell (list built at setup) for W^T W.
template <typename Number>
__global__ void __launch_bounds__(256) dg_copy_plain_kernel(Number *dst, const Number *src, const uint8_t *masks, const long long n_entries, const int n3,
                                                             const int apply_constraints)
{
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_entries; i += (long long)gridDim.x * blockDim.x)
    if (!apply_constraints || masks[i / n3] == 0) dst[i] += src[i];
}
template <int n, typename Number>
__global__ void dg_copy_hn_kernel(Number *dst, const Number *src, const uint8_t *masks, const int32_t *hn_cells)
{
  __shared__ Number s[n * n * n];
  const long long cell = hn_cells[blockIdx.x];
  const int l          = threadIdx.x;
  const unsigned mask  = masks[cell];
  const Number *g      = src + cell * (n * n * n);
  Number *o            = dst + cell * (n * n * n);
  for (int z = 0; z < n; ++z) s[l + n * n * z] = g[l + n * n * z];
  __syncthreads();
  hn_passes_block<n>(s, l, mask, false);
  hn_passes_block<n>(s, l, mask, true);
  for (int z = 0; z < n; ++z) o[l + n * n * z] += s[l + n * n * z];
}

template <typename Number>
void hn_only(int degree, void *values, const uint8_t *d_masks, long long n_cells, int transpose, cudaStream_t st)
{
  if (n_cells == 0) return;
  const unsigned grid = (unsigned)n_cells;
#define HN_CASE(N)                                                                                    \
  case N - 1:                                                                                         \
    hn_only_kernel<N, Number><<<grid, N * N, 0, st>>>((Number *)values, d_masks, n_cells, transpose); \
    break;
  switch (degree)
    {
      HN_CASE(2) HN_CASE(3) HN_CASE(4) HN_CASE(5) HN_CASE(6) HN_CASE(7) HN_CASE(8) HN_CASE(9)
      default: throw std::runtime_error("unsupported degree");
    }
#undef HN_CASE
  check_launch("hanging-node kernel");
}

template <typename Number>
void dg_copy(int degree, void *dst, const void *src, const uint8_t *d_masks, long long n_cells, const int32_t *d_hn_cells, long long n_hn,
             int apply_constraints, cudaStream_t st)
{
  if (n_cells == 0) return;
  const int n3 = (degree + 1) * (degree + 1) * (degree + 1);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  dg_copy_plain_kernel<Number><<<sms * 16, 256, 0, st>>>((Number *)dst, (const Number *)src, d_masks, n_cells * n3, n3, apply_constraints);
  check_launch("DG copy kernel");
  if (!apply_constraints || n_hn == 0) return;
#define DG_CASE(N)                                                                                                \
  case N - 1:                                                                                                     \
    dg_copy_hn_kernel<N, Number><<<(unsigned)n_hn, N * N, 0, st>>>((Number *)dst, (const Number *)src, d_masks, d_hn_cells); \
    break;
  switch (degree)
    {
      DG_CASE(2) DG_CASE(3) DG_CASE(4) DG_CASE(5) DG_CASE(6) DG_CASE(7) DG_CASE(8) DG_CASE(9)
      default: throw std::runtime_error("unsupported degree");
    }
#undef DG_CASE
  check_launch("DG copy kernel (constrained cells)");
}

template <typename Number>
__global__ void fma_bench_kernel(Number *out, int iters)
{
  Number a[8], x = Number(1.0) + Number(1e-9) * threadIdx.x, y = Number(0.5);
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = Number(i) * Number(0.125) + x;
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = a[i] * x + y;
    }
  Number s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  if (s == Number(-1)) out[0] = s;
}

template <typename Number>
__global__ void pack_kernel(Number *buf, const Number *vec, const int32_t *idx, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = vec[idx[i]];
}
// atomic: several peers may contribute to the same owned entry while other kernels add to the vector
template <typename Number, bool ATOMIC>
__global__ void unpack_add_kernel(Number *vec, const Number *buf, const int32_t *idx, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    {
      if (ATOMIC)
        atomicAdd(vec + idx[i], buf[i]);
      else
        vec[idx[i]] += buf[i];
    }
}
// Barrier between the ranks (one process per GPU) of a peer-memory operator.  flags = unsigned[world + 1] in memory that
// the peers have mapped (CUDA IPC): slot r = last epoch rank r has announced here, entry `world` = this rank's epoch
// counter.  Thread p announces the new epoch in peer p's array (release, system scope: everything earlier kernels of
// this stream wrote -- also over NVLink -- is visible first), then waits until peer p has announced it here.
// All ranks call it the same number of times.
__global__ void peer_barrier_kernel(unsigned *flags, unsigned *const *peer_flags, const int rank, const int world)
{
  __shared__ unsigned epoch;
  if (threadIdx.x == 0) epoch = ++flags[world];
  __syncthreads();
  const int p = threadIdx.x;
  if (p < world && p != rank)
    {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[p] + rank), "r"(epoch) : "memory");
      unsigned seen;
      do
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + p) : "memory");
      while ((int)(seen - epoch) < 0);
    }
}
} // namespace

void run_peer_barrier(unsigned *flags_local, unsigned *const *d_peer_flags, int rank, int world, cudaStream_t stream)
{
  peer_barrier_kernel<<<1, 32, 0, stream>>>(flags_local, d_peer_flags, rank, world);
  check_launch("peer barrier");
}

void run_baseline(int degree, int number, BaselineArrays &arrays, const uint32_t *d_idx, const void *d_h, long long n_cells, const CellLoopParams &p,
                  int device, cudaStream_t stream)
{
  ensure_shape_tables(device);
  if (number == 0)
    launch_baseline_number<double>(degree, arrays, d_idx, d_h, n_cells, p, stream);
  else
    launch_baseline_number<float>(degree, arrays, d_idx, d_h, n_cells, p, stream);
}

void run_hn_only(int degree, int number, void *values, const uint8_t *d_masks, long long n_cells, int transpose, cudaStream_t stream)
{
  int device = 0;
  cudaGetDevice(&device);
  ensure_shape_tables(device);
  if (number == 0)
    hn_only<double>(degree, values, d_masks, n_cells, transpose, stream);
  else
    hn_only<float>(degree, values, d_masks, n_cells, transpose, stream);
}

void run_dg_copy(int degree, int number, void *dst, const void *src, const uint8_t *d_masks, long long n_cells, const int32_t *d_hn_cells, long long n_hn,
                 int apply_constraints, cudaStream_t stream)
{
  int device = 0;
  cudaGetDevice(&device);
  ensure_shape_tables(device);
  if (number == 0)
    dg_copy<double>(degree, dst, src, d_masks, n_cells, d_hn_cells, n_hn, apply_constraints, stream);
  else
    dg_copy<float>(degree, dst, src, d_masks, n_cells, d_hn_cells, n_hn, apply_constraints, stream);
}

double run_fma_bench(int number, int iters)
{
  void *out = nullptr;
  if (cudaMalloc(&out, 64) != cudaSuccess) throw std::runtime_error("cudaMalloc failed");
  cudaDeviceProp prop;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaGetDeviceProperties(&prop, dev);
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep)
    {
      cudaEventRecord(e0);
      if (number == 0)
        fma_bench_kernel<double><<<blocks, threads>>>((double *)out, iters);
      else
        fma_bench_kernel<float><<<blocks, threads>>>((float *)out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  check_launch("fma benchmark");
  return 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
}

void run_pack(int number, void *buffer, const void *vec, const int32_t *idx, long long n, cudaStream_t stream)
{
  if (n <= 0) return;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (number == 0)
    pack_kernel<double><<<grid, 256, 0, stream>>>((double *)buffer, (const double *)vec, idx, n);
  else
    pack_kernel<float><<<grid, 256, 0, stream>>>((float *)buffer, (const float *)vec, idx, n);
  check_launch("pack kernel");
}
void run_unpack_add(int number, void *vec, const void *buffer, const int32_t *idx, long long n, bool atomic, cudaStream_t stream)
{
  if (n <= 0) return;
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (number == 0)
    {
      if (atomic)
        unpack_add_kernel<double, true><<<grid, 256, 0, stream>>>((double *)vec, (const double *)buffer, idx, n);
      else
        unpack_add_kernel<double, false><<<grid, 256, 0, stream>>>((double *)vec, (const double *)buffer, idx, n);
    }
  else
    {
      if (atomic)
        unpack_add_kernel<float, true><<<grid, 256, 0, stream>>>((float *)vec, (const float *)buffer, idx, n);
      else
        unpack_add_kernel<float, false><<<grid, 256, 0, stream>>>((float *)vec, (const float *)buffer, idx, n);
    }
  check_launch("unpack kernel");
}
} // namespace mfhn
