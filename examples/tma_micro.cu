// micro-benchmark: throughput of many small 1D bulk copies (global->shared) and bulk reduce-adds (shared->global)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> // 0: loads only, 1: loads + reduces, 2: hex only loads + reduces, 3: reduces only
__global__ void __launch_bounds__(128, 4) k(double *g, const double *s, long long n_batches)
{
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *W = sm + warp * 4352; // 4224 B staging + mbarrier
  const unsigned sa = (unsigned)__cvta_generic_to_shared(W), ba = sa + 4224;
  const long long batch = (long long)blockIdx.x * 4 + warp;
  if (batch >= n_batches) return;
  if (lane == 0)
    {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
      asm volatile("fence.mbarrier_init.release.cluster;");
    }
  __syncwarp();
  // 6 cells x (1 hex of 28 doubles + 6 quads of 10 doubles)
  unsigned total = 0;
  for (int r = 0; r < 2; ++r)
    {
      const int i = r * 32 + lane;
      if (i < 42)
        {
          const int c = i / 7, o = i % 7;
          if (MODE == 2 && o != 0) continue;
          total += o == 0 ? 224 : 80;
        }
    }
  total = __reduce_add_sync(0xffffffffu, total);
  if (MODE != 3)
    {
      if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ba), "r"(total) : "memory");
      __syncwarp();
      for (int r = 0; r < 2; ++r)
        {
          const int i = r * 32 + lane;
          if (i < 42)
            {
              const int c = i / 7, o = i % 7;
              if (MODE == 2 && o != 0) continue;
              const unsigned bytes = o == 0 ? 224 : 80;
              const unsigned soff  = c * 704 + (o == 0 ? 0 : 224 + (o - 1) * 80);
              // cell block of 66 doubles per cell: hex then quads, scattered a little like object numbering
              const long long e = ((batch * 6 + c) * 100 + (o == 0 ? 0 : 30 + (o - 1) * 11)) & ~1ll;
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa + soff), "l"(s + e),
                           "r"(bytes), "r"(ba)
                           : "memory");
            }
        }
      asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@!p bra W;\n}" ::"r"(ba) : "memory");
    }
  if (MODE >= 1)
    {
      double *S = (double *)W;
      for (int j = lane; j < 528; j += 32) S[j] = S[j] * 0.5 + 1.0;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      for (int r = 0; r < 2; ++r)
        {
          const int i = r * 32 + lane;
          if (i < 42)
            {
              const int c = i / 7, o = i % 7;
              if (MODE == 2 && o != 0) continue;
              const unsigned bytes = o == 0 ? 224 : 80;
              const unsigned soff  = c * 704 + (o == 0 ? 0 : 224 + (o - 1) * 80);
              const long long e = ((batch * 6 + c) * 100 + (o == 0 ? 0 : 30 + (o - 1) * 11)) & ~1ll;
              asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(g + e), "r"(sa + soff), "r"(bytes)
                           : "memory");
            }
        }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  else
    {
      double *S = (double *)W;
      double a = 0;
      for (int j = lane; j < 528; j += 32) a += S[j];
      if (a == 1.2345) g[0] = a;
    }
}

// LSU reference: same bytes with plain coalesced loads + REDs
__global__ void __launch_bounds__(128, 4) k_lsu(double *g, const double *s, long long n_batches)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long batch = (long long)blockIdx.x * 4 + warp;
  if (batch >= n_batches) return;
  for (int c = 0; c < 6; ++c)
    {
      const long long base = (batch * 6 + c) * 100;
      for (int j = lane; j < 96; j += 32) atomicAdd(g + base + j, s[base + j] * 0.5 + 1.0);
    }
}

template <typename F>
void run(const char *name, F f, int reps, long long nb, double ops_per_batch)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= reps;
  cudaError_t e = cudaGetLastError();
  const double cyc_per_op = ms * 1e-3 * 1.9e9 / (nb * ops_per_batch / 148.0);
  printf("%-28s %.3f ms  ops/batch %.0f  ~%.1f cyc/op/SM @1.9GHz  (%s)\n", name, ms, ops_per_batch, cyc_per_op, cudaGetErrorString(e));
}

int main()
{
  const long long nb = 360268;
  const size_t n = (size_t)nb * 6 * 100 + 1024;
  double *s, *g;
  cudaMalloc(&s, n * 8);
  cudaMalloc(&g, n * 8);
  cudaMemset(s, 0, n * 8);
  cudaMemset(g, 0, n * 8);
  const int smem = 4 * 4352;
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const unsigned grid = (unsigned)((nb + 3) / 4);
  run("bulk loads only (42/warp)", [&] { k<0><<<grid, 128, smem>>>(g, s, nb); }, 20, nb, 42);
  run("bulk loads+reduces (84/warp)", [&] { k<1><<<grid, 128, smem>>>(g, s, nb); }, 20, nb, 84);
  run("hex only loads+reduces (12)", [&] { k<2><<<grid, 128, smem>>>(g, s, nb); }, 20, nb, 12);
  run("bulk reduces only (42/warp)", [&] { k<3><<<grid, 128, smem>>>(g, s, nb); }, 20, nb, 42);
  run("LSU coalesced ld+red", [&] { k_lsu<<<grid, 128>>>(g, s, nb); }, 20, nb, 1);
  cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
