// Micro-benchmark behind DESIGN.md "FP64 tensor cores": one 1D sweep of the sum factorisation at degree 7
// (n = 8: an 8 x 8 matrix applied to the 64 lines of a cell, data in shared memory) written three ways:
//   fma    one thread per line, dense 8 x 8 product            (64 DFMA per line)
//   eo     one thread per line, even-odd form of the kernels   (2 x 16 DFMA + 16 DADD per line)
//   dmma   mma.sync.aligned.m8n8k4.f64: a warp multiplies the matrix with 8 lines at a time (2 DMMA per 8 lines)
// Every variant reads a line from shared memory and writes the result back, like a sweep between two changes of the
// thread axis.  Output: lines per second and the FP64 rate each variant sustains.  build: make dmma_sweep
#include <cstdio>
#include <cuda_runtime.h>

constexpr int n = 8, ls = 9, cells_per_cta = 8, threads = 256; // ls: line stride, padded against bank conflicts
__constant__ double Tm[n * n];

template <int MODE>
__global__ void __launch_bounds__(threads) sweep(double *out, int iters)
{
  __shared__ double A[cells_per_cta][n * n * ls]; // [cell][line][entry], line stride ls
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < cells_per_cta * n * n * ls; i += threads) (&A[0][0])[i] = 1.0 + 1e-3 * (i % 17);
  __syncthreads();
  for (int it = 0; it < iters; ++it)
    {
      if (MODE < 2)
        {
          // 512 lines per CTA, 2 per thread
#pragma unroll
          for (int r = 0; r < 2; ++r)
            {
              const int l = tid + r * threads, c = l / (n * n), q = l % (n * n);
              double *line = &A[c][q * ls];
              double x[n], y[n];
#pragma unroll
              for (int i = 0; i < n; ++i) x[i] = line[i];
              if (MODE == 0)
                {
#pragma unroll
                  for (int i = 0; i < n; ++i)
                    {
                      double s = Tm[i * n] * x[0];
#pragma unroll
                      for (int j = 1; j < n; ++j) s += Tm[i * n + j] * x[j];
                      y[i] = s;
                    }
                }
              else
                {
                  double xs[4], xd[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    {
                      xs[j] = x[j] + x[7 - j];
                      xd[j] = x[j] - x[7 - j];
                    }
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    {
                      double e = Tm[i * n] * xs[0], o = Tm[(4 + i) * n] * xd[0];
#pragma unroll
                      for (int j = 1; j < 4; ++j)
                        {
                          e += Tm[i * n + j] * xs[j];
                          o += Tm[(4 + i) * n + j] * xd[j];
                        }
                      y[i]     = e + o;
                      y[7 - i] = e - o;
                    }
                }
#pragma unroll
              for (int i = 0; i < n; ++i) line[i] = y[i];
            }
        }
      else
        {
          // a warp owns one cell: 8 blocks of 8 lines; D[i][l] = sum_j T[i][j] X[j][l], two k-steps of 4
          // fragments (m8n8k4, f64): A: row = lane / 4, col = lane % 4;  B: row = lane % 4, col = lane / 4;
          //                          C/D: row = lane / 4, cols 2 (lane % 4), 2 (lane % 4) + 1
          const int c = warp;
          const double a0 = Tm[(lane / 4) * n + lane % 4], a1 = Tm[(lane / 4) * n + 4 + lane % 4];
#pragma unroll
          for (int blk = 0; blk < n; ++blk)
            {
              double *X = &A[c][blk * n * ls]; // 8 lines of 8 entries: X[l * ls + j]
              const double b0 = X[(lane / 4) * ls + lane % 4], b1 = X[(lane / 4) * ls + 4 + lane % 4];
              double d0 = 0.0, d1 = 0.0;
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a0), "d"(b0));
              asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a1), "d"(b1));
              __syncwarp();
              // D[i = lane / 4][l = 2 (lane % 4) + {0, 1}] -> line l, entry i
              X[(2 * (lane % 4)) * ls + lane / 4]     = d0;
              X[(2 * (lane % 4) + 1) * ls + lane / 4] = d1;
              __syncwarp();
            }
        }
      __syncthreads();
    }
  double s = 0;
  for (int i = tid; i < cells_per_cta * n * n * ls; i += threads) s += (&A[0][0])[i];
  if (s == -1.0) out[0] = s;
}

template <int MODE>
double run(const char *name, double flops_per_line, int iters)
{
  double *out;
  cudaMalloc(&out, 8);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int grid = prop.multiProcessorCount * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  sweep<MODE><<<grid, threads>>>(out, 10);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep)
    {
      cudaEventRecord(e0);
      sweep<MODE><<<grid, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = ms < best ? ms : best;
    }
  const double lines = (double)grid * cells_per_cta * n * n * iters;
  printf("{\"variant\": \"%s\", \"ms\": %.3f, \"glines_per_s\": %.2f, \"tflops_executed\": %.2f, \"error\": \"%s\"}\n", name, best, lines / best / 1e6,
         lines * flops_per_line / best / 1e9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  return best;
}

int main()
{
  double T[n * n];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) T[i * n + j] = (i == j ? 0.5 : 0.0) + 0.01 * ((i * 7 + j * 3) % 5 - 2); // keeps the iterates bounded
  cudaMemcpyToSymbol(Tm, T, sizeof(T));
  const int iters = 2000;
  run<0>("fma dense 8x8 (128 flop/line)", 128, iters);
  run<1>("fma even-odd (2x16 DFMA + 16 DADD = 80 flop/line)", 80, iters);
  run<2>("dmma m8n8k4 (2 DMMA per 8 lines = 128 flop/line)", 128, iters);
  return 0;
}
