// Run-wise bulk-copy cell kernel (kernels_runs.cuh) and the host analysis of its layout.
#include "kernels_runs.cuh"

namespace mfhn
{
#define MFHN_RUNS_DISPATCH(nn, CALL_D, CALL_F)                                              \
  switch (nn)                                                                                \
    {                                                                                        \
      case 2: return f64 ? CALL_D(2) : CALL_F(2);                                            \
      case 3: return f64 ? CALL_D(3) : CALL_F(3);                                            \
      case 4: return f64 ? CALL_D(4) : CALL_F(4);                                            \
      case 5: return f64 ? CALL_D(5) : CALL_F(5);                                            \
      case 6: return f64 ? CALL_D(6) : CALL_F(6);                                            \
      default: throw std::runtime_error("MFHN_KERNEL_RUNS is available for degrees 1..5");   \
    }

void run_runs(int degree, int number, const RunsLayout &L, const CellLoopParams &p, int device, cudaStream_t stream)
{
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
  ensure_shape_tables(device);
  const bool f64 = number == 0;
#define D(n) launch_runs<n, double>(L, p, device, stream)
#define F(n) launch_runs<n, float>(L, p, device, stream)
  MFHN_RUNS_DISPATCH(degree + 1, D, F)
#undef D
#undef F
}
void runs_analyze(RunsHostLayout &L, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx, int max_gap, int min_run, bool place)
{
  const bool f64 = number == 0;
#define D(n) runs_analyze_impl<n, double>(L, n_cells, n_vec, idx, max_gap, min_run, place)
#define F(n) runs_analyze_impl<n, float>(L, n_cells, n_vec, idx, max_gap, min_run, place)
  MFHN_RUNS_DISPATCH(n, D, F)
#undef D
#undef F
}
long long runs_verify(const RunsHostLayout &L, int number, const uint32_t *idx, long long *staging_wavefronts)
{
  const bool f64 = number == 0;
#define D(n) runs_verify_impl<n, double>(L, idx, staging_wavefronts)
#define F(n) runs_verify_impl<n, float>(L, idx, staging_wavefronts)
  MFHN_RUNS_DISPATCH(L.n, D, F)
#undef D
#undef F
}
} // namespace mfhn
