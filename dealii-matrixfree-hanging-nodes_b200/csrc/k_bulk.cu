// Bulk-copy cell kernel (kernels_bulk.cuh) and the host analysis of its layout.
#include "kernels_bulk.cuh"

namespace mfhn
{
void run_bulk(int degree, int number, const BulkLayout &L, const CellLoopParams &p, int device, cudaStream_t stream)
{
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
  ensure_shape_tables(device);
  const bool f64 = number == 0;
  switch (degree + 1)
    {
      case 4: return f64 ? launch_bulk<4, double>(L, p, device, stream) : launch_bulk<4, float>(L, p, device, stream);
      case 5: return f64 ? launch_bulk<5, double>(L, p, device, stream) : launch_bulk<5, float>(L, p, device, stream);
      case 6: return f64 ? launch_bulk<6, double>(L, p, device, stream) : launch_bulk<6, float>(L, p, device, stream);
      default: throw std::runtime_error("MFHN_KERNEL_BULK is available for degrees 3..5");
    }
}
void bulk_analyze(BulkHostLayout &L, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx)
{
  const bool f64 = number == 0;
  switch (n)
    {
      case 4: f64 ? bulk_analyze_impl<4, double>(L, n_cells, n_vec, idx) : bulk_analyze_impl<4, float>(L, n_cells, n_vec, idx); break;
      case 5: f64 ? bulk_analyze_impl<5, double>(L, n_cells, n_vec, idx) : bulk_analyze_impl<5, float>(L, n_cells, n_vec, idx); break;
      case 6: f64 ? bulk_analyze_impl<6, double>(L, n_cells, n_vec, idx) : bulk_analyze_impl<6, float>(L, n_cells, n_vec, idx); break;
      default: throw std::runtime_error("MFHN_KERNEL_BULK is available for degrees 3..5");
    }
}
long long bulk_verify(const BulkHostLayout &L, int number, const uint32_t *idx)
{
  const bool f64 = number == 0;
  switch (L.n)
    {
      case 4: return f64 ? bulk_verify_impl<4, double>(L, idx) : bulk_verify_impl<4, float>(L, idx);
      case 5: return f64 ? bulk_verify_impl<5, double>(L, idx) : bulk_verify_impl<5, float>(L, idx);
      case 6: return f64 ? bulk_verify_impl<6, double>(L, idx) : bulk_verify_impl<6, float>(L, idx);
      default: throw std::runtime_error("MFHN_KERNEL_BULK is available for degrees 3..5");
    }
}
} // namespace mfhn
