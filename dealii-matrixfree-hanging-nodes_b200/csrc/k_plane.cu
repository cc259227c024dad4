// Plane kernels (kernels_plane.cuh: register-tiled, k <= 5; kernels_plane_smem.cuh: plane in shared memory, k = 6..8).
#include "kernels_plane_smem.cuh"

namespace mfhn
{
namespace
{
template <int n, typename Number>
void run_plane_n(const PlaneLayout &L, const CellLoopParams &p, int device, cudaStream_t stream, const PeerTables *peer)
{
  if constexpr (plane_supported(n))
    launch_plane<n, Number>(L, p, device, stream, peer);
  else
    launch_plane_smem<n, Number>(L, p, device, stream, peer);
}
template <typename Number>
void run_plane_number(int degree, const PlaneLayout &L, const CellLoopParams &p, int device, cudaStream_t stream, const PeerTables *peer)
{
  switch (degree)
    {
      case 1: return run_plane_n<2, Number>(L, p, device, stream, peer);
      case 2: return run_plane_n<3, Number>(L, p, device, stream, peer);
      case 3: return run_plane_n<4, Number>(L, p, device, stream, peer);
      case 4: return run_plane_n<5, Number>(L, p, device, stream, peer);
      case 5: return run_plane_n<6, Number>(L, p, device, stream, peer);
      case 6: return run_plane_n<7, Number>(L, p, device, stream, peer);
      case 7: return run_plane_n<8, Number>(L, p, device, stream, peer);
      case 8: return run_plane_n<9, Number>(L, p, device, stream, peer);
      default: throw std::runtime_error("unsupported degree");
    }
}
} // namespace

void run_plane(int degree, int number, const PlaneLayout &L, const CellLoopParams &p, int device, cudaStream_t stream, const PeerTables *peer)
{
  if (device < 0 || device >= 64) throw std::runtime_error("device ordinal out of range");
  ensure_shape_tables(device);
  if (number == 0)
    run_plane_number<double>(degree, L, p, device, stream, peer);
  else
    run_plane_number<float>(degree, L, p, device, stream, peer);
}
} // namespace mfhn
