// Operator object behind the C ABI: device-resident MatrixFree data + kernel
// dispatch.  Stands in for CUDAWrappers::MatrixFree::reinit / cell_loop as used
// by LaplaceOperator<..., MemorySpace::CUDA> (benchmark_03.h:319-357).
#include "../../include/mfhn.h"
#include "error.hpp"
#include "fe1d.hpp"
#include "kernels_generic.cuh"
#include "kernels_plane.cuh"
#include "kernels_bulk.cuh"
#include "kernels_patch.cuh"
#include "kernels_plane_smem.cuh"
#include "dist.cuh"
#include "kernels_baseline.cuh"
#include "octree.hpp"

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <cstdlib>
#include <memory>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

namespace mfhn
{

#define CUDA_CHECK(x)                                                                            \
  do                                                                                             \
    {                                                                                            \
      cudaError_t e_ = (x);                                                                      \
      if (e_ != cudaSuccess)                                                                     \
        throw CudaError(std::string(#x) + ": " + cudaGetErrorString(e_));                        \
    }                                                                                            \
  while (0)

namespace
{
std::mutex g_table_mutex;
bool g_tables_uploaded[64] = {};

void upload_tables(int device)
{
  std::lock_guard<std::mutex> lock(g_table_mutex);
  if (device >= 0 && device < 64 && g_tables_uploaded[device]) return;
  static ShapeTables<double> hd;
  static ShapeTables<float> hf;
  std::memset(&hd, 0, sizeof(hd));
  for (int k = 1; k <= 8; ++k)
    {
      const Shape1D s = make_shape(k);
      const int n = k + 1, h = n / 2, he = (n + 1) / 2;
      auto &t = hd.full[k - 1];
      for (int i = 0; i < n * n; ++i)
        {
          t[T_S][i]  = s.S[i];
          t[T_DC][i] = s.Dc[i];
          t[T_W0][i] = s.W[0][i];
          t[T_M][i]  = s.M[i];
          t[T_K][i]  = s.K[i];
        }
      for (int i = 0; i < n; ++i) hd.qw[k - 1][i] = s.qw[i];
      // even-odd halves of the persymmetric M and K: E = (A[i][j] + A[i][n-1-j]) / 2 (middle
      // column: A[i][m]), O = (A[i][j] - A[i][n-1-j]) / 2
      if (he <= MAX_HE)
        for (int which = 0; which < 2; ++which)
          {
            const std::vector<double> &A = which == 0 ? s.M : s.K;
            double *E = hd.eo[k - 1][which == 0 ? T_ME : T_KE], *O = hd.eo[k - 1][which == 0 ? T_MO : T_KO];
            for (int i = 0; i < he; ++i)
              for (int j = 0; j < he; ++j)
                {
                  if (j < h)
                    {
                      E[i * he + j] = 0.5 * (A[i * n + j] + A[i * n + (n - 1 - j)]);
                      O[i * he + j] = 0.5 * (A[i * n + j] - A[i * n + (n - 1 - j)]);
                    }
                  else
                    {
                      E[i * he + j] = A[i * n + j];
                      O[i * he + j] = 0;
                    }
                }
          }
    }
  {
    const double *ps = reinterpret_cast<const double *>(&hd);
    float *pf        = reinterpret_cast<float *>(&hf);
    for (size_t i = 0; i < sizeof(hd) / sizeof(double); ++i) pf[i] = (float)ps[i];
  }
  CUDA_CHECK(cudaMemcpyToSymbol(c_shape_d, &hd, sizeof(hd)));
  CUDA_CHECK(cudaMemcpyToSymbol(c_shape_f, &hf, sizeof(hf)));
  if (device >= 0 && device < 64) g_tables_uploaded[device] = true;
}

template <typename T>
T *to_device(const std::vector<T> &h)
{
  T *d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
  if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}
} // namespace

void PlaneLayout::build(int n_, long long n_cells_, const uint32_t *idx)
{
  n               = n_;
  n_cells         = n_cells_;
  const int cpw   = 32 / n;
  n_batches       = (n_cells + cpw - 1) / cpw;
  const int n2    = n * n;
  const long long n3 = (long long)n2 * n;
  std::vector<uint32_t> p((size_t)std::max<long long>(n_batches, 1) * n2 * 32, 0xffffffffu);
#pragma omp parallel for schedule(static)
  for (long long c = 0; c < n_cells; ++c)
    {
      const long long batch = c / cpw;
      const int slot        = (int)(c % cpw);
      // kernel axes (X,Y,Z) = physical (y,z,x): thread t = x, plane slot j = y + n z
      for (int t = 0; t < n; ++t)
        for (int j = 0; j < n2; ++j) p[((size_t)batch * n2 + j) * 32 + slot * n + t] = idx[c * n3 + t + (long long)n * j];
    }
  d_pidx = to_device(p);
}

// ---- bulk-copy layout -------------------------------------------------------------
static void bulk_analyze(BulkHostLayout &L, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx)
{
  const bool f64 = number == MFHN_F64;
  switch (n)
    {
      case 4: f64 ? bulk_analyze_impl<4, double>(L, n_cells, n_vec, idx) : bulk_analyze_impl<4, float>(L, n_cells, n_vec, idx); break;
      case 5: f64 ? bulk_analyze_impl<5, double>(L, n_cells, n_vec, idx) : bulk_analyze_impl<5, float>(L, n_cells, n_vec, idx); break;
      case 6: f64 ? bulk_analyze_impl<6, double>(L, n_cells, n_vec, idx) : bulk_analyze_impl<6, float>(L, n_cells, n_vec, idx); break;
      default: throw NotImplemented("MFHN_KERNEL_BULK is available for degrees 3..5");
    }
}
static long long bulk_verify(const BulkHostLayout &L, int number, const uint32_t *idx)
{
  const bool f64 = number == MFHN_F64;
  switch (L.n)
    {
      case 4: return f64 ? bulk_verify_impl<4, double>(L, idx) : bulk_verify_impl<4, float>(L, idx);
      case 5: return f64 ? bulk_verify_impl<5, double>(L, idx) : bulk_verify_impl<5, float>(L, idx);
      case 6: return f64 ? bulk_verify_impl<6, double>(L, idx) : bulk_verify_impl<6, float>(L, idx);
      default: throw NotImplemented("MFHN_KERNEL_BULK is available for degrees 3..5");
    }
}
// more irregular cells than this: the layout does not fit the numbering, the plane kernel runs everything
constexpr long long bulk_max_irregular = 32;

static void bulk_build(BulkLayout &B, int n, int number, long long n_cells, long long n_vec, const uint32_t *idx)
{
  BulkHostLayout L;
  bulk_analyze(L, n, number, n_cells, n_vec, idx);
  B.n         = n;
  B.n_cells   = n_cells;
  B.n_batches = L.n_batches;
  B.irregular = L.irregular;
  B.usable    = (long long)L.irregular.size() <= bulk_max_irregular;
  if (!B.usable) return;
  B.d_bidx  = to_device(L.bidx);
  B.d_lvidx = to_device(L.lvidx);
  B.d_cinfo = to_device(L.cinfo);
}

template <int n, typename Number>
static int patch_slot(int s, int j)
{
  using Cfg = PatchCfg<n, Number>;
  return Cfg::slot(s, j % n, (j / n) % n, j / (n * n));
}
template <typename Number>
static int patch_slot_n(int n, int s, int j)
{
  switch (n)
    {
      case 2: return patch_slot<2, Number>(s, j);
      case 3: return patch_slot<3, Number>(s, j);
      case 4: return patch_slot<4, Number>(s, j);
      case 5: return patch_slot<5, Number>(s, j);
      case 6: return patch_slot<6, Number>(s, j);
      default: throw InvalidArgument("patch kernel not available for this degree");
    }
}

void PatchLayout::build(int n_, int number_, long long n_cells_, const uint32_t *idx)
{
  n       = n_;
  number  = number_;
  n_cells = n_cells_;
  const int cpw = 32 / n, n2 = n * n, n3 = n2 * n, ent_stride = cpw * n3;
  const int rounds = (ent_stride + 31) / 32, u_stride = rounds * 32;
  n_patches = (n_cells + cpw - 1) / cpw;
  const size_t np = (size_t)std::max<long long>(n_patches, 1);
  std::vector<uint32_t> uidx(np * u_stride, 0u);
  std::vector<uint16_t> lidx(np * n2 * 32, 0);
  std::vector<uint16_t> ent(np * ent_stride, 0);
  std::vector<PatchInfo> info(np);
  long long total = 0;
#pragma omp parallel reduction(+ : total)
  {
    struct Entry { uint32_t g; uint16_t slot, lpos; }; // lpos = plane slot * 32 + lane
    std::vector<Entry> entries;
    struct Group { uint32_t g; int first, count; };
    std::vector<Group> groups;
#pragma omp for schedule(dynamic, 64)
    for (long long pt = 0; pt < n_patches; ++pt)
      {
        const long long cb = pt * cpw;
        const int nc       = (int)std::min<long long>(cpw, n_cells - cb);
        entries.clear();
        for (int s = 0; s < nc; ++s)
          for (int j = 0; j < n3; ++j)
            {
              const int x = j % n, y = (j / n) % n, z = j / n2;
              const int slot = number == MFHN_F64 ? patch_slot_n<double>(n, s, j) : patch_slot_n<float>(n, s, j);
              // thread t = x of cell s sits in lane s n + x; its plane slot is y + n z (kernel axes (X,Y,Z) = (y,z,x))
              entries.push_back(Entry{idx[(cb + s) * n3 + j], (uint16_t)slot, (uint16_t)((y + n * z) * 32 + s * n + x)});
            }
        std::sort(entries.begin(), entries.end(), [](const Entry &a, const Entry &b) { return a.g < b.g; });
        groups.clear();
        for (int i = 0; i < (int)entries.size();)
          {
            int j = i;
            while (j < (int)entries.size() && entries[j].g == entries[i].g) ++j;
            // multiplicities other than 1, 2, 4, 8 are split (3 = 2 + 1, ...): such a DoF is listed more
            // than once, i.e. read twice and sent two REDs
            int first = i, left = j - i;
            while (left > 0)
              {
                const int piece = left >= 8 ? 8 : left >= 4 ? 4 : left >= 2 ? 2 : 1;
                groups.push_back(Group{entries[i].g, first, piece});
                first += piece;
                left -= piece;
              }
            i = j;
          }
        // classes by multiplicity (uniform trip counts), ascending address inside a class
        std::stable_sort(groups.begin(), groups.end(), [](const Group &a, const Group &b) { return a.count < b.count; });
        PatchInfo &pi = info[pt];
        pi.n_unique   = (unsigned short)groups.size();
        pi.pad[0] = pi.pad[1] = pi.pad[2] = 0;
        for (int m = 0; m < N_CLASSES; ++m) pi.count[m] = 0;
        for (const Group &g : groups) ++pi.count[g.count == 1 ? 0 : g.count == 2 ? 1 : g.count == 4 ? 2 : 3];
        uint32_t *u = uidx.data() + pt * u_stride;
        uint16_t *l = lidx.data() + pt * (size_t)n2 * 32;
        uint16_t *e = ent.data() + pt * (size_t)ent_stride;
        size_t gi = 0;
        int ebase = 0;
        for (int ci = 0; ci < N_CLASSES; ++ci)
          {
            const int m = 1 << ci, cnt = pi.count[ci];
            for (int i = 0; i < cnt; ++i, ++gi)
              {
                u[gi] = groups[gi].g;
                for (int q = 0; q < m; ++q)
                  {
                    const Entry &en        = entries[groups[gi].first + q];
                    e[ebase + q * cnt + i] = en.slot;
                    l[en.lpos]             = (uint16_t)gi;
                  }
              }
            ebase += m * cnt;
          }
        for (size_t i = groups.size(); i < (size_t)u_stride; ++i) u[i] = groups.empty() ? 0u : groups.back().g; // padding: valid address
        total += (long long)groups.size();
      }
  }
  unique_per_cell = n_cells > 0 ? (double)total / (double)n_cells : 0;
  index_bytes     = total * 4 + n_patches * (long long)n2 * 64 + n_cells * (long long)n3 * 2 + n_patches * (long long)sizeof(PatchInfo);
  d_patches       = to_device(info);
  d_uidx          = to_device(uidx);
  d_lidx          = to_device(lidx);
  d_ent           = to_device(ent);
}

struct Operator
{
  int degree = 0, number = 0, device = 0, geometry_type = 0;
  int apply_constraints = 1, kernel = MFHN_KERNEL_AUTO;
  long long n_cells = 0, n_owned = 0, n_ghost = 0, n_cells_hn = 0;
  uint32_t *d_idx = nullptr;    // reference layout [cell][lexicographic]
  uint8_t *d_masks = nullptr;
  void *d_geom = nullptr;       // Number h[cell] or Number G[cell][6]
  PlaneLayout plane;            // warp-interleaved layout of the register-tiled kernel
  PatchLayout patch;            // sorted-unique / CSR layout of the patch kernel
  BulkLayout bulk;              // block descriptors of the bulk-copy kernel (degrees 3..5)
  // padded deal.II-CUDA-style arrays of the baseline kernel (built on first use: 84 B per padded slot)
  uint32_t *d_base_l2g = nullptr;
  void *d_base_invjac = nullptr, *d_base_jxw = nullptr;
  int base_pad = 0;
  std::vector<long long> segments;
  long long launches = 0;
  void *d_stage_src[2] = {nullptr, nullptr}, *d_stage_dst[2] = {nullptr, nullptr}; // device staging of the host-vector entry point (2 slots)
  // src vectors bound as linear textures (gathers through the TEX pipe), cached per pointer
  int use_texture = 0; // measured: no gain over plain loads (profiles/), kept as a switch (MFHN_TEXTURE=1)
  std::vector<std::pair<const void *, cudaTextureObject_t>> tex_cache;
  cudaTextureObject_t texture_for(const void *src)
  {
    if (!use_texture) return 0;
    for (auto &e : tex_cache)
      if (e.first == src) return e.second;
    const long long nvec = n_owned + n_ghost;
    cudaResourceDesc rd{};
    rd.resType                = cudaResourceTypeLinear;
    rd.res.linear.devPtr      = const_cast<void *>(src);
    rd.res.linear.desc        = number == MFHN_F64 ? cudaCreateChannelDesc<int2>() : cudaCreateChannelDesc<float>();
    rd.res.linear.sizeInBytes = (size_t)nvec * (number == MFHN_F64 ? 8 : 4);
    cudaTextureDesc td{};
    td.readMode = cudaReadModeElementType;
    cudaTextureObject_t t = 0;
    if (cudaCreateTextureObject(&t, &rd, &td, nullptr) != cudaSuccess)
      {
        cudaGetLastError(); // vector too long for a linear texture: fall back to plain loads
        use_texture = 0;
        return 0;
      }
    if (tex_cache.size() >= 16)
      {
        cudaDestroyTextureObject(tex_cache.front().second);
        tex_cache.erase(tex_cache.begin());
      }
    tex_cache.emplace_back(src, t);
    return t;
  }

  ~Operator()
  {
    cudaFree(d_idx);
    cudaFree(d_masks);
    cudaFree(d_geom);
    for (int i = 0; i < 2; ++i)
      {
        cudaFree(d_stage_src[i]);
        cudaFree(d_stage_dst[i]);
      }
    for (auto &e : tex_cache) cudaDestroyTextureObject(e.second);
    cudaFree(d_base_l2g);
    cudaFree(d_base_invjac);
    cudaFree(d_base_jxw);
    plane.free();
    patch.free();
    bulk.free();
  }
};

// ---------------------------------------------------------------------------
template <int n, typename Number, int V, bool DIAG = false>
void launch_generic(const Operator &op, const CellLoopParams &p, cudaStream_t stream)
{
  using Cfg           = GenericCfg<n>;
  const size_t smem   = (size_t)Cfg::cpb * generic_n_arrays<V>() * Cfg::cs * sizeof(Number);
  static bool attr[64] = {};
  if (smem > 48 * 1024 && !attr[op.device])
    {
      CUDA_CHECK(cudaFuncSetAttribute(generic_cell_kernel<n, Number, V, DIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr[op.device] = true;
    }
  const long long nc = p.cell_end - p.cell_begin;
  if (nc <= 0) return;
  const unsigned grid = (unsigned)((nc + Cfg::cpb - 1) / Cfg::cpb);
  generic_cell_kernel<n, Number, V, DIAG><<<grid, Cfg::threads, smem, stream>>>(p);
  CUDA_CHECK(cudaGetLastError());
}

// diagonal of the operator (for point-Jacobi): Cartesian cells through the separable form, affine through the q-point form
template <int n, typename Number>
void launch_diagonal_n(Operator &op, const CellLoopParams &p, cudaStream_t stream)
{
  if (op.geometry_type == MFHN_GEOM_AFFINE)
    launch_generic<n, Number, GV_QPOINT_METRIC, true>(op, p, stream);
  else if (op.geometry_type == MFHN_GEOM_GENERAL)
    launch_generic<n, Number, GV_QPOINT_GENERAL, true>(op, p, stream);
  else
    launch_generic<n, Number, GV_SEPARABLE, true>(op, p, stream);
  ++op.launches;
}
template <typename Number>
void launch_diagonal(Operator &op, const CellLoopParams &p, cudaStream_t stream)
{
  switch (op.degree)
    {
      case 1: launch_diagonal_n<2, Number>(op, p, stream); break;
      case 2: launch_diagonal_n<3, Number>(op, p, stream); break;
      case 3: launch_diagonal_n<4, Number>(op, p, stream); break;
      case 4: launch_diagonal_n<5, Number>(op, p, stream); break;
      case 5: launch_diagonal_n<6, Number>(op, p, stream); break;
      case 6: launch_diagonal_n<7, Number>(op, p, stream); break;
      case 7: launch_diagonal_n<8, Number>(op, p, stream); break;
      case 8: launch_diagonal_n<9, Number>(op, p, stream); break;
      default: throw InvalidArgument("unsupported degree");
    }
}

template <int n, typename Number>
void launch_baseline(Operator &op, const CellLoopParams &p, cudaStream_t stream)
{
  using Cfg = BaselineCfg<n>;
  if (!op.d_base_l2g)
    {
      const size_t slots = (size_t)std::max<long long>(op.n_cells, 1) * Cfg::pad;
      CUDA_CHECK(cudaMalloc(&op.d_base_l2g, slots * sizeof(uint32_t)));
      CUDA_CHECK(cudaMalloc(&op.d_base_invjac, slots * 9 * sizeof(Number)));
      CUDA_CHECK(cudaMalloc(&op.d_base_jxw, slots * sizeof(Number)));
      op.base_pad = Cfg::pad;
      if (op.n_cells > 0)
        baseline_setup_kernel<n, Number><<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(
          op.d_base_l2g, (Number *)op.d_base_invjac, (Number *)op.d_base_jxw, op.d_idx, (const Number *)op.d_geom, op.n_cells, Cfg::pad);
      CUDA_CHECK(cudaGetLastError());
    }
  BaselineParams b;
  b.local_to_global   = op.d_base_l2g;
  b.inv_jacobian      = op.d_base_invjac;
  b.JxW               = op.d_base_jxw;
  b.masks             = p.masks;
  b.src               = p.src;
  b.dst               = p.dst;
  b.n_cells           = op.n_cells;
  b.cell_begin        = p.cell_begin;
  b.cell_end          = p.cell_end;
  b.pad               = Cfg::pad;
  b.apply_constraints = p.apply_constraints;
  const long long nc  = p.cell_end - p.cell_begin;
  if (nc <= 0) return;
  baseline_kernel<n, Number><<<(unsigned)((nc + Cfg::cpb - 1) / Cfg::cpb), Cfg::n3 * Cfg::cpb, 0, stream>>>(b);
  CUDA_CHECK(cudaGetLastError());
}

template <int n, typename Number>
void launch_n(Operator &op, int kernel, const CellLoopParams &p, cudaStream_t stream)
{
  if (kernel == MFHN_KERNEL_BASELINE)
    {
      launch_baseline<n, Number>(op, p, stream);
      ++op.launches;
      return;
    }
  if (kernel == MFHN_KERNEL_BULK)
    {
      // the bulk-copy kernel takes whole warp batches; the (at most cpw - 1) cells at either end of an unaligned
      // range and the cells that do not show the block pattern (at most bulk_max_irregular) go to the plane kernel
      constexpr long long cpw = PlaneCfg<n, Number>::cpw;
      // (the last batch of the mesh may be incomplete: its missing cells are marked in the layout)
      const long long b0 = (p.cell_begin + cpw - 1) / cpw * cpw;
      const long long b1 = p.cell_end == op.n_cells ? (op.n_cells + cpw - 1) / cpw * cpw : p.cell_end / cpw * cpw;
      auto plane_part = [&](const long long cb, const long long ce) {
        if (ce <= cb) return;
        CellLoopParams q = p;
        q.cell_begin     = cb;
        q.cell_end       = ce;
        launch_plane<n, Number>(op.plane, q, op.device, stream, 0);
        ++op.launches;
      };
      if (b1 <= b0)
        {
          plane_part(p.cell_begin, p.cell_end);
          return;
        }
      plane_part(p.cell_begin, b0);
      {
        CellLoopParams q = p;
        q.cell_begin     = b0;
        q.cell_end       = b1;
        launch_bulk<n, Number>(op.bulk, q, op.device, stream);
      }
      plane_part(b1, p.cell_end);
      for (const long long c : op.bulk.irregular)
        if (c >= b0 && c < std::min(b1, p.cell_end)) plane_part(c, c + 1);
    }
  else if (kernel == MFHN_KERNEL_PATCH)
    launch_patch<n, Number>(op.patch, p, op.device, stream);
  else if (kernel == MFHN_KERNEL_PLANE && plane_supported(n))
    launch_plane<n, Number>(op.plane, p, op.device, stream, op.texture_for(p.src));
  else if (kernel == MFHN_KERNEL_PLANE) // degrees 6..8: plane in shared memory
    launch_plane_smem<n, Number>(op.plane, p, op.device, stream);
  else if (op.geometry_type == MFHN_GEOM_AFFINE)
    launch_generic<n, Number, GV_QPOINT_METRIC>(op, p, stream);
  else if (op.geometry_type == MFHN_GEOM_GENERAL)
    launch_generic<n, Number, GV_QPOINT_GENERAL>(op, p, stream);
  else if (kernel == MFHN_KERNEL_SEPARABLE)
    launch_generic<n, Number, GV_SEPARABLE>(op, p, stream);
  else
    launch_generic<n, Number, GV_QPOINT_CARTESIAN>(op, p, stream);
  ++op.launches;
}

template <typename Number>
void launch_number(Operator &op, int kernel, const CellLoopParams &p, cudaStream_t stream)
{
  switch (op.degree)
    {
      case 1: launch_n<2, Number>(op, kernel, p, stream); break;
      case 2: launch_n<3, Number>(op, kernel, p, stream); break;
      case 3: launch_n<4, Number>(op, kernel, p, stream); break;
      case 4: launch_n<5, Number>(op, kernel, p, stream); break;
      case 5: launch_n<6, Number>(op, kernel, p, stream); break;
      case 6: launch_n<7, Number>(op, kernel, p, stream); break;
      case 7: launch_n<8, Number>(op, kernel, p, stream); break;
      case 8: launch_n<9, Number>(op, kernel, p, stream); break;
      default: throw InvalidArgument("unsupported degree");
    }
}

int resolve_kernel(const Operator &op)
{
  int kernel = op.kernel;
  if (kernel == MFHN_KERNEL_AUTO)
    {
      kernel = op.geometry_type == MFHN_GEOM_CARTESIAN ? MFHN_KERNEL_PLANE : MFHN_KERNEL_QPOINT;
      // measured on B200 (profiles/): the bulk-copy kernel wins for double at k = 4, 5 (the plane kernel is bound by
      // the L1 data stage there); for float and k = 3 the plane kernel's gathers are cheap enough
      if (kernel == MFHN_KERNEL_PLANE && op.number == MFHN_F64 && (op.degree == 4 || op.degree == 5) && op.bulk.usable) kernel = MFHN_KERNEL_BULK;
    }
  if (op.geometry_type != MFHN_GEOM_CARTESIAN && kernel != MFHN_KERNEL_QPOINT)
    throw InvalidArgument("affine / general geometry requires MFHN_KERNEL_QPOINT");
  if (kernel == MFHN_KERNEL_BULK && !bulk_supported(op.degree + 1)) throw NotImplemented("MFHN_KERNEL_BULK is available for degrees 3..5");
  if (kernel == MFHN_KERNEL_BULK && op.geometry_type == MFHN_GEOM_CARTESIAN && !op.bulk.usable)
    throw InvalidArgument("MFHN_KERNEL_BULK: the DoF numbering does not show contiguous cell-interior / face blocks");
  if (kernel == MFHN_KERNEL_PATCH && !plane_supported(op.degree + 1))
    throw NotImplemented("MFHN_KERNEL_PATCH is not available for this degree");
  if (kernel == MFHN_KERNEL_PATCH && op.patch.d_uidx == nullptr)
    throw InvalidArgument("the patch layout is built only when the operator is created with MFHN_KERNEL_PATCH");
  if ((kernel == MFHN_KERNEL_PLANE || kernel == MFHN_KERNEL_PATCH || kernel == MFHN_KERNEL_BULK || kernel == MFHN_KERNEL_SEPARABLE) && op.geometry_type != MFHN_GEOM_CARTESIAN)
    throw InvalidArgument("this kernel requires Cartesian geometry");
  if (kernel == MFHN_KERNEL_BASELINE && op.geometry_type != MFHN_GEOM_CARTESIAN)
    throw InvalidArgument("the baseline kernel is set up for Cartesian cells");
  return kernel;
}

void op_vmult_range(Operator &op, void *dst, const void *src, cudaStream_t stream, long long cb, long long ce)
{
  if (ce < 0) ce = op.n_cells;
  if (cb < 0 || ce > op.n_cells || cb > ce) throw InvalidArgument("cell range out of bounds");
  CellLoopParams p;
  p.idx               = op.d_idx;
  p.masks             = op.d_masks;
  p.geom              = op.d_geom;
  p.src               = src;
  p.dst               = dst;
  p.cell_begin        = cb;
  p.cell_end          = ce;
  p.apply_constraints = op.apply_constraints;
  int kernel          = resolve_kernel(op);
  if (kernel == MFHN_KERNEL_BULK && (((uintptr_t)src | (uintptr_t)dst) & 15u))
    {
      if (op.kernel != MFHN_KERNEL_AUTO)
        throw InvalidArgument("MFHN_KERNEL_BULK needs 16-byte aligned vectors (use MFHN_KERNEL_PLANE for unaligned views)");
      kernel = MFHN_KERNEL_PLANE; // AUTO: unaligned views take the plane kernel
    }
  if (op.number == MFHN_F64)
    launch_number<double>(op, kernel, p, stream);
  else
    launch_number<float>(op, kernel, p, stream);
}

Operator *op_create(const mfhn_op_desc &d)
{
  if (d.degree < 1 || d.degree > 8) throw InvalidArgument("degree must be in 1..8");
  if (d.number != MFHN_F64 && d.number != MFHN_F32) throw InvalidArgument("number must be MFHN_F64 or MFHN_F32");
  if (d.n_cells < 0 || d.n_owned < 0 || d.n_ghost < 0) throw InvalidArgument("negative size");
  if (d.n_cells > 0 && (!d.dof_indices || !d.masks || !d.geometry)) throw InvalidArgument("null array");
  if (d.geometry_type != MFHN_GEOM_CARTESIAN && d.geometry_type != MFHN_GEOM_AFFINE && d.geometry_type != MFHN_GEOM_GENERAL)
    throw InvalidArgument("unknown geometry type");
  if (d.kernel < MFHN_KERNEL_AUTO || d.kernel > MFHN_KERNEL_BULK) throw InvalidArgument("unknown kernel");
  int device = d.device;
  if (device < 0)
    CUDA_CHECK(cudaGetDevice(&device));
  else
    CUDA_CHECK(cudaSetDevice(device));
  upload_tables(device);
  std::unique_ptr<Operator> op(new Operator);
  op->degree            = d.degree;
  op->number            = d.number;
  op->device            = device;
  op->geometry_type     = d.geometry_type;
  op->apply_constraints = d.apply_constraints;
  if (const char *e = std::getenv("MFHN_TEXTURE")) op->use_texture = std::atoi(e);
  op->kernel            = d.kernel;
  op->n_cells           = d.n_cells;
  op->n_owned           = d.n_owned;
  op->n_ghost           = d.n_ghost;
  const int n           = d.degree + 1;
  const long long n3    = (long long)n * n * n;
  const long long nvec  = d.n_owned + d.n_ghost;
  for (long long i = 0; i < d.n_cells * n3; ++i)
    if (d.dof_indices[i] >= (unsigned long long)nvec) throw InvalidArgument("dof index out of range");
  for (long long c = 0; c < d.n_cells; ++c)
    {
      if (!check_kind(decompress_kind(d.masks[c])) || compress_kind(decompress_kind(d.masks[c])) != d.masks[c])
        throw InvalidArgument("invalid compressed constraint mask");
      op->n_cells_hn += d.masks[c] != 0;
    }
  CUDA_CHECK(cudaMalloc(&op->d_idx, std::max<size_t>(1, d.n_cells * n3) * sizeof(uint32_t)));
  CUDA_CHECK(cudaMemcpy(op->d_idx, d.dof_indices, d.n_cells * n3 * sizeof(uint32_t), cudaMemcpyHostToDevice));
  CUDA_CHECK(cudaMalloc(&op->d_masks, std::max<size_t>(1, d.n_cells)));
  CUDA_CHECK(cudaMemcpy(op->d_masks, d.masks, d.n_cells, cudaMemcpyHostToDevice));
  // geometry -> per-cell factors in the operator's Number type
  std::vector<double> g;
  if (d.geometry_type == MFHN_GEOM_CARTESIAN)
    g.assign(d.geometry, d.geometry + d.n_cells);
  else if (d.geometry_type == MFHN_GEOM_GENERAL)
    g.assign(d.geometry, d.geometry + d.n_cells * 6 * n3); // [cell][6][q], JxW J^-1 J^-T per quadrature point
  else
    {
      g.resize(d.n_cells * 6);
      for (long long c = 0; c < d.n_cells; ++c)
        {
          const double *J = d.geometry + 9 * c;
          const double det = J[0] * (J[4] * J[8] - J[5] * J[7]) - J[1] * (J[3] * J[8] - J[5] * J[6]) + J[2] * (J[3] * J[7] - J[4] * J[6]);
          if (!(det > 0)) throw InvalidArgument("non-positive Jacobian determinant");
          double inv[9] = {(J[4] * J[8] - J[5] * J[7]) / det, (J[2] * J[7] - J[1] * J[8]) / det, (J[1] * J[5] - J[2] * J[4]) / det,
                           (J[5] * J[6] - J[3] * J[8]) / det, (J[0] * J[8] - J[2] * J[6]) / det, (J[2] * J[3] - J[0] * J[5]) / det,
                           (J[3] * J[7] - J[4] * J[6]) / det, (J[1] * J[6] - J[0] * J[7]) / det, (J[0] * J[4] - J[1] * J[3]) / det};
          // G = det * inv * inv^T  (symmetric): rows of inv are gradients of xi_r
          int t = 0;
          for (int r = 0; r < 3; ++r)
            for (int s = r; s < 3; ++s)
              g[6 * c + t++] = det * (inv[3 * r] * inv[3 * s] + inv[3 * r + 1] * inv[3 * s + 1] + inv[3 * r + 2] * inv[3 * s + 2]);
        }
    }
  if (d.number == MFHN_F64)
    op->d_geom = to_device(g);
  else
    {
      std::vector<float> gf(g.begin(), g.end());
      op->d_geom = to_device(gf);
    }
  if (d.geometry_type == MFHN_GEOM_CARTESIAN)
    {
      op->plane.build(n, d.n_cells, d.dof_indices);
      op->segments.assign(1, 0);
      if (d.segments)
        {
          if (d.n_segments < 1 || d.segments[0] != 0) throw InvalidArgument("segments must start at cell 0");
          op->segments.assign(d.segments, d.segments + d.n_segments);
          for (int i = 1; i < d.n_segments; ++i)
            if (d.segments[i] < d.segments[i - 1] || d.segments[i] > d.n_cells) throw InvalidArgument("segments must be ascending");
        }
      if (bulk_supported(n)) bulk_build(op->bulk, n, d.number, d.n_cells, nvec, d.dof_indices);
      if (plane_supported(n) && (d.kernel == MFHN_KERNEL_PATCH || std::getenv("MFHN_BUILD_PATCH"))) op->patch.build(n, d.number, d.n_cells, d.dof_indices);
    }
  resolve_kernel(*op);
  return op.release();
}

// ---------------------------------------------------------------------------
// auxiliary kernels
template <typename Number>
__global__ void pack_kernel(Number *buf, const Number *vec, const int32_t *idx, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) buf[i] = vec[idx[i]];
}
template <typename Number>
__global__ void unpack_add_kernel(Number *vec, const Number *buf, const int32_t *idx, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) vec[idx[i]] += buf[i];
}

template <int n, typename Number>
__global__ void hn_only_kernel(Number *values, const uint8_t *masks, long long n_cells, int transpose)
{
  // FEEvaluationHangingNodesFactory::apply on cell-local values (benchmark_00_likwid.cc:56-59)
  __shared__ Number s[n * n * n];
  const long long cell = blockIdx.x;
  if (masks[cell] == 0) return; // unconstrained cell: nothing to interpolate (block-uniform)
  const int l = threadIdx.x, a = l % n, b = l / n;
  Number *g = values + cell * (n * n * n);
  for (int z = 0; z < n; ++z) s[l + n * n * z] = g[l + n * n * z];
  unsigned face, edge, cb;
  const unsigned mask = masks[cell];
  decode_mask(mask, face, edge, cb);
  __syncthreads();
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
  for (int z = 0; z < n; ++z) g[l + n * n * z] = s[l + n * n * z];
}

template <typename Number>
void hn_only(Operator &op, void *values, int transpose, cudaStream_t st)
{
  if (op.n_cells == 0) return;
  const unsigned grid = (unsigned)op.n_cells;
#define HN_CASE(N)                                                                                            \
  case N - 1:                                                                                                 \
    hn_only_kernel<N, Number><<<grid, N * N, 0, st>>>((Number *)values, op.d_masks, op.n_cells, transpose); \
    break;
  switch (op.degree)
    {
      HN_CASE(2) HN_CASE(3) HN_CASE(4) HN_CASE(5) HN_CASE(6) HN_CASE(7) HN_CASE(8) HN_CASE(9)
    }
#undef HN_CASE
  CUDA_CHECK(cudaGetLastError());
  ++op.launches;
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

// ---------------------------------------------------------------------------
// partitioned operator: ghost import / compress over NCCL, overlapped with the cell partitions
struct Dist
{
  Operator *op = nullptr;
  void *comm   = nullptr;
  int rank = 0, world = 1;
  std::vector<int> import_peers, ghost_peers;
  std::vector<long long> import_off, ghost_begin, ghost_end;
  long long n_import = 0;
  long long seg[4]   = {0, 0, 0, 0};
  int32_t *d_import_idx = nullptr;
  void *d_send = nullptr, *d_recv = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev[4]        = {nullptr, nullptr, nullptr, nullptr};
  long long launches       = 0;
  // peer mode: registered vector pair + per-ghost addresses inside the owners' vectors
  void *peer_src_local = nullptr, *peer_dst_local = nullptr;
  void **d_ghost_src = nullptr, **d_ghost_dst = nullptr;
  int *d_barrier = nullptr;

  ~Dist()
  {
    cudaFree(d_ghost_src);
    cudaFree(d_ghost_dst);
    cudaFree(d_barrier);
    if (comm) NcclApi::get().CommDestroy(comm);
    cudaFree(d_import_idx);
    cudaFree(d_send);
    cudaFree(d_recv);
    if (comm_stream) cudaStreamDestroy(comm_stream);
    for (auto e : ev)
      if (e) cudaEventDestroy(e);
  }
};

#define NCCL_CHECK(x)                                                                            \
  do                                                                                             \
    {                                                                                            \
      int r_ = (x);                                                                              \
      if (r_ != 0) throw CudaError(std::string(#x) + ": " + NcclApi::get().GetErrorString(r_)); \
    }                                                                                            \
  while (0)

void dist_vmult(Dist &d, void *dst, const void *src, cudaStream_t main, int zero_dst)
{
  Operator &op  = *d.op;
  NcclApi &nccl = NcclApi::get();
  const size_t s = op.number == MFHN_F64 ? 8 : 4;
  char *dstb = static_cast<char *>(dst);
  char *srcb = static_cast<char *>(const_cast<void *>(src));
  const unsigned grid = (unsigned)((d.n_import + 255) / 256);
  if (zero_dst) CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)(op.n_owned + op.n_ghost) * s, main));
  // pack the entries the peers ghost
  if (d.n_import > 0)
    {
      if (op.number == MFHN_F64)
        pack_all_kernel<double><<<grid, 256, 0, main>>>((double *)d.d_send, (const double *)src, d.d_import_idx, d.n_import);
      else
        pack_all_kernel<float><<<grid, 256, 0, main>>>((float *)d.d_send, (const float *)src, d.d_import_idx, d.n_import);
      CUDA_CHECK(cudaGetLastError());
      ++d.launches;
    }
  CUDA_CHECK(cudaEventRecord(d.ev[0], main));
  CUDA_CHECK(cudaStreamWaitEvent(d.comm_stream, d.ev[0], 0));
  // owners -> ghosts (update_ghost_values)
  NCCL_CHECK(nccl.GroupStart());
  for (size_t i = 0; i < d.ghost_peers.size(); ++i)
    NCCL_CHECK(nccl.Recv(srcb + (size_t)(op.n_owned + d.ghost_begin[i]) * s, (size_t)(d.ghost_end[i] - d.ghost_begin[i]) * s, 0,
                         d.ghost_peers[i], d.comm, d.comm_stream));
  for (size_t i = 0; i < d.import_peers.size(); ++i)
    NCCL_CHECK(nccl.Send(static_cast<char *>(d.d_send) + (size_t)d.import_off[i] * s, (size_t)(d.import_off[i + 1] - d.import_off[i]) * s, 0,
                         d.import_peers[i], d.comm, d.comm_stream));
  NCCL_CHECK(nccl.GroupEnd());
  CUDA_CHECK(cudaEventRecord(d.ev[1], d.comm_stream));
  if (d.seg[1] > d.seg[0]) op_vmult_range(op, dst, src, main, d.seg[0], d.seg[1]); // interior A overlaps the import
  CUDA_CHECK(cudaStreamWaitEvent(main, d.ev[1], 0));
  if (d.seg[3] > d.seg[2]) op_vmult_range(op, dst, src, main, d.seg[2], d.seg[3]); // boundary cells need the ghosts
  CUDA_CHECK(cudaEventRecord(d.ev[2], main));
  CUDA_CHECK(cudaStreamWaitEvent(d.comm_stream, d.ev[2], 0));
  // ghosts -> owners (compress, add)
  NCCL_CHECK(nccl.GroupStart());
  for (size_t i = 0; i < d.import_peers.size(); ++i)
    NCCL_CHECK(nccl.Recv(static_cast<char *>(d.d_recv) + (size_t)d.import_off[i] * s, (size_t)(d.import_off[i + 1] - d.import_off[i]) * s, 0,
                         d.import_peers[i], d.comm, d.comm_stream));
  for (size_t i = 0; i < d.ghost_peers.size(); ++i)
    NCCL_CHECK(nccl.Send(dstb + (size_t)(op.n_owned + d.ghost_begin[i]) * s, (size_t)(d.ghost_end[i] - d.ghost_begin[i]) * s, 0,
                         d.ghost_peers[i], d.comm, d.comm_stream));
  NCCL_CHECK(nccl.GroupEnd());
  CUDA_CHECK(cudaEventRecord(d.ev[3], d.comm_stream));
  if (d.seg[2] > d.seg[1]) op_vmult_range(op, dst, src, main, d.seg[1], d.seg[2]); // interior B overlaps the compress
  CUDA_CHECK(cudaStreamWaitEvent(main, d.ev[3], 0));
  if (d.n_import > 0)
    {
      if (op.number == MFHN_F64)
        unpack_add_all_kernel<double><<<grid, 256, 0, main>>>((double *)dst, (const double *)d.d_recv, d.d_import_idx, d.n_import);
      else
        unpack_add_all_kernel<float><<<grid, 256, 0, main>>>((float *)dst, (const float *)d.d_recv, d.d_import_idx, d.n_import);
      CUDA_CHECK(cudaGetLastError());
      ++d.launches;
    }
  if (op.n_ghost > 0) CUDA_CHECK(cudaMemsetAsync(dstb + (size_t)op.n_owned * s, 0, (size_t)op.n_ghost * s, main));
}

// Peer-memory variant: no pack / unpack and no data-path collective.  After a barrier the
// boundary cells read the owners' src entries and add into the owners' dst entries directly
// over NVLink; a second barrier makes the remote contributions visible before anyone consumes
// dst.  The barriers are 4-byte NCCL all-reduces.
template <typename Number>
void dist_vmult_peer_n(Dist &d, cudaStream_t main)
{
  Operator &op = *d.op;
  CellLoopParams p;
  p.idx               = op.d_idx;
  p.masks             = op.d_masks;
  p.geom              = op.d_geom;
  p.src               = d.peer_src_local;
  p.dst               = d.peer_dst_local;
  p.apply_constraints = op.apply_constraints;
  PeerTables pt;
  pt.n_owned   = op.n_owned;
  pt.ghost_src = d.d_ghost_src;
  pt.ghost_dst = d.d_ghost_dst;
  NcclApi &nccl = NcclApi::get();
  auto launch = [&](long long cb, long long ce, const PeerTables *peer, cudaStream_t st) {
    if (ce <= cb) return;
    p.cell_begin = cb;
    p.cell_end   = ce;
    switch (op.degree)
      {
        case 1: launch_plane<2, Number>(op.plane, p, op.device, st, 0, peer); break;
        case 2: launch_plane<3, Number>(op.plane, p, op.device, st, 0, peer); break;
        case 3: launch_plane<4, Number>(op.plane, p, op.device, st, 0, peer); break;
        case 4: launch_plane<5, Number>(op.plane, p, op.device, st, 0, peer); break;
        case 5: launch_plane<6, Number>(op.plane, p, op.device, st, 0, peer); break;
        default: throw NotImplemented("peer mode covers the register-tiled plane kernel (degree <= 5)");
      }
    ++op.launches;
    ++d.launches;
  };
  // high-priority stream: barrier (peers' src final, peers' dst zeroed) -> boundary cells with remote
  // entries over NVLink -> barrier (remote contributions landed); the interior cells (local entries
  // only) run concurrently on the compute stream and hide the whole chain
  CUDA_CHECK(cudaEventRecord(d.ev[0], main));
  CUDA_CHECK(cudaStreamWaitEvent(d.comm_stream, d.ev[0], 0));
  NCCL_CHECK(nccl.AllReduce(d.d_barrier, d.d_barrier, 1, 2 /*ncclInt32*/, 0 /*ncclSum*/, d.comm, d.comm_stream));
  launch(d.seg[2], d.seg[3], &pt, d.comm_stream);
  NCCL_CHECK(nccl.AllReduce(d.d_barrier, d.d_barrier, 1, 2, 0, d.comm, d.comm_stream));
  CUDA_CHECK(cudaEventRecord(d.ev[3], d.comm_stream));
  launch(d.seg[0], d.seg[2], nullptr, main);
  CUDA_CHECK(cudaStreamWaitEvent(main, d.ev[3], 0));
}
} // namespace mfhn

using namespace mfhn;

extern "C" {
int mfhn_op_create(const mfhn_op_desc *desc, mfhn_op *out)
{
  return guard([&] {
    if (!desc || !out) throw InvalidArgument("null argument");
    *out = reinterpret_cast<mfhn_op>(op_create(*desc));
  });
}
void mfhn_op_destroy(mfhn_op op) { delete reinterpret_cast<Operator *>(op); }

int mfhn_op_vmult(mfhn_op h, void *dst, const void *src, void *stream, int zero_dst)
{
  return guard([&] {
    if (!h || !dst || !src) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (zero_dst)
      CUDA_CHECK(cudaMemsetAsync(dst, 0, (size_t)(op.n_owned + op.n_ghost) * (op.number == MFHN_F64 ? 8 : 4), st));
    op_vmult_range(op, dst, src, st, 0, op.n_cells);
  });
}
int mfhn_op_vmult_range(mfhn_op h, void *dst, const void *src, void *stream, int64_t cb, int64_t ce)
{
  return guard([&] {
    if (!h || !dst || !src) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    op_vmult_range(op, dst, src, static_cast<cudaStream_t>(stream), cb, ce);
  });
}
int mfhn_op_vmult_host_slot(mfhn_op h, void *dst_host, const void *src_host, void *stream, int zero_dst, int slot)
{
  return guard([&] {
    if (!h || !dst_host || !src_host) throw InvalidArgument("null argument");
    if (slot < 0 || slot > 1) throw InvalidArgument("slot must be 0 or 1");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t st    = static_cast<cudaStream_t>(stream);
    const size_t bytes = (size_t)(op.n_owned + op.n_ghost) * (op.number == MFHN_F64 ? 8 : 4);
    if (!op.d_stage_src[slot])
      {
        CUDA_CHECK(cudaMalloc(&op.d_stage_src[slot], std::max<size_t>(bytes, 8)));
        CUDA_CHECK(cudaMalloc(&op.d_stage_dst[slot], std::max<size_t>(bytes, 8)));
      }
    CUDA_CHECK(cudaMemcpyAsync(op.d_stage_src[slot], src_host, bytes, cudaMemcpyHostToDevice, st));
    if (zero_dst)
      CUDA_CHECK(cudaMemsetAsync(op.d_stage_dst[slot], 0, bytes, st));
    else
      CUDA_CHECK(cudaMemcpyAsync(op.d_stage_dst[slot], dst_host, bytes, cudaMemcpyHostToDevice, st));
    op_vmult_range(op, op.d_stage_dst[slot], op.d_stage_src[slot], st, 0, op.n_cells);
    CUDA_CHECK(cudaMemcpyAsync(dst_host, op.d_stage_dst[slot], bytes, cudaMemcpyDeviceToHost, st));
  });
}
int mfhn_op_vmult_host(mfhn_op h, void *dst_host, const void *src_host, void *stream, int zero_dst)
{
  return mfhn_op_vmult_host_slot(h, dst_host, src_host, stream, zero_dst, 0);
}
int mfhn_op_diagonal(mfhn_op h, void *diag, void *stream)
{
  return guard([&] {
    if (!h || !diag) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    CellLoopParams p;
    p.idx               = op.d_idx;
    p.masks             = op.d_masks;
    p.geom              = op.d_geom;
    p.src               = diag; // unused
    p.dst               = diag;
    p.cell_begin        = 0;
    p.cell_end          = op.n_cells;
    p.apply_constraints = op.apply_constraints;
    if (op.number == MFHN_F64)
      launch_diagonal<double>(op, p, static_cast<cudaStream_t>(stream));
    else
      launch_diagonal<float>(op, p, static_cast<cudaStream_t>(stream));
  });
}
int mfhn_op_set_apply_constraints(mfhn_op h, int v)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    reinterpret_cast<Operator *>(h)->apply_constraints = v != 0;
  });
}
int mfhn_op_set_kernel(mfhn_op h, int kernel)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    Operator &op  = *reinterpret_cast<Operator *>(h);
    const int old = op.kernel;
    if (kernel < MFHN_KERNEL_AUTO || kernel > MFHN_KERNEL_BULK) throw InvalidArgument("unknown kernel");
    op.kernel = kernel;
    try
      {
        resolve_kernel(op);
      }
    catch (...)
      {
        op.kernel = old;
        throw;
      }
  });
}
int mfhn_op_apply_hn(mfhn_op h, void *values, int transpose, void *stream)
{
  return guard([&] {
    if (!h || !values) throw InvalidArgument("null argument");
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    if (op.number == MFHN_F64)
      hn_only<double>(op, values, transpose, static_cast<cudaStream_t>(stream));
    else
      hn_only<float>(op, values, transpose, static_cast<cudaStream_t>(stream));
  });
}
int mfhn_op_query(mfhn_op h, const char *what, double *value)
{
  return guard([&] {
    if (!h || !what || !value) throw InvalidArgument("null argument");
    const Operator &op = *reinterpret_cast<Operator *>(h);
    const double n = op.degree + 1, n3 = n * n * n, s = op.number == MFHN_F64 ? 8 : 4;
    const double nvec = (double)(op.n_owned + op.n_ghost);
    const std::string w(what);
    if (w == "n_cells")
      *value = (double)op.n_cells;
    else if (w == "n_cells_hn")
      *value = (double)op.n_cells_hn;
    else if (w == "algorithmic_bytes") // 2 s n_dofs + n_cells (4 (k+1)^3 + 1 + G)
      *value = 2 * s * nvec + (double)op.n_cells * (4 * n3 + 1 + (op.geometry_type == MFHN_GEOM_CARTESIAN ? 3 * s : op.geometry_type == MFHN_GEOM_AFFINE ? 10 * s : 6 * s * n3));
    else if (w == "algorithmic_bytes_accumulate") // + s n_dofs: dst is read as well when vmult accumulates
      *value = 3 * s * nvec + (double)op.n_cells * (4 * n3 + 1 + (op.geometry_type == MFHN_GEOM_CARTESIAN ? 3 * s : op.geometry_type == MFHN_GEOM_AFFINE ? 10 * s : 6 * s * n3));
    else if (w == "algorithmic_flops") // even-odd sum factorisation count of SURVEY 8d, without HN terms
      *value = (double)op.n_cells * (12 * n * n * (n * n + 2 * n) + 3 * n3);
    else if (w == "unique_dofs_per_cell")
      *value = op.patch.unique_per_cell;
    else if (w == "patch_index_bytes")
      *value = (double)op.patch.index_bytes;
    else if (w == "kernel")
      *value = (double)resolve_kernel(op);
    else if (w == "bulk_irregular_cells") // cells the bulk-copy kernel leaves to the plane kernel (-1: layout not usable)
      *value = op.bulk.usable ? (double)op.bulk.irregular.size() : -1.0;
    else
      throw InvalidArgument("unknown query '" + w + "'");
  });
}
int64_t mfhn_op_launch_count(mfhn_op h) { return h ? reinterpret_cast<Operator *>(h)->launches : 0; }

int mfhn_bulk_layout_check(int degree, int number, int64_t n_cells, int64_t n_vec, const uint32_t *dof_indices, int64_t *n_irregular,
                           int64_t *n_mismatch)
{
  return guard([&] {
    if (n_cells > 0 && !dof_indices) throw InvalidArgument("null argument");
    if (number != MFHN_F64 && number != MFHN_F32) throw InvalidArgument("number must be MFHN_F64 or MFHN_F32");
    BulkHostLayout L;
    bulk_analyze(L, degree + 1, number, n_cells, n_vec, dof_indices);
    if (n_irregular) *n_irregular = (int64_t)L.irregular.size();
    if (n_mismatch) *n_mismatch = bulk_verify(L, number, dof_indices);
  });
}

int mfhn_pack(int number, void *buffer, const void *vec, const int32_t *idx, int64_t n, void *stream)
{
  return guard([&] {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st     = static_cast<cudaStream_t>(stream);
    if (number == MFHN_F64)
      pack_kernel<double><<<grid, 256, 0, st>>>((double *)buffer, (const double *)vec, idx, n);
    else
      pack_kernel<float><<<grid, 256, 0, st>>>((float *)buffer, (const float *)vec, idx, n);
    CUDA_CHECK(cudaGetLastError());
  });
}
int mfhn_unpack_add(int number, void *vec, const void *buffer, const int32_t *idx, int64_t n, void *stream)
{
  return guard([&] {
    if (n <= 0) return;
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st     = static_cast<cudaStream_t>(stream);
    if (number == MFHN_F64)
      unpack_add_kernel<double><<<grid, 256, 0, st>>>((double *)vec, (const double *)buffer, idx, n);
    else
      unpack_add_kernel<float><<<grid, 256, 0, st>>>((float *)vec, (const float *)buffer, idx, n);
    CUDA_CHECK(cudaGetLastError());
  });
}
int mfhn_bench_dfma(int number, int iters, double *tflops)
{
  return guard([&] {
    if (!tflops) throw InvalidArgument("null argument");
    void *out = nullptr;
    CUDA_CHECK(cudaMalloc(&out, 64));
    cudaDeviceProp prop;
    int dev;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep)
      {
        CUDA_CHECK(cudaEventRecord(e0));
        if (number == MFHN_F64)
          fma_bench_kernel<double><<<blocks, threads>>>((double *)out, iters);
        else
          fma_bench_kernel<float><<<blocks, threads>>>((float *)out, iters);
        CUDA_CHECK(cudaEventRecord(e1));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
      }
    *tflops = 2.0 * 8.0 * iters * (double)blocks * threads / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
  });
}

int mfhn_dist_unique_id(void *id128)
{
  return guard([&] {
    if (!id128) throw InvalidArgument("null argument");
    NcclApi &nccl = NcclApi::get();
    if (!nccl.ok) throw CudaError(nccl.error);
    NCCL_CHECK(nccl.GetUniqueId(id128));
  });
}
int mfhn_dist_create(mfhn_op h, const mfhn_dist_desc *dd, mfhn_dist *out)
{
  return guard([&] {
    if (!h || !dd || !out || !dd->unique_id) throw InvalidArgument("null argument");
    NcclApi &nccl = NcclApi::get();
    if (!nccl.ok) throw CudaError(nccl.error);
    Operator &op = *reinterpret_cast<Operator *>(h);
    CUDA_CHECK(cudaSetDevice(op.device));
    std::unique_ptr<Dist> d(new Dist);
    d->op    = &op;
    d->rank  = dd->rank;
    d->world = dd->world;
    for (int i = 0; i < 4; ++i) d->seg[i] = dd->segments[i];
    if (!(0 == d->seg[0] && d->seg[0] <= d->seg[1] && d->seg[1] <= d->seg[2] && d->seg[2] <= d->seg[3] && d->seg[3] == op.n_cells))
      throw InvalidArgument("segments must be 0 <= a <= b <= n_cells");
    d->import_peers.assign(dd->import_peers, dd->import_peers + dd->n_import_peers);
    d->import_off.assign(dd->import_offsets, dd->import_offsets + dd->n_import_peers + 1);
    d->ghost_peers.assign(dd->ghost_peers, dd->ghost_peers + dd->n_ghost_peers);
    d->ghost_begin.assign(dd->ghost_begin, dd->ghost_begin + dd->n_ghost_peers);
    d->ghost_end.assign(dd->ghost_end, dd->ghost_end + dd->n_ghost_peers);
    d->n_import = d->import_off.back();
    for (long long i = 0; i < d->n_import; ++i)
      if (dd->import_indices[i] < 0 || dd->import_indices[i] >= op.n_owned) throw InvalidArgument("import index out of range");
    for (size_t i = 0; i < d->ghost_peers.size(); ++i)
      if (d->ghost_begin[i] < 0 || d->ghost_end[i] > op.n_ghost || d->ghost_begin[i] > d->ghost_end[i]) throw InvalidArgument("ghost range out of bounds");
    const size_t s = op.number == MFHN_F64 ? 8 : 4;
    CUDA_CHECK(cudaMalloc(&d->d_import_idx, std::max<size_t>(1, d->n_import) * sizeof(int32_t)));
    CUDA_CHECK(cudaMemcpy(d->d_import_idx, dd->import_indices, d->n_import * sizeof(int32_t), cudaMemcpyHostToDevice));
    CUDA_CHECK(cudaMalloc(&d->d_send, std::max<size_t>(8, d->n_import * s)));
    CUDA_CHECK(cudaMalloc(&d->d_recv, std::max<size_t>(8, d->n_import * s)));
    {
      // highest priority: the NCCL kernels must not queue behind the blocks of an interior cell partition
      int lo = 0, hi = 0;
      CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      CUDA_CHECK(cudaStreamCreateWithPriority(&d->comm_stream, cudaStreamNonBlocking, hi));
    }
    for (auto &e : d->ev) CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    NcclApi::Id id;
    std::memcpy(&id, dd->unique_id, sizeof(id));
    NCCL_CHECK(nccl.CommInitRank(&d->comm, dd->world, id, dd->rank));
    *out = reinterpret_cast<mfhn_dist>(d.release());
  });
}
void mfhn_dist_destroy(mfhn_dist d) { delete reinterpret_cast<Dist *>(d); }
int mfhn_dist_vmult(mfhn_dist h, void *dst, const void *src, void *stream, int zero_dst)
{
  return guard([&] {
    if (!h || !dst || !src) throw InvalidArgument("null argument");
    Dist &d = *reinterpret_cast<Dist *>(h);
    CUDA_CHECK(cudaSetDevice(d.op->device));
    dist_vmult(d, dst, src, static_cast<cudaStream_t>(stream), zero_dst);
  });
}
int64_t mfhn_dist_launch_count(mfhn_dist h) { return h ? reinterpret_cast<Dist *>(h)->launches : 0; }

int mfhn_vec_alloc(int64_t bytes, void **ptr)
{
  return guard([&] {
    if (!ptr || bytes < 0) throw InvalidArgument("bad argument");
    CUDA_CHECK(cudaMalloc(ptr, std::max<size_t>((size_t)bytes, 8)));
    CUDA_CHECK(cudaMemset(*ptr, 0, std::max<size_t>((size_t)bytes, 8)));
  });
}
int mfhn_vec_free(void *ptr)
{
  return guard([&] { CUDA_CHECK(cudaFree(ptr)); });
}
int mfhn_ipc_get_handle(void *ptr, void *handle64)
{
  return guard([&] {
    if (!ptr || !handle64) throw InvalidArgument("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CUDA_CHECK(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t *>(handle64), ptr));
  });
}
int mfhn_ipc_open_handle(const void *handle64, void **ptr)
{
  return guard([&] {
    if (!ptr || !handle64) throw InvalidArgument("null argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, sizeof(h));
    CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  });
}
int mfhn_ipc_close_handle(void *ptr)
{
  return guard([&] { CUDA_CHECK(cudaIpcCloseMemHandle(ptr)); });
}
int mfhn_dist_enable_peer(mfhn_dist h, void *src_local, void *dst_local, void *const *peer_src, void *const *peer_dst,
                          const int32_t *ghost_owner, const int64_t *ghost_remote_index)
{
  return guard([&] {
    if (!h || !src_local || !dst_local || !peer_src || !peer_dst) throw InvalidArgument("null argument");
    Dist &d      = *reinterpret_cast<Dist *>(h);
    Operator &op = *d.op;
    if (!plane_supported(op.degree + 1) || op.geometry_type != MFHN_GEOM_CARTESIAN)
      throw NotImplemented("peer mode covers the register-tiled plane kernel (Cartesian cells, degree <= 5)");
    CUDA_CHECK(cudaSetDevice(op.device));
    const size_t s = op.number == MFHN_F64 ? 8 : 4;
    std::vector<void *> gs((size_t)std::max<long long>(op.n_ghost, 1)), gd(gs.size());
    for (long long g = 0; g < op.n_ghost; ++g)
      {
        if (!ghost_owner || !ghost_remote_index) throw InvalidArgument("null argument");
        const int o = ghost_owner[g];
        if (o < 0 || o >= d.world || o == d.rank || !peer_src[o] || !peer_dst[o]) throw InvalidArgument("bad ghost owner");
        gs[g] = static_cast<char *>(peer_src[o]) + (size_t)ghost_remote_index[g] * s;
        gd[g] = static_cast<char *>(peer_dst[o]) + (size_t)ghost_remote_index[g] * s;
      }
    cudaFree(d.d_ghost_src);
    cudaFree(d.d_ghost_dst);
    d.d_ghost_src = to_device(gs);
    d.d_ghost_dst = to_device(gd);
    if (!d.d_barrier)
      {
        CUDA_CHECK(cudaMalloc(&d.d_barrier, sizeof(int)));
        CUDA_CHECK(cudaMemset(d.d_barrier, 0, sizeof(int)));
      }
    d.peer_src_local = src_local;
    d.peer_dst_local = dst_local;
  });
}
int mfhn_dist_vmult_peer(mfhn_dist h, void *stream, int zero_dst)
{
  return guard([&] {
    if (!h) throw InvalidArgument("null argument");
    Dist &d = *reinterpret_cast<Dist *>(h);
    if (!d.peer_src_local) throw InvalidArgument("mfhn_dist_enable_peer has not been called");
    Operator &op = *d.op;
    CUDA_CHECK(cudaSetDevice(op.device));
    cudaStream_t main = static_cast<cudaStream_t>(stream);
    if (zero_dst) CUDA_CHECK(cudaMemsetAsync(d.peer_dst_local, 0, (size_t)(op.n_owned + op.n_ghost) * (op.number == MFHN_F64 ? 8 : 4), main));
    if (op.number == MFHN_F64)
      dist_vmult_peer_n<double>(d, main);
    else
      dist_vmult_peer_n<float>(d, main);
  });
}
}
