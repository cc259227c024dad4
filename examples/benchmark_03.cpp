// C++ driver over the C ABI, laid out like the reference's benchmark_03
// (/root/reference/benchmark_03.h:382-546, `./benchmark_03 cuda annulus 4`):
// for n_refinements = first..last: mesh -> count cells with hanging nodes -> FE_Q(degree)
// -> MatrixFree -> LaplaceOperator -> src = interpolate(sum sin x_d) -> 100 x { barrier; vmult; synchronize }
// -> min / max / avg time (max over the ranks, benchmark_03.h:501-505).
//
//   ./benchmark_03 <geometry> <degree> [first_refinement last_refinement [n_ranks]]
//
// n_ranks > 1 is the reference's `mpirun -np n_ranks`: the image has no MPI, so the ranks are forked processes, one
// per GPU; the NCCL id, the ghost requests and the timings travel through a shared-memory board.  Every rank builds
// the (replicated) coarse description of the mesh and its own MatrixFree, like p4est ranks do.
#include "../include/mfhn.hpp"

#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>

namespace
{
constexpr int max_ranks = 8;
// shared between the forked ranks (MPI_Barrier / MPI_Bcast / Utilities::MPI::max stand-ins)
struct Board
{
  std::atomic<int> arrived, generation;
  char nccl_id[128];
  double times[max_ranks];
  double norms[max_ranks];
  long long counts[max_ranks][4];
  char tmpdir[64];
};
struct Comm
{
  Board *board;
  int rank, size;
  void barrier() const
  {
    if (size == 1) return;
    const int gen = board->generation.load();
    if (board->arrived.fetch_add(1) + 1 == size)
      {
        board->arrived.store(0);
        board->generation.fetch_add(1);
      }
    else
      while (board->generation.load() == gen) usleep(20);
  }
  double max(double v) const
  {
    if (size == 1) return v;
    board->times[rank] = v;
    barrier();
    double m = 0;
    for (int r = 0; r < size; ++r) m = std::max(m, board->times[r]);
    barrier();
    return m;
  }
  std::string file(int from, int to) const { return std::string(board->tmpdir) + "/req_" + std::to_string(from) + "_" + std::to_string(to); }
};

// Partitioner setup: every rank tells the owners which of their entries it ghosts
void exchange_ghost_requests(mfhn::MatrixFree &mf, const Comm &comm)
{
  for (int p = 0; p < mf.sizes().n_ghost_peers; ++p)
    {
      const int64_t *idx;
      int64_t n;
      const int owner = mf.ghost_peer(p, idx, n);
      std::ofstream f(comm.file(comm.rank, owner), std::ios::binary);
      f.write(reinterpret_cast<const char *>(idx), sizeof(int64_t) * n);
    }
  comm.barrier();
  for (int r = 0; r < comm.size; ++r)
    {
      if (r == comm.rank) continue;
      std::ifstream f(comm.file(r, comm.rank), std::ios::binary | std::ios::ate);
      if (!f) continue;
      std::vector<int64_t> idx((size_t)f.tellg() / sizeof(int64_t));
      f.seekg(0);
      f.read(reinterpret_cast<char *>(idx.data()), sizeof(int64_t) * idx.size());
      mf.set_imports(r, idx.data(), (int64_t)idx.size());
    }
  comm.barrier();
  for (int r = 0; r < comm.size; ++r) std::remove(comm.file(comm.rank, r).c_str());
}
} // namespace

template <int degree>
void run(const std::string &geometry_type, int first, int last, const Comm &comm)
{
  using Number     = double; // benchmark_03.h:390
  using VectorType = mfhn::Vector<Number>;
  const unsigned n_repetitions = 100; // benchmark_03.h:393
  if (comm.rank == 0) std::printf("n_ranks n_levels degree geometry n_cells n_cells_hn n_dofs time_min time_max time_avg GDoF/s\n");
  for (int n_refinements = first; n_refinements <= last; ++n_refinements)
    {
      mfhn::Triangulation tria(geometry_type, n_refinements); // p4est flavour, benchmark_03.h:397-404
      const long long n_cells_w_hn = tria.n_cells_with_hanging_nodes();
      mfhn::DoFHandler dof_handler(tria, degree, comm.size);
      mfhn::MatrixFree matrix_free(dof_handler, comm.rank);
      if (comm.size > 1) exchange_ghost_requests(matrix_free, comm);
      mfhn::LaplaceOperator<3, degree, Number> laplace_operator(matrix_free, /*apply_constraints*/ true);
      if (comm.size > 1)
        {
          if (comm.rank == 0) mfhn::check(mfhn_dist_unique_id(comm.board->nccl_id));
          comm.barrier();
          laplace_operator.attach_communicator(comm.board->nccl_id);
        }
      VectorType src, dst;
      laplace_operator.initialize_dof_vector(src);
      laplace_operator.initialize_dof_vector(dst);
      {
        // VectorTools::interpolate(dof_handler, AnalyticalFunction, src_host)  (benchmark_03.h:455-468)
        const std::vector<double> xyz = matrix_free.owned_support_points();
        std::vector<Number> host((size_t)src.size(), Number(0));
        for (int64_t i = 0; i < matrix_free.sizes().n_owned; ++i) host[i] = std::sin(xyz[3 * i]) + std::sin(xyz[3 * i + 1]) + std::sin(xyz[3 * i + 2]);
        src.import_from_host(host);
        dst = 0.0;
      }
      for (int i = 0; i < 3; ++i) laplace_operator.vmult(dst, src); // connections, lazy initialisations
      cudaDeviceSynchronize();
      dst = 0.0;
      double min_time = 1e10, max_time = 0, avg_time = 0;
      for (unsigned i = 0; i < n_repetitions; ++i)
        {
          cudaDeviceSynchronize();
          comm.barrier(); // MPI_Barrier, benchmark_03.h:477
          const auto t0 = std::chrono::system_clock::now();
          laplace_operator.vmult(dst, src);
          cudaDeviceSynchronize(); // benchmark_03.h:485
          double dt = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now() - t0).count() / 1e9;
          dt        = comm.max(dt);
          min_time = std::min(min_time, dt);
          max_time = std::max(max_time, dt);
          avg_time += dt / n_repetitions;
        }
      // dst accumulated n_repetitions applications (the reference never zeroes it): checksum over the owned entries
      const std::vector<Number> d = dst.to_host();
      double nrm = 0;
      for (int64_t i = 0; i < matrix_free.sizes().n_owned; ++i) nrm += d[i] * d[i];
      if (comm.size > 1)
        {
          comm.board->norms[comm.rank] = nrm;
          comm.barrier();
          nrm = 0;
          for (int r = 0; r < comm.size; ++r) nrm += comm.board->norms[r];
          comm.barrier();
        }
      if (comm.rank == 0)
        {
          std::printf("%d %d %d %s %lld %lld %lld %.4e %.4e %.4e %.2f\n", comm.size, tria.n_global_levels(), degree, geometry_type.c_str(),
                      (long long)tria.n_global_active_cells(), n_cells_w_hn, (long long)dof_handler.n_dofs(), min_time, max_time, avg_time,
                      dof_handler.n_dofs() / avg_time / 1e9);
          std::printf("  |dst|_2 / n_repetitions = %.12e\n", std::sqrt(nrm) / n_repetitions);
          std::fflush(stdout);
        }
    }
}

int run_rank(const std::string &geometry_type, int fe_degree, int first, int last, const Comm &comm)
{
  try
    {
      if (cudaSetDevice(comm.rank) != cudaSuccess) throw mfhn::ExcMessage("cudaSetDevice failed (one GPU per rank)");
      switch (fe_degree)
        {
          case 1: run<1>(geometry_type, first, last, comm); break;
          case 2: run<2>(geometry_type, first, last, comm); break;
          case 3: run<3>(geometry_type, first, last, comm); break;
          case 4: run<4>(geometry_type, first, last, comm); break;
          case 5: run<5>(geometry_type, first, last, comm); break;
          case 6: run<6>(geometry_type, first, last, comm); break;
          case 7: run<7>(geometry_type, first, last, comm); break;
          case 8: run<8>(geometry_type, first, last, comm); break;
          default: throw mfhn::ExcNotImplemented("degree not compiled");
        }
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "rank %d: %s\n", comm.rank, e.what());
      return 1;
    }
  return 0;
}

int main(int argc, char **argv)
{
  const std::string geometry_type = argc > 1 ? argv[1] : "quadrant";
  const int fe_degree             = argc > 2 ? std::atoi(argv[2]) : 4;
  const int first = argc > 3 ? std::atoi(argv[3]) : 4, last = argc > 4 ? std::atoi(argv[4]) : 7;
  const int n_ranks = argc > 5 ? std::atoi(argv[5]) : 1;
  if (n_ranks < 1 || n_ranks > max_ranks)
    {
      std::fprintf(stderr, "n_ranks must be in 1..%d\n", max_ranks);
      return 1;
    }
  if (n_ranks == 1) return run_rank(geometry_type, fe_degree, first, last, Comm{nullptr, 0, 1});
  Board *board = static_cast<Board *>(mmap(nullptr, sizeof(Board), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0));
  if (board == MAP_FAILED) return 1;
  new (board) Board();
  board->arrived.store(0);
  board->generation.store(0);
  std::strcpy(board->tmpdir, "/tmp/mfhn_b03_XXXXXX");
  if (!mkdtemp(board->tmpdir)) return 1;
  std::vector<pid_t> children;
  for (int r = 0; r < n_ranks; ++r) // fork BEFORE any CUDA call: every rank creates its own context
    {
      const pid_t pid = fork();
      if (pid == 0) _exit(run_rank(geometry_type, fe_degree, first, last, Comm{board, r, n_ranks}));
      children.push_back(pid);
    }
  int rc = 0;
  for (pid_t pid : children)
    {
      int status = 0;
      waitpid(pid, &status, 0);
      if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
    }
  rmdir(board->tmpdir);
  return rc;
}
