// C++ driver over the C ABI, laid out like the reference's benchmark_03
// (/root/reference/benchmark_03.h:382-546, `./benchmark_03 cuda annulus 4`):
// for n_refinements = first..last: mesh -> count cells with hanging nodes -> FE_Q(degree)
// -> LaplaceOperator -> src = interpolate(sum sin x_d) -> 100 x { vmult; synchronize }
// -> min / max / avg time.  Single process (the image has no MPI); the partitioned
// path is driven from bench.py through torch.distributed.
//
//   ./benchmark_03 <geometry> <degree> [first_refinement last_refinement]
#include "../include/mfhn.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

template <int degree>
void run(const std::string &geometry_type, int first, int last)
{
  using Number     = double; // benchmark_03.h:390
  using VectorType = mfhn::Vector<Number>;
  const unsigned n_repetitions = 100; // benchmark_03.h:393
  std::printf("n_levels degree geometry n_cells n_cells_hn n_dofs time_min time_max time_avg GDoF/s\n");
  for (int n_refinements = first; n_refinements <= last; ++n_refinements)
    {
      mfhn::Triangulation tria(geometry_type, n_refinements); // p4est flavour, benchmark_03.h:397-404
      const long long n_cells_w_hn = tria.n_cells_with_hanging_nodes();
      mfhn::DoFHandler dof_handler(tria, degree);
      mfhn::LaplaceOperator<3, degree, Number> laplace_operator(dof_handler, /*apply_constraints*/ true);
      VectorType src, dst;
      laplace_operator.initialize_dof_vector(src);
      laplace_operator.initialize_dof_vector(dst);
      {
        // VectorTools::interpolate(dof_handler, AnalyticalFunction, src_host)  (benchmark_03.h:455-468)
        std::vector<double> xyz(3 * (size_t)dof_handler.n_dofs());
        mfhn::check(mfhn_dofs_support_points(dof_handler.handle(), 0, dof_handler.n_dofs(), xyz.data()));
        std::vector<Number> host(dof_handler.n_dofs());
        for (size_t i = 0; i < host.size(); ++i) host[i] = std::sin(xyz[3 * i]) + std::sin(xyz[3 * i + 1]) + std::sin(xyz[3 * i + 2]);
        src.import_from_host(host);
        dst = 0.0;
      }
      double min_time = 1e10, max_time = 0, avg_time = 0;
      for (unsigned i = 0; i < n_repetitions; ++i)
        {
          cudaDeviceSynchronize();
          const auto t0 = std::chrono::system_clock::now();
          laplace_operator.vmult(dst, src);
          cudaDeviceSynchronize(); // benchmark_03.h:485
          const double dt = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now() - t0).count() / 1e9;
          min_time = std::min(min_time, dt);
          max_time = std::max(max_time, dt);
          avg_time += dt / n_repetitions;
        }
      std::printf("%d %d %s %lld %lld %lld %.4e %.4e %.4e %.2f\n", tria.n_global_levels(), degree, geometry_type.c_str(),
                  (long long)tria.n_global_active_cells(), n_cells_w_hn, (long long)dof_handler.n_dofs(), min_time, max_time, avg_time,
                  dof_handler.n_dofs() / avg_time / 1e9);
      // dst accumulated n_repetitions applications (the reference never zeroes it): report a checksum
      const std::vector<Number> d = dst.to_host();
      double nrm = 0;
      for (const Number v : d) nrm += v * v;
      std::printf("  |dst|_2 / n_repetitions = %.12e\n", std::sqrt(nrm) / n_repetitions);
    }
}

int main(int argc, char **argv)
{
  const std::string geometry_type = argc > 1 ? argv[1] : "quadrant";
  const int fe_degree             = argc > 2 ? std::atoi(argv[2]) : 4;
  const int first = argc > 3 ? std::atoi(argv[3]) : 4, last = argc > 4 ? std::atoi(argv[4]) : 7;
  try
    {
      switch (fe_degree)
        {
          case 1: run<1>(geometry_type, first, last); break;
          case 2: run<2>(geometry_type, first, last); break;
          case 3: run<3>(geometry_type, first, last); break;
          case 4: run<4>(geometry_type, first, last); break;
          case 5: run<5>(geometry_type, first, last); break;
          case 6: run<6>(geometry_type, first, last); break;
          case 7: run<7>(geometry_type, first, last); break;
          case 8: run<8>(geometry_type, first, last); break;
          default: throw mfhn::ExcNotImplemented("degree not compiled");
        }
    }
  catch (const std::exception &e)
    {
      std::fprintf(stderr, "%s\n", e.what());
      return 1;
    }
  return 0;
}
