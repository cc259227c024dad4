// Error convention of the C ABI: int status + mfhn_last_error(); nothing throws
// across the boundary.  The C++ wrapper (include/mfhn.hpp) turns non-zero into
// exceptions, mirroring the reference's AssertThrow(..., ExcMessage /
// ExcNotImplemented) use (benchmark_01.h:204-217, benchmark_03.h:332,404,616).
#pragma once
#include <stdexcept>
#include <string>

namespace mfhn
{
struct InvalidArgument : std::runtime_error
{
  using std::runtime_error::runtime_error;
};
struct CudaError : std::runtime_error
{
  using std::runtime_error::runtime_error;
};
struct NotImplemented : std::runtime_error
{
  using std::runtime_error::runtime_error;
};

void set_last_error(const std::string &msg);

template <typename F>
int guard(F &&f)
{
  try
    {
      f();
      return 0;
    }
  catch (const InvalidArgument &e)
    {
      set_last_error(e.what());
      return 1;
    }
  catch (const std::invalid_argument &e)
    {
      set_last_error(e.what());
      return 1;
    }
  catch (const CudaError &e)
    {
      set_last_error(e.what());
      return 2;
    }
  catch (const NotImplemented &e)
    {
      set_last_error(e.what());
      return 3;
    }
  catch (const std::exception &e)
    {
      set_last_error(e.what());
      return 1;
    }
  catch (...)
    {
      set_last_error("unknown error");
      return 1;
    }
}
} // namespace mfhn
