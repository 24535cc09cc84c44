// TEST INFRASTRUCTURE (oracle).  Stand-in for the reference's src/utils/invoke_kernel.hh in serial builds of its unit
// tests: apply_parallel_for(kernel, [name,] N) as a plain loop.  The reference's own serial branch lacks the named
// overload and apply_parallel_for_tuple_impl (invoke_kernel.hh:41-47,64-74), which test_SurfAlb / test_CanFlux pull in
// through the aerosol and soil-temperature headers.
#pragma once
#include <cstddef>
#include <functional>
#include <string>
#include <tuple>
#include <utility>
#include "compile_options.hh"
namespace ELM {
template <typename F> void apply_parallel_for(F&& kernel, int N) { for (int i = 0; i < N; ++i) std::invoke(kernel, i); }
template <typename F> void apply_parallel_for(F&& kernel, const std::string&, int N) { for (int i = 0; i < N; ++i) std::invoke(kernel, i); }
template <typename F, typename T> void apply_parallel_for_tuple(F&& kernel, T&& args) {
  const int N = static_cast<int>(std::get<std::tuple_size_v<std::remove_reference_t<T>> - 1>(args));
  for (int i = 0; i < N; ++i) std::invoke(kernel, i);
}
} // namespace ELM
