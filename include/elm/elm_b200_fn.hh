// elm_b200_fn.hh - marshalling behind the library-level API of include/elm/*.h.
//
// The reference's physics library is a set of free function templates ELM::<namespace>::<function>(const LandType&,
// scalars by reference, per-column rows by value as ArrayD1 ...) that a driver calls per column from inside its own
// parallel loop (e.g. src/physics/canopy_hydrology.h, canopy_hydrology_impl.hh:8-357).  The headers of include/elm/
// declare the same functions - names, namespaces, argument order and meaning - and run each call on the device:
// the arguments are packed into one flat array of doubles in argument order (rows expanded in place, ints and bools
// as doubles, LandType checked on the host), elmk_fn_call executes the function's device code - the same code the
// fused column kernels run (csrc/phys_*.h, namespaces hyd / rad / tmp) - and the outputs are unpacked into the
// caller's variables.  One launch per call: this is the drop-in for code written against the function API (the
// reference's own test/*.cc compile against these headers unchanged), not the fast path - that is elmk_step.
#pragma once
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../elmk_b200.h"

namespace ELM {
namespace b200 {
namespace fn {

class Args {
public:
  explicit Args(int fn) : fn_(fn) {}
  Args& in(double v) { buf_.push_back(v); return *this; }
  Args& in(int v) { buf_.push_back(static_cast<double>(v)); return *this; }
  Args& in(bool v) { buf_.push_back(v ? 1.0 : 0.0); return *this; }
  // a scalar the function writes (or reads and writes): remembered for unpack()
  Args& io(double& v) { outs_.push_back({buf_.size(), &v, nullptr, nullptr}); buf_.push_back(v); return *this; }
  Args& io(int& v) { outs_.push_back({buf_.size(), nullptr, &v, nullptr}); buf_.push_back(static_cast<double>(v)); return *this; }
  Args& io(bool& v) { outs_.push_back({buf_.size(), nullptr, nullptr, &v}); buf_.push_back(v ? 1.0 : 0.0); return *this; }
  // a per-column row (ArrayD1 view): n elements through operator()(i); written back when `writes`
  template <class Row> Args& row(Row a, int n, bool writes) {
    const size_t at = buf_.size();
    for (int i = 0; i < n; ++i) buf_.push_back(static_cast<double>(a(i)));
    if (writes) rows_.push_back([a, at, n](const std::vector<double>& b) mutable {
      for (int i = 0; i < n; ++i) a(i) = static_cast<std::decay_t<decltype(a(0))>>(b[at + i]);
    });
    return *this;
  }
  void call() {
    const int rc = elmk_fn_call(0, fn_, buf_.data(), static_cast<int64_t>(buf_.size()));
    if (rc != ELMK_OK) throw std::runtime_error("ELM (B200 backend): elmk_fn_call(" + std::to_string(fn_) + ") failed with " + std::to_string(rc));
    for (auto& o : outs_) {
      if (o.d) *o.d = buf_[o.at];
      if (o.i) *o.i = static_cast<int>(buf_[o.at]);
      if (o.b) *o.b = buf_[o.at] != 0.0;
    }
    for (auto& r : rows_) r(buf_);
  }

private:
  struct Out { size_t at; double* d; int* i; bool* b; };
  int fn_;
  std::vector<double> buf_;
  std::vector<Out> outs_;
  std::vector<std::function<void(const std::vector<double>&)>> rows_;
};

// the device code resolves the land-unit branches for soil / crop columns without lake or urban points - what the
// reference's driver runs (elm_kokkos_interface.cc:84-90); anything else is refused, as elmk_set_tables does
template <class Land> void require_soil(const Land& L) {
  if (!(L.ltype == 1 || L.ltype == 2) || L.lakpoi || L.urbpoi)
    throw std::runtime_error("ELM (B200 backend): only soil / crop land units without lake or urban points are on the hot path");
}

} // namespace fn
} // namespace b200
} // namespace ELM
