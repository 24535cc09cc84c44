// TEST INFRASTRUCTURE (oracle).  Stand-in for the reference's src/utils/compile_options.hh in serial (non-Kokkos) builds
// of its own unit tests (oracle/build_ref.py --tests): the same three definitions as the reference's non-Kokkos branch,
// with namespace ELM declared first - the reference's `namespace NS = ELM;` fails to compile when this header is the
// first ELM header of a translation unit (test_SurfAlb.cc, test_CanFlux.cc), SURVEY.md section 8(c).
#pragma once
#define ACCELERATE
#define ELM_LAMBDA [=]
namespace ELM {}
namespace NS = ELM;
