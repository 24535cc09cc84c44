// TEST INFRASTRUCTURE (oracle).  Linked into the serial builds of the reference's test_SurfAlb and test_CanFlux
// (oracle/build_ref.py --tests) and nowhere else: a replaceable global operator new that hands out zero-filled memory.
// Both tests read scratch rows of ELM::Array (new double[n], src/utils/array.hh) that nothing has written - the
// SNICAR and two-stream scratch of test_SurfAlb.cc:150-290, the 2-D work arrays of test_CanFlux.cc:180-320; with
// whatever the allocator recycles they abort before their first comparison, with zeros (what a fresh process usually
// gets from the kernel, and what the Kokkos Views of the driver guarantee) they run to the end and print the counts
// SURVEY.md section 4 lists: 2350 of 2350 and 8560 of 8633 comparisons at 1e-15.
#include <cstdlib>
#include <new>
void* operator new(std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { void* p = std::calloc(1, n ? n : 1); if (!p) throw std::bad_alloc(); return p; }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
