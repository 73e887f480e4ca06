// Test-infrastructure shim (NOT product code): open_karto's only Boost use is
// boost::shared_mutex / shared_lock / unique_lock (Karto.h:37, 5195-5343).
// Mapping them onto the C++17 standard library lets the reference's own
// Karto.cpp / Mapper.cpp compile unmodified in an image without Boost.
#pragma once
#include <mutex>
#include <shared_mutex>
namespace boost {
using shared_mutex = std::shared_mutex;
template <class M> using shared_lock = std::shared_lock<M>;
template <class M> using unique_lock = std::unique_lock<M>;
}  // namespace boost
