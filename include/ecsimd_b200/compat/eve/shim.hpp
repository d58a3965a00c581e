// The few EVE names the reference's P-256 tests use (eve::all / any on comparison results, eve::zero(eve::as<mask>())),
// over ecsimd::wide_mask -- EVE itself (a CPU SIMD abstraction) has no role on the GPU path.
#pragma once
#include "../../ecsimd.hpp"
namespace eve {
template <class T>
struct as { using type = T; };
template <class T>
inline T zero(as<T>) { return T{}; }
inline bool all(ecsimd::wide_mask const& m) { return ecsimd::all(m); }
inline bool any(ecsimd::wide_mask const& m) { return ecsimd::any(m); }
inline bool all(bool v) { return v; }
inline bool any(bool v) { return v; }
}  // namespace eve
