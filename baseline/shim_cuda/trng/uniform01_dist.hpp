#pragma once
#include "philox_engine.hpp"
namespace trng
{
template <typename T = double>
struct uniform01_dist
{
  template <typename R> NQS_SHIM_HD T operator()(R & r) const { return static_cast<T>(r.next01()); }
};
} // namespace trng
