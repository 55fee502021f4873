#pragma once
#include "philox_engine.hpp"
namespace trng
{
struct uniform_int_dist
{
  int a_, b_;
  NQS_SHIM_HD uniform_int_dist(int a, int b): a_(a), b_(b) {}
  template <typename R> NQS_SHIM_HD int operator()(R & r) const { return a_+static_cast<int>(r.next01()*(b_-a_)); }
};
} // namespace trng
