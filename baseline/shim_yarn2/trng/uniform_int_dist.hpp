// only referenced by samplers outside the path (parallel tempering / Kawasaki); present so that the reference headers parse
#pragma once
#include "uniform01_dist.hpp"
namespace trng
{
struct uniform_int_dist
{
  int a_, b_;
  NQS_SHIM_HD uniform_int_dist(int a, int b): a_(a), b_(b) {}
  template <typename R> NQS_SHIM_HD int operator()(R & r) const { return a_+static_cast<int>(uniform01_dist<double>()(r)*(b_-a_)); }
};
} // namespace trng
