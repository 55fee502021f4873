#pragma once
#include "feed.hpp"
namespace trng
{
// only referenced by out-of-scope samplers (parallel tempering / Kawasaki); present so the headers parse.
struct uniform_int_dist
{
  int a, b;
  uniform_int_dist(int a_ = 0, int b_ = 1): a(a_), b(b_) {}
  template <typename Engine> int operator()(Engine & e) const { return a+static_cast<int>(e.draw()*(b-a)); }
};
} // namespace trng
