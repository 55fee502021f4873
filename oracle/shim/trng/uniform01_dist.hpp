#pragma once
#include "feed.hpp"
namespace trng
{
template <typename T = double>
struct uniform01_dist
{
  template <typename Engine> T operator()(Engine & e) const { return static_cast<T>(e.draw()); }
};
} // namespace trng
