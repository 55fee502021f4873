// trng::uniform01_dist<T> restated (see yarn2.hpp): one engine call scaled into [0, 1): (x - min) * (1/(max - min + 1)).
#pragma once
#include "yarn2.hpp"
namespace trng
{
template <typename T = double>
struct uniform01_dist
{
  template <typename R> NQS_SHIM_HD T operator()(R & r) const
  { return static_cast<T>(r()-R::min())*(static_cast<T>(1)/(static_cast<T>(R::max())-static_cast<T>(R::min())+static_cast<T>(1))); }
};
} // namespace trng
