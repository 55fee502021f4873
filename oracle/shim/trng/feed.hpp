// Test-infrastructure stand-in for TRNG4 (v4.22 is pinned by the reference's cmake/FindTRNG4.cmake:46-48 but is
// third-party, not under /root/reference and not installable offline).  The reference samplers only use
//   engine.seed(s); engine.jump(2*seedDistance*k);   (cpu/include/impl_mcmc_sampler.hpp:19-23)
//   trng::uniform01_dist<T>()(engine)                (cpu/include/impl_mcmc_sampler.hpp:52)
// so this shim replaces the generator by a FEED of pre-drawn uniforms u[step][chain]: the n-th draw of the engine
// that was jumped to chain k returns u[n][k].  That is exactly the "same pre-drawn uniforms" protocol the
// north star prescribes for accept/reject parity.  The harness constructs samplers with seedDistance = 1, so
// jump(2*k) identifies chain k.  PARITY UNPINNED at the RNG boundary (no yarn2 stream is reproduced).
#pragma once
#include <cstddef>
#include <stdexcept>
namespace trng
{
struct uniform_feed
{
  const double * u = nullptr;   // [steps][nChains]
  long steps = 0;
  int nChains = 0;
};
inline uniform_feed *& current_feed() { static uniform_feed * p = nullptr; return p; }

struct feed_engine
{
  uniform_feed * feed = nullptr;
  long chain = 0, ndraw = 0;
  void seed(unsigned long) { feed = current_feed(); ndraw = 0; chain = 0; }
  void jump(unsigned long long s) { chain = static_cast<long>(s/2ull); }
  double draw()
  {
    if (feed == nullptr || feed->u == nullptr || ndraw >= feed->steps)
      throw std::runtime_error("trng shim: uniform feed exhausted or not set");
    return feed->u[(ndraw++)*static_cast<long>(feed->nChains)+chain];
  }
};
} // namespace trng
