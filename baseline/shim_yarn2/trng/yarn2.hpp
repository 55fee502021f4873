// Restatement of trng::yarn2 (TRNG4 v4.22, pinned by the reference's cmake/FindTRNG4.cmake:46-48; third-party, absent from
// /root/reference and not installable offline) in the library's own class shape, for the SAME-BOX REFERENCE CUDA BUILD
// (baseline/Makefile, targets *-ref-yarn2).  TEST / BENCHMARK INFRASTRUCTURE ONLY: nothing under
// neural_network_quantum_state_b200/ includes this; the product's generator is csrc/yarn2.cuh, written separately (tables +
// binary matrix powers) -- this one follows the documented structure of the library (step / jump2 / jump, exponentiation by
// repeated squaring) so that the two can be compared through the reference's own drivers.
//
// Published algorithm (H. Bauke, S. Mertens, Phys. Rev. E 75, 066701 (2007); TRNG documentation, "yarn2"):
//   r_i = (a1 r_{i-1} + a2 r_{i-2}) mod (2^31-1),  a = (1498809829, 1160990996) [L'Ecuyer],  output g^{r_i} mod (2^31-1)
//   with g = 123567893 and output 0 for r_i = 0;  status after construction (0, 1);  seed(s): (s mod m, 1).
// PARITY UNPINNED against the real library (no golden vector of the stream exists in the reference).
// The reference uses (gpu/include/trng4cuda.cuh:14-65):  host  rng[k].seed(seedNumber); rng[k].jump(2ul*seedDistance*k);
//                                                         device  trng::uniform01_dist<FloatType>()(rng[idx])
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#define NQS_SHIM_HD __host__ __device__
#else
#define NQS_SHIM_HD
#endif
namespace trng
{
class yarn2
{
public:
  typedef int32_t result_type;
  static constexpr result_type modulus = 2147483647;
  static constexpr result_type gen = 123567893;
  NQS_SHIM_HD static constexpr result_type min() { return 0; }
  NQS_SHIM_HD static constexpr result_type max() { return modulus-1; }

  NQS_SHIM_HD yarn2() { a_[0] = 1498809829; a_[1] = 1160990996; r_[0] = 0; r_[1] = 1; }
  void seed(unsigned long s)
  {
    int64_t t = static_cast<int64_t>(s);
    t %= modulus;
    if (t < 0) t += modulus;
    r_[0] = static_cast<result_type>(t);
    r_[1] = 1;
  }
  NQS_SHIM_HD result_type operator()()
  {
    step();
    return r_[0] == 0 ? 0 : power(r_[0]);
  }
  // 2^s steps: the companion matrix squared s times
  void jump2(unsigned int s)
  {
    result_type b[4] = {a_[0], a_[1], 1, 0}, c[4];
    for (unsigned int i = 0; i < s; ++i)
    {
      matmul(b, b, c);
      for (int q = 0; q < 4; ++q) b[q] = c[q];
    }
    const result_type d0 = mod(prod(b[0], r_[0])+prod(b[1], r_[1])), d1 = mod(prod(b[2], r_[0])+prod(b[3], r_[1]));
    r_[0] = d0; r_[1] = d1;
  }
  void jump(unsigned long long s)
  {
    if (s < 16)
    {
      for (unsigned int i = 0; i < s; ++i) step();
      return;
    }
    unsigned int i = 0;
    while (s > 0)
    {
      if (s%2 == 1) jump2(i);
      ++i;
      s >>= 1;
    }
  }
private:
  NQS_SHIM_HD static uint64_t prod(result_type x, result_type y) { return static_cast<uint64_t>(x)*static_cast<uint64_t>(y); }
  NQS_SHIM_HD static result_type mod(uint64_t t) { return static_cast<result_type>(t%static_cast<uint64_t>(modulus)); }
  NQS_SHIM_HD void step()
  {
    const result_type n = mod(prod(a_[0], r_[0])+prod(a_[1], r_[1]));
    r_[1] = r_[0]; r_[0] = n;
  }
  NQS_SHIM_HD static result_type power(result_type n)
  { // gen^n mod modulus by repeated squaring
    uint64_t p = 1, b = gen;
    while (n > 0)
    {
      if (n&1) p = (p*b)%static_cast<uint64_t>(modulus);
      b = (b*b)%static_cast<uint64_t>(modulus);
      n >>= 1;
    }
    return static_cast<result_type>(p);
  }
  static void matmul(const result_type * x, const result_type * y, result_type * z)
  {
    z[0] = mod(prod(x[0], y[0])+prod(x[1], y[2])); z[1] = mod(prod(x[0], y[1])+prod(x[1], y[3]));
    z[2] = mod(prod(x[2], y[0])+prod(x[3], y[2])); z[3] = mod(prod(x[2], y[1])+prod(x[3], y[3]));
  }
  result_type a_[2], r_[2];
};
} // namespace trng
