// trng::yarn2 + trng::uniform01_dist<double> on the device: the parallel random number stream of the reference sampler
// (ref gpu/include/trng4cuda.cuh:14-65 -- one engine per Markov chain, `seed(seedNumber); jump(2ul*seedDistance*k)`, one
// uniform01 draw per chain and proposal; call sites gpu/include/mcmc_sampler.cuh:34, impl_mcmc_sampler.cuh:34).
//
// TRNG4 (v4.22, pinned by the reference's cmake/FindTRNG4.cmake:46-48) is a third-party library whose source is neither under
// /root/reference nor installable offline, so this file restates its PUBLISHED algorithm (H. Bauke, S. Mertens, "Random numbers
// for large-scale distributed Monte Carlo simulations", Phys. Rev. E 75, 066701 (2007), and the TRNG documentation of yarn2):
//   state  (r0, r1) in F_p^2, p = 2^31-1;  default-constructed (0, 1);  seed(s): r0 = int64(s) mod p, r1 = 1
//   step   r0' = (a0 r0 + a1 r1) mod p, r1' = r0      with L'Ecuyer's multipliers a0 = 1498809829, a1 = 1160990996 ("LEcuyer1")
//   output x = 0 if r0 == 0 else g^r0 mod p, g = 123567893 (TRNG tabulates g^i for i < 2^16 and g^(i 2^16) for i < 2^15)
//   jump(s) advances s steps: the 2x2 companion matrix raised to the power s (TRNG composes jump2(i) = matrix^(2^i) for the
//          set bits of s, steps one by one for s < 16; all three are the same map since the arithmetic is exact)
//   uniform01_dist<double>: x * (1/(max-min+1)) = x * (1/2147483647.0) in [0,1), ONE engine call per uniform
// PARITY UNPINNED: no golden vector of the stream exists in the reference and the library cannot be run here; the restatement is
// checked against two independent restatements (oracle/yarn2.py, baseline/shim_yarn2) and the algebraic identities of the
// generator (tests/test_yarn2_cpu.py).  The last-bit form of the scaling (multiply by the reciprocal) is from the documentation.
//
// Layout here: the stream is COUNTER-ADDRESSED.  The state of chain k after n draws is matrix^(2 d k + n) applied to the seed
// state, so the handle keeps only (seedNumber, seedDistance, draws so far): yarn2_fill_kernel either continues from the
// per-chain state cached by the previous launch or rebuilds it by two matrix powers (first launch, restart from a checkpoint,
// different number of GPUs), and writes the uniforms of the next `nsteps` proposals as u[step][chain] -- the feed layout every
// sweep kernel already reads (nqs_set_uniforms).  One thread per chain, stores coalesced over chains; 16.8 MB per sweep at
// N = 128, K = 16384, i.e. ~1 % of the bytes of one S*v.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nqs
{
constexpr uint32_t YARN2_P = 2147483647u;      // modulus 2^31-1
constexpr uint32_t YARN2_GEN = 123567893u;     // generator of the multiplicative group used by the output map
constexpr uint32_t YARN2_A0 = 1498809829u, YARN2_A1 = 1160990996u;
constexpr int YARN2_TAB0 = 0x10000, YARN2_TAB1 = 0x08000;

// t mod (2^31-1) for t < 2^63 (two folds of the Mersenne modulus, canonical residue in [0, p))
__host__ __device__ __forceinline__ uint32_t yarn2_mod(uint64_t t)
{
  t = (t&YARN2_P)+(t>>31);
  t = (t&YARN2_P)+(t>>31);
  return (uint32_t)(t >= YARN2_P ? t-YARN2_P : t);
}
__host__ __device__ __forceinline__ uint32_t yarn2_mulmod(uint32_t a, uint32_t b) { return yarn2_mod((uint64_t)a*b); }

// seed(s): the library takes the unsigned long over into a signed 64-bit integer before reducing it into [0, p)
__host__ __device__ __forceinline__ uint32_t yarn2_seed_state(unsigned long long s)
{
  long long t = (long long)s;
  t %= (long long)YARN2_P;
  if (t < 0) t += (long long)YARN2_P;
  return (uint32_t)t;
}

struct Yarn2Mat { uint32_t m00, m01, m10, m11; };
__host__ __device__ __forceinline__ Yarn2Mat yarn2_matmul(const Yarn2Mat & a, const Yarn2Mat & b)
{
  Yarn2Mat c;
  c.m00 = yarn2_mod((uint64_t)a.m00*b.m00+(uint64_t)a.m01*b.m10);
  c.m01 = yarn2_mod((uint64_t)a.m00*b.m01+(uint64_t)a.m01*b.m11);
  c.m10 = yarn2_mod((uint64_t)a.m10*b.m00+(uint64_t)a.m11*b.m10);
  c.m11 = yarn2_mod((uint64_t)a.m10*b.m01+(uint64_t)a.m11*b.m11);
  return c;
}
// (r0, r1) <- companion^s (r0, r1)
__host__ __device__ inline void yarn2_jump(uint32_t & r0, uint32_t & r1, unsigned long long s)
{
  Yarn2Mat acc = {1u, 0u, 0u, 1u}, b = {YARN2_A0, YARN2_A1, 1u, 0u};
  while (s != 0ull)
  {
    if (s&1ull) acc = yarn2_matmul(b, acc);
    s >>= 1;
    if (s != 0ull) b = yarn2_matmul(b, b);
  }
  const uint32_t n0 = yarn2_mod((uint64_t)acc.m00*r0+(uint64_t)acc.m01*r1);
  const uint32_t n1 = yarn2_mod((uint64_t)acc.m10*r0+(uint64_t)acc.m11*r1);
  r0 = n0; r1 = n1;
}
__host__ __device__ inline uint32_t yarn2_powmod(uint32_t base, uint32_t e)
{
  uint32_t acc = 1u;
  while (e != 0u)
  {
    if (e&1u) acc = yarn2_mulmod(acc, base);
    e >>= 1;
    if (e != 0u) base = yarn2_mulmod(base, base);
  }
  return acc;
}

// tab[i] = g^i (i < 2^16), tab[2^16 + i] = g^(i 2^16) (i < 2^15)
__global__ void yarn2_table_kernel(uint32_t * __restrict__ tab)
{
  const int i = blockIdx.x*blockDim.x+threadIdx.x;
  if (i < YARN2_TAB0) tab[i] = yarn2_powmod(YARN2_GEN, (uint32_t)i);
  else if (i < YARN2_TAB0+YARN2_TAB1) tab[i] = yarn2_powmod(YARN2_GEN, (uint32_t)(i-YARN2_TAB0)<<16);
}

struct Yarn2FillArgs
{
  long long K;                     // chains of this rank
  long long chain_offset;          // global id of local chain 0 (the jump distance is a function of the GLOBAL chain)
  unsigned long long seed, seed_distance, draws_done;
  long long nsteps;
  int rebuild;                     // 1: state from (seed, jump, draws_done); 0: continue from `state`
  const uint32_t * tab;
  uint2 * state;                   // [K] (r0, r1) after the last draw
  double * u;                      // [nsteps][K]
};

__global__ void __launch_bounds__(128) yarn2_fill_kernel(const Yarn2FillArgs a)
{
  const long long k = (long long)blockIdx.x*blockDim.x+threadIdx.x;
  if (k >= a.K) return;
  uint32_t r0, r1;
  if (a.rebuild)
  {
    r0 = yarn2_seed_state(a.seed); r1 = 1u;                                          // seed(seedNumber)
    yarn2_jump(r0, r1, 2ull*a.seed_distance*(unsigned long long)(a.chain_offset+k)); // jump(2ul*seedDistance*k), 64-bit wrap as in the ref
    yarn2_jump(r0, r1, a.draws_done);
  }
  else
  {
    const uint2 s = a.state[k];
    r0 = s.x; r1 = s.y;
  }
  const uint32_t * __restrict__ t0 = a.tab;
  const uint32_t * __restrict__ t1 = a.tab+YARN2_TAB0;
  double * __restrict__ out = a.u+k;
  for (long long n = 0; n < a.nsteps; ++n)
  {
    const uint32_t nr = yarn2_mod((uint64_t)YARN2_A0*r0+(uint64_t)YARN2_A1*r1);
    r1 = r0; r0 = nr;
    const uint32_t x = (r0 == 0u) ? 0u : yarn2_mulmod(__ldg(t1+(r0>>16)), __ldg(t0+(r0&0xffffu)));
    out[n*a.K] = (double)x*(1.0/2147483647.0);
  }
  a.state[k] = make_uint2(r0, r1);
}
} // namespace nqs
