// Device-side fp64 complex helpers shared by every kernel of libnqs_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nqs
{
typedef double2 cd; // interleaved {re, im} == thrust::complex<double> == nqs_cdouble

#define NQS_LN2 0.6931471805599453 // std::log(2.0); ref: kln2d, gpu/include/impl_neural_quantum_state.cuh:54-57

__host__ __device__ __forceinline__ cd cmake(double re, double im) { cd r; r.x = re; r.y = im; return r; }
__host__ __device__ __forceinline__ cd cadd(cd a, cd b) { return cmake(a.x+b.x, a.y+b.y); }
__host__ __device__ __forceinline__ cd csub(cd a, cd b) { return cmake(a.x-b.x, a.y-b.y); }
__host__ __device__ __forceinline__ cd cmul(cd a, cd b) { return cmake(a.x*b.x-a.y*b.y, a.x*b.y+a.y*b.x); }
__host__ __device__ __forceinline__ cd cscale(cd a, double s) { return cmake(a.x*s, a.y*s); }
__host__ __device__ __forceinline__ cd cconj(cd a) { return cmake(a.x, -a.y); }
__host__ __device__ __forceinline__ double cnorm(cd a) { return a.x*a.x+a.y*a.y; }

// log(cosh(z)), the reference's overflow-safe form (ref: gpu_device::logcosh, impl_neural_quantum_state.cuh:1238-1245):
//   e = exp(-2|x|);  log((1+e) cos y + i (1-e) sin y sgn x) + |x| - ln 2
__device__ __forceinline__ cd c_logcosh(cd z)
{
  const double ax = fabs(z.x);
  double s, c;
  sincos(z.y, &s, &c);
  const double e = exp(-2.0*ax);
  const double re = (1.0+e)*c, im = (1.0-e)*s*copysign(1.0, z.x);
  return cmake(0.5*log(fma(re, re, im*im))+(ax-NQS_LN2), atan2(im, re));
}

// Re log cosh(z) only (what the Metropolis test needs).
__device__ __forceinline__ double re_logcosh(cd z)
{
  const double ax = fabs(z.x);
  double s, c;
  sincos(z.y, &s, &c);
  const double e = exp(-2.0*ax);
  const double re = (1.0+e)*c, im = (1.0-e)*s;
  return 0.5*log(fma(re, re, im*im))+(ax-NQS_LN2);
}

// tanh(x+iy) = ((1-e^2) sgn x + 2i e sin 2y) / ((1-e)^2 + 4 e cos^2 y),  e = exp(-2|x|): no overflow, and the
// denominator is a sum of squares (no cancellation near the poles of tanh).  ref uses thrust::tanh (:1441,1446,1634).
__device__ __forceinline__ cd c_tanh(cd z)
{
  const double ax = fabs(z.x);
  double s, c;
  sincos(z.y, &s, &c);
  const double e = exp(-2.0*ax);
  const double ome = 1.0-e;
  const double den = fma(ome, ome, 4.0*e*c*c);
  const double inv = 1.0/den;
  return cmake(copysign((1.0-e*e)*inv, z.x), 4.0*e*s*c*inv);
}

__device__ __forceinline__ cd c_exp(cd z)
{
  double s, c;
  sincos(z.y, &s, &c);
  const double m = exp(z.x);
  return cmake(m*c, m*s);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ cd warp_sum(cd v) { return cmake(warp_sum(v.x), warp_sum(v.y)); }

// ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX) shared by the pipelined kernels (sv_fused.cuh, fast_kernels.cuh) ----
__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t * bar, const uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, const uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, const uint32_t parity)
{ // try_wait suspends the thread in hardware (up to the time hint) and wakes it when the phase completes: no busy polling
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// 1-D TMA bulk copy global -> this CTA's shared memory, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void * dst, const void * src, const uint32_t bytes, uint64_t * bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
    :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- Philox4x32-10 counter RNG (Salmon et al. SC'11); restated in oracle/nqs_oracle.py:philox4x32_10 ----------------
__host__ __device__ __forceinline__ void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1)
{
  const uint64_t p0 = (uint64_t)0xD2511F53u*c[0], p1 = (uint64_t)0xCD9E8D57u*c[2];
  const uint32_t hi0 = (uint32_t)(p0>>32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1>>32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1^c[1]^k0, n2 = hi0^c[3]^k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// U[0,1) with 53 bits for (seed, global chain id, proposal index)
__host__ __device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t chain, uint64_t step)
{
  uint32_t c[4] = {(uint32_t)chain, (uint32_t)(chain>>32), (uint32_t)step, (uint32_t)(step>>32)};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed>>32);
#pragma unroll
  for (int r = 0; r < 10; ++r)
  {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return ((double)(c[0]>>5)*67108864.0+(double)(c[1]>>6))*(1.0/9007199254740992.0);
}
} // namespace nqs
