// Metropolis sweep of the complex RBM with an fp32 FILTER in front of the exact fp64 accept test.
// EXPERIMENT, OPT-IN (NQS_SWEEP_F32=1): exact, but measured slower than fast_kernels.cuh (see engine.cu: launch_sweep).
//
// ref: BaseParallelSampler::do_mcmc_steps (gpu/include/impl_mcmc_sampler.cuh:28-39) with RBM::forward(int) / spin_flip
// (impl_neural_quantum_state.cuh:93-104,172-182) -- same decisions as fast_kernels.cuh: rbm_sweep_fast_kernel, i.e. as the reference
// given the same uniforms (exact accept/reject parity).
//
// Why: ncu of the all-fp64 kernel (profiles/r2_sweep_full_summary.md) shows 784 instructions per proposal and warp at 8 warps per
// SM (255 registers for the fp64 state of two chains), issue slots 38 % busy, fp64 pipe 30 %: the sweep is bound by dependent
// instruction latency at low occupancy.  A Metropolis decision is a COMPARISON, u R0 < P' A with P' = prod_j f_j the flip
// product; it needs P' only to the accuracy that separates it from u R0.  So:
//   * every proposal forms P' in fp32 from an fp32 copy of the state (registers: 4 floats per hidden unit, a third of the fp64
//     state's registers) and fp32 tables, TOGETHER WITH a rigorous bound delta on its relative error: each factor
//     f = T1 + T2 + sigma (T3 - T4) is a sum of four products, so |f~ - f| <= 8 u (|T1|+|T2|+|T3|+|T4|), u = 2^-24, and the
//     relative errors of the factors add up along the product (cancellation inside a factor -- f -> 0 near theta = i pi/2 --
//     shows up as a large |T|/f and widens delta for that proposal only);
//   * if u R0 lies outside [P'A (1-delta), P'A (1+delta)] the decision is the exact one by construction; otherwise (~1e-4 of the
//     proposals) the warp recomputes P' in fp64 from the fp64 state and the fp64 tables, exactly as rbm_sweep_fast_kernel does;
//   * the fp64 state (double angles, as in fast_kernels.cuh) lives in SHARED memory and is touched only on accepted flips,
//     which also refresh the fp32 copy and the tracked reference product R0 = prod_j (cosh 2x_j + cos 2y_j) from the new state.
// theta is replayed exactly after each sweep and the state rebuilt from it, as in rbm_sweep_fast_kernel (no drift; theta, spins
// bit-identical to the generic kernel).  One chain per warp, 16+ warps per SM.
#pragma once
#include "fast_kernels.cuh"

namespace nqs
{
#define NQS_S32_STAGES 3
struct F32SweepArgs
{
  FastSweepArgs b;           // everything rbm_sweep_fast_kernel takes (ftab_a/b: fp64 tables of 4W)
  const float4 * ftab32;     // [N][Mpad] (cosh 4ReW, sinh 4ReW, cos 4ImW, sin 4ImW) rounded to fp32
  float delta_scale;         // 1: the rigorous bound; tests widen it (every proposal takes the fp64 path) or set it huge
  unsigned long long * stats;  // optional [2]: proposals, proposals decided by the fp64 path
};

inline size_t f32_sweep_smem_bytes(int N, int warps, int Mpad)
{
  const size_t npad = (size_t)((N+15)/16)*16;
  size_t b = (size_t)warps*Mpad*32;                                   // fp64 state (sinh 2x, cosh 2x | cos 2y, sin 2y)
  b += (size_t)NQS_S32_STAGES*Mpad*(16+32);                           // per stage: fp32 row | fp64 rows a, b
  b += (size_t)warps*(npad+N)+(size_t)N*sizeof(int);
  return (b+15)/16*16+2*NQS_S32_STAGES*sizeof(uint64_t)+16;
}

__device__ __forceinline__ void split_me32(const float p, float & m, int & e)
{ // positive finite float -> mantissa in [1,2), exponent
  const int bits = __float_as_int(p);
  e = ((bits>>23)&0xff)-127;
  m = __int_as_float((bits&0x007fffff)|0x3f800000);
}

template <int JPL>
__global__ void __launch_bounds__(256, 2) rbm_sweep_f32_kernel(const F32SweepArgs fa)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FastSweepArgs & a = fa.b;
  const int warps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  const int N = a.N, M = a.M;
  constexpr int Mpad = 32*JPL;
  const int npad = ((N+15)/16)*16;
  double2 * st_all = reinterpret_cast<double2*>(smem_raw);                         // [warps][2][Mpad]: (S, Ch) then (cy, sy)
  unsigned char * stage0 = smem_raw+(size_t)warps*Mpad*32;                         // [STAGES][ float4[Mpad] | double2[Mpad] | double2[Mpad] ]
  constexpr size_t stage_bytes = (size_t)Mpad*48;
  int * ord = reinterpret_cast<int*>(stage0+(size_t)NQS_S32_STAGES*stage_bytes);   // [N]
  int8_t * spall = reinterpret_cast<int8_t*>(ord+N);                               // [warps][npad]
  int8_t * recall = spall+(size_t)warps*npad;                                      // [warps][N]
  uint64_t * full = reinterpret_cast<uint64_t*>(smem_raw+(((size_t)(reinterpret_cast<unsigned char*>(recall)+(size_t)warps*N-smem_raw)+15)/16)*16);
  uint64_t * empty = full+NQS_S32_STAGES;
  double2 * stA = st_all+(size_t)w*2*Mpad;          // (sinh 2x, cosh 2x)
  double2 * stB = stA+Mpad;                         // (cos 2y, sin 2y)
  int8_t * sp = spall+(size_t)w*npad;
  int8_t * rec = recall+(size_t)w*N;
  for (int i = threadIdx.x; i < N; i += blockDim.x) ord[i] = a.order[i];
  const long long kblock = (long long)blockIdx.x*warps;
  if (threadIdx.x == 0)
  {
    long long nact = a.K-kblock;
    if (nact > warps) nact = warps;
    if (nact < 1) nact = 1;
    for (int q = 0; q < NQS_S32_STAGES; ++q) { mbar_init(full+q, 1); mbar_init(empty+q, (uint32_t)nact); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long k = kblock+w;
  if (k >= a.K) return;
  const cd * avis = a.params+(size_t)N*M;
  for (int i = lane; i < N; i += 32) sp[i] = a.spins[k*N+i];
  cd ln0 = a.lnpsi0[k], sa = a.sa[k];
  bool any_acc = false;
  // R0 = exp(2 (Re lnpsi0 - Re sa)) 2^Mpad = the product of the TRACKED amplitude's factors, as mantissa * 2^exponent
  double r0m; int r0e;
  {
    const double v = 2.0*(ln0.x-sa.x)*1.4426950408889634;
    const double fl = floor(v);
    r0m = exp2(v-fl);
    r0e = (int)fmax(fmin(fl, 100000.0), -100000.0)+Mpad;
  }
  __syncwarp();
  int pos = a.pos0;
  long long t_glob = 0;
  const long long t_end = (long long)a.nsweeps*N;
  auto issue_rows = [&](const long long q)
  {
    const int sq = ord[(int)(((long long)a.pos0+q)%N)];
    unsigned char * dst = stage0+(size_t)(q%NQS_S32_STAGES)*stage_bytes;
    uint64_t * bar = full+(int)(q%NQS_S32_STAGES);
    mbar_expect_tx(bar, (uint32_t)stage_bytes);
    tma_load_1d(dst, fa.ftab32+(size_t)sq*Mpad, (uint32_t)(Mpad*16), bar);
    tma_load_1d(dst+(size_t)Mpad*16, a.ftab_a+(size_t)sq*Mpad, (uint32_t)(Mpad*16), bar);
    tma_load_1d(dst+(size_t)Mpad*32, a.ftab_b+(size_t)sq*Mpad, (uint32_t)(Mpad*16), bar);
  };
  if (w == 0 && lane == 0)
    for (long long q = 0; q < NQS_S32_STAGES-1 && q < t_end; ++q) issue_rows(q);
  double ubuf = 0.0;
  unsigned long long n_slow = 0;
  const float uerr = 5.9604645e-8f*fa.delta_scale;      // 2^-24 times the test scale

  for (int sweep = 0; sweep < a.nsweeps; ++sweep)
  {
    // ---- (1) state from the exact theta: fp64 in shared memory, fp32 copy in registers
    float4 sf[JPL];                                        // (sinh 2x, cosh 2x, cos 2y, sin 2y)
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj)
    {
      const int j = lane+32*jj;
      cd th = cmake(0.0, 0.0);
      if (j < M) th = a.theta[k*M+j];
      const double ex = exp(2.0*th.x), emx = 1.0/ex;
      const double S = 0.5*(ex-emx), Ch = 0.5*(ex+emx);
      double sy, cy;
      sincos(2.0*th.y, &sy, &cy);
      stA[j] = make_double2(S, Ch); stB[j] = make_double2(cy, sy);
      sf[jj] = make_float4((float)S, (float)Ch, (float)cy, (float)sy);
    }
    __syncwarp();
    const int pos_sweep0 = pos;
    // ---- (2) N proposals
    for (int t = 0; t < N; ++t, ++t_glob)
    {
      if ((t_glob&31) == 0)
      {
        const long long tt = t_glob+lane;
        if (tt < t_end)
          ubuf = a.uniforms ? a.uniforms[tt*a.K+k] : philox_uniform(a.seed, (unsigned long long)(a.chain_offset+k), a.step0+(unsigned long long)tt);
      }
      const double u = __shfl_sync(0xffffffffu, ubuf, (int)(t_glob&31));
      const int site = ord[pos];
      pos = (pos+1 == N) ? 0 : pos+1;
      const int slot = (int)(t_glob%NQS_S32_STAGES);
      if (w == 0 && lane == 0)
      {
        const long long q = t_glob+NQS_S32_STAGES-1;
        if (q < t_end)
        {
          if (t_glob > 0) mbar_wait(empty+(int)(q%NQS_S32_STAGES), (uint32_t)(((t_glob-1)/NQS_S32_STAGES)&1));
          issue_rows(q);
        }
      }
      mbar_wait(full+slot, (uint32_t)((t_glob/NQS_S32_STAGES)&1));
      const unsigned char * stg = stage0+(size_t)slot*stage_bytes;
      const float4 * t32 = reinterpret_cast<const float4*>(stg)+lane;
      const double2 * tA = reinterpret_cast<const double2*>(stg+(size_t)Mpad*16)+lane;     // (cosh 4ReW, sinh 4ReW)
      const double2 * tB = reinterpret_cast<const double2*>(stg+(size_t)Mpad*32)+lane;     // (cos 4ImW, sin 4ImW)
      const bool up = sp[site] > 0;
      const float sgf = up ? 1.0f : -1.0f;
      // ---- fp32 product with its error bound
      float prod = 1.0f, esum = 0.0f;
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj)
      {
        const float4 tf = t32[32*jj];
        const float T1 = sf[jj].y*tf.x, T2 = sf[jj].z*tf.z, T3 = sf[jj].w*tf.w, T4 = sf[jj].x*tf.y;
        const float f = (T1+T2)+sgf*(T3-T4);
        const float mag = (T1+fabsf(T2))+(fabsf(T3)+fabsf(T4));
        esum += __fdividef(mag, fabsf(f));
        prod *= f;
      }
      float pm; int pe;
      split_me32(fmaxf(prod, 1e-37f), pm, pe);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        pm *= __shfl_xor_sync(0xffffffffu, pm, o);
        pe += __shfl_xor_sync(0xffffffffu, pe, o);
        esum += __shfl_xor_sync(0xffffffffu, esum, o);
      }
      // relative error of pm 2^pe against the exact product: 8 u per unit of |T|/f inside the factors, u per multiplication
      // (Mpad factors + 32 lane products + 5 butterfly stages), first order times 1.25 for the higher orders and the
      // approximate division; a bound above 0.05 (a factor cancelled almost completely, or prod under/overflowed) is not trusted
      const float delta = 1.25f*uerr*(8.0f*esum+(float)(Mpad+40));
      const double Afac = a.afac[2*site+(up ? 0 : 1)];
      int de = max(-2000, min(2000, pe-r0e));
      const double lhs = u*r0m, rhs = scalbn((double)pm*Afac, de);
      bool acc;
      const bool trust = (delta < 0.05f) && (prod > 1e-30f) && (prod < 1e30f) && (esum == esum);
      if (trust && lhs < rhs*(1.0-(double)delta)) acc = true;
      else if (trust && lhs > rhs*(1.0+1.1*(double)delta)) acc = false;      // (1/(1-delta) <= 1 + 1.06 delta below 0.05)
      else
      { // ---- exact: the fp64 flip product of rbm_sweep_fast_kernel from the fp64 state and tables (warp-uniform branch)
        ++n_slow;
        const double sg = up ? 1.0 : -1.0;
        double p64 = 1.0;
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const double2 sA = stA[lane+32*jj], sB = stB[lane+32*jj], Ta = tA[32*jj], Tb = tB[32*jj];
          const double A = fma(sB.x, Tb.x, sA.y*Ta.x);
          const double B = fma(sB.y, Tb.y, -(sA.x*Ta.y));
          p64 *= fma(sg, B, A);
        }
        double m64; int e64;
        split_me(fmax(p64, 1e-300), m64, e64);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          m64 *= __shfl_xor_sync(0xffffffffu, m64, o);
          e64 += __shfl_xor_sync(0xffffffffu, e64, o);
        }
        de = max(-2000, min(2000, e64-r0e));
        acc = (lhs < scalbn(m64*Afac, de));
      }
      if (lane == 0)
      {
        if (a.acc_log) a.acc_log[t_glob*a.K+k] = acc ? 1 : 0;
        rec[t] = acc ? (int8_t)(up ? 1 : -1) : (int8_t)0;
      }
      if (acc)
      { // ---- accepted: move the fp64 state (angle addition, as rbm_sweep_fast_kernel), refresh the fp32 copy, re-derive R0
        const double sg = up ? 1.0 : -1.0;
        const cd ai = avis[site];
        sa = cmake(sa.x-2.0*sg*ai.x, sa.y-2.0*sg*ai.y);
        any_acc = true;
        double r64 = 1.0;
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const int j = lane+32*jj;
          const double2 sA = stA[j], sB = stB[j], Ta = tA[32*jj], Tb = tB[32*jj];
          const double tys = sg*Ta.y, tbs = sg*Tb.y;
          const double S = fma(-sA.y, tys, sA.x*Ta.x), Ch = fma(-sA.x, tys, sA.y*Ta.x);
          const double cy = fma(sB.y, tbs, sB.x*Tb.x), sy = fma(-sB.x, tbs, sB.y*Tb.x);
          stA[j] = make_double2(S, Ch); stB[j] = make_double2(cy, sy);
          sf[jj] = make_float4((float)S, (float)Ch, (float)cy, (float)sy);
          r64 *= (Ch+cy);
        }
        double m64; int e64;
        split_me(fmax(r64, 1e-300), m64, e64);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
          m64 *= __shfl_xor_sync(0xffffffffu, m64, o);
          e64 += __shfl_xor_sync(0xffffffffu, e64, o);
        }
        double m2; int e2;
        split_me(m64, m2, e2);
        r0m = m2; r0e = e64+e2;
        if (lane == 0) sp[site] = (int8_t)(up ? -1 : 1);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty+slot);
    }
    // ---- (3) replay the accepted flips of this sweep on the exact theta, in order (bit-identical to the generic kernel)
    {
      cd th[JPL];
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj)
      {
        const int j = lane+32*jj;
        th[jj] = (j < M) ? a.theta[k*M+j] : cmake(0.0, 0.0);
      }
      int rp = pos_sweep0;
      for (int t = 0; t < N; ++t)
      {
        const int site = ord[rp];
        rp = (rp+1 == N) ? 0 : rp+1;
        const int r = rec[t];
        if (r == 0) continue;
        const cd * wrow = a.w2+(size_t)site*Mpad;
        const double s = (double)r;
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const cd wv = ld_tab(wrow+lane+32*jj);
          th[jj].x -= wv.x*s; th[jj].y -= wv.y*s;
        }
      }
      const bool last = (sweep+1 == a.nsweeps);
      cd lsum = cmake(0.0, 0.0);
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj)
      {
        const int j = lane+32*jj;
        if (j < M)
        {
          a.theta[k*M+j] = th[jj];
          if (last && any_acc) lsum = cadd(lsum, c_logcosh(th[jj]));
        }
      }
      if (last && any_acc) ln0 = cadd(warp_sum(lsum), sa);
      __syncwarp();
    }
  }
  for (int i = lane; i < N; i += 32) a.spins[k*N+i] = sp[i];
  if (lane == 0)
  {
    a.lnpsi0[k] = ln0;
    a.sa[k] = sa;
    if (any_acc) a.fresh[k] = 1;
    if (fa.stats != nullptr)
    {
      atomicAdd(fa.stats, (unsigned long long)t_end);
      atomicAdd(fa.stats+1, n_slow);
    }
  }
}
} // namespace nqs
