// Register-resident, transcendental-free kernels for the complex RBM (the configuration the headline metric is quoted on).
//
// Observation: for the RBM,  |psi(s^(i))/psi(s)|^2 = prod_j |cosh(theta_j - 2 s_i W_ij)|^2 / |cosh(theta_j)|^2 * exp(-4 s_i Re a_i)
// and |cosh(x+iy)|^2 = sinh^2 x + cos^2 y.  Keeping (sinh x, cosh x, cos y, sin y) per (chain, hidden unit) in REGISTERS and
// tabulating (cosh 2 Re W_ij, sinh 2 Re W_ij, cos 2 Im W_ij, sin 2 Im W_ij) once per parameter update turns every Metropolis
// proposal into 7 fp64 multiply-adds per hidden unit (angle-addition formulas, the spin enters as a sign-bit XOR) -- no exp /
// sincos / log inside the sweep (the generic kernel spends ~250 fp64 instructions per hidden unit per proposal on them).  The accept test  u < |psi'/psi0|^2  is evaluated on the products
// themselves in (mantissa, exponent) form, so it needs no log/exp either.  theta itself is NOT carried through the sweep:
// the accepted flips of a sweep are recorded and replayed afterwards on the exact fp64 theta in the reference's order
// (theta -= 2 s W_i, ref conditional_y_update, impl_neural_quantum_state.cuh:1314-1329), so theta stays bit-identical to the
// generic kernel's and the (sinh x, cosh x, cos y, sin y) registers are rebuilt from it at every sweep (no drift).
// lnpsi0 is recomputed from the final theta for chains that accepted at least once (it equals the lnpsi' of their last
// accepted proposal up to rounding); chains that never accepted keep their tracked -- possibly stale -- value, which is
// exactly the reference's behaviour.
//
// One warp owns C chains and ALL M hidden units (JPL = ceil(M/32) per lane), so every table entry fetched from L1 is used
// for C chains and the only cross-lane traffic is one shuffle butterfly per proposal.
#pragma once
#include "device_math.cuh"
#include "sampler_kernels.cuh"

namespace nqs
{
// Tables are split into 16-byte halves so that a warp's access is one fully coalesced LDG.128 per half (a 32-byte struct
// read with 8-byte loads costs 4x the L1 wavefronts -- measured, profiles/r1b_fast_kernels.md).
//   ftab_a = (cosh 4 Re W, sinh 4 Re W)   ftab_b = (cos 4 Im W, sin 4 Im W)
//   ctab_a = cosh(2W)                      ctab_b = sinh(2W)
typedef double2 FlipTab;
typedef double2 CoshTab;
__device__ __forceinline__ double2 ld_tab(const double2 * p) { return __ldg(p); }

// tables [N][Mpad] (Mpad = 32*JPL, neutral padding), rebuilt after every parameter change
__global__ void build_fast_tables_kernel(const int N, const int M, const int Mpad, const cd * __restrict__ params,
  FlipTab * __restrict__ ftab_a, FlipTab * __restrict__ ftab_b, CoshTab * __restrict__ ctab_a, CoshTab * __restrict__ ctab_b,
  CoshTab * __restrict__ ctabT_a, CoshTab * __restrict__ ctabT_b, const int Npad,
  cd * __restrict__ w2, double * __restrict__ afac, cd * __restrict__ aexp, double * __restrict__ bound, const int has_visible_bias,
  float4 * __restrict__ ftab32)
{
  const cd * W = params;
  const cd * a = params+(size_t)N*M;
  const long long total = (long long)N*Mpad;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x)
  {
    const int i = (int)(idx/Mpad), j = (int)(idx-(long long)i*Mpad);
    cd w = cmake(0.0, 0.0);
    if (j < M) w = W[(size_t)i*M+j];
    const double ex = exp(2.0*w.x), emx = exp(-2.0*w.x);
    double s, co;
    sincos(2.0*w.y, &s, &co);
    const double ch = 0.5*(ex+emx), sh = sinh(2.0*w.x);
    { // the sweep carries the DOUBLE angles (cosh 2x, sinh 2x, cos 2y, sin 2y): its flip tables are those of 4W
      double s4, c4;
      sincos(4.0*w.y, &s4, &c4);
      const double ch4 = cosh(4.0*w.x), sh4 = sinh(4.0*w.x);
      ftab_a[idx] = make_double2(ch4, sh4); ftab_b[idx] = make_double2(c4, s4);
      if (ftab32 != nullptr) ftab32[idx] = make_float4((float)ch4, (float)sh4, (float)c4, (float)s4);   // sweep_f32.cuh
    }
    ctab_a[idx] = make_double2(ch*co, sh*s); ctab_b[idx] = make_double2(sh*co, ch*s);
    w2[idx] = cmake(2.0*w.x, 2.0*w.y);
    if (j < M)
    {
      ctabT_a[(size_t)j*Npad+i] = make_double2(ch*co, sh*s);
      ctabT_b[(size_t)j*Npad+i] = make_double2(sh*co, ch*s);
    }
  }
  // neutral padding of the transposed tables (sites N..Npad-1): cosh = 1, sinh = 0
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < (long long)M*(Npad-N); idx += (long long)gridDim.x*blockDim.x)
  {
    const int j = (int)(idx/(Npad-N)), i = N+(int)(idx-(long long)j*(Npad-N));
    ctabT_a[(size_t)j*Npad+i] = make_double2(1.0, 0.0);
    ctabT_b[(size_t)j*Npad+i] = make_double2(0.0, 0.0);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && bound != nullptr) *bound = 0.0;   // theta_bound_kernel (launched next) accumulates a maximum
  for (int i = blockIdx.x*blockDim.x+threadIdx.x; i < N && has_visible_bias; i += gridDim.x*blockDim.x)
  {
    const cd ai = a[i];
    afac[2*i] = exp(-4.0*ai.x);   // sigma = +1
    afac[2*i+1] = exp(4.0*ai.x);  // sigma = -1
    aexp[2*i] = c_exp(cmake(-2.0*ai.x, -2.0*ai.y));
    aexp[2*i+1] = c_exp(cmake(2.0*ai.x, 2.0*ai.y));
  }
}

// bound[0] = max_j (|Re b_j| + sum_i |Re W_ij|) >= max |Re theta|: decides whether the product form cannot overflow.
// Grid of column tiles (32 hidden units x 8 row groups per CTA): the single-CTA version walked the N rows of W serially, one L2
// round trip each (22 us at N = 128, on the critical path of every parameter update).  bound[0] must be zero at launch
// (build_fast_tables_kernel, which always runs just before, clears it); non-negative doubles order like their bit patterns,
// so the maximum is an integer atomicMax.
#define NQS_TB_THREADS 256
__global__ void __launch_bounds__(NQS_TB_THREADS) theta_bound_kernel(const int N, const int M, const cd * __restrict__ params, double * __restrict__ bound)
{
  __shared__ double sh[8][32];
  const cd * W = params;
  const cd * b = params+(size_t)N*M+N;
  const int jl = threadIdx.x&31, ig = threadIdx.x>>5;
  const int j = blockIdx.x*32+jl;
  double s = 0.0;
  if (j < M)
    for (int i = ig; i < N; i += 8) s += fabs(W[(size_t)i*M+j].x);
  sh[ig][jl] = s;
  __syncthreads();
  if (ig == 0)
  {
    double t = (j < M) ? fabs(b[j].x) : 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sh[q][jl];
    if (!(t == t)) t = __longlong_as_double(0x7ff0000000000000ll);     // NaN parameters: report +inf (the product form is refused)
    for (int o = 16; o > 0; o >>= 1) t = fmax(t, __shfl_xor_sync(0xffffffffu, t, o));
    if (jl == 0) atomicMax(reinterpret_cast<unsigned long long*>(bound), (unsigned long long)__double_as_longlong(t));
  }
}

// split a positive finite double into mantissa in [1,2) and exponent
__device__ __forceinline__ void split_me(const double p, double & m, int & e)
{
  const int hi = __double2hiint(p);
  e = ((hi>>20)&0x7ff)-1023;
  m = __hiloint2double((hi&0x800fffff)|0x3ff00000, __double2loint(p));
}

__device__ __forceinline__ double flip_sign(const double v, const int mask)
{ // v * (+-1) as one integer XOR on the sign bit
  return __hiloint2double(__double2hiint(v)^mask, __double2loint(v));
}

struct FastSweepArgs
{
  int N, M, Mpad;
  long long K;
  const cd * params;
  const FlipTab * ftab_a;
  const FlipTab * ftab_b;
  const cd * w2;
  const double * afac;
  int8_t * spins;
  cd * theta;
  cd * lnpsi0;
  cd * sa;
  unsigned char * fresh;   // [K] 1 = lnpsi0 equals lnpsi(current theta) (set here when the chain accepted at least once)
  const int * order;
  int pos0;
  int nsweeps;
  const double * uniforms;
  unsigned long long seed, step0;
  long long chain_offset;
  unsigned char * acc_log;
};

// The flip-table rows of the next proposals are staged in shared memory by TMA bulk copies (one 2 x Mpad x 16 B row pair per
// proposal, shared by the warps of the CTA) through a ring of NQS_SW_STAGES slots with full / empty mbarriers: the sweep walks
// a fixed site order, so the loads are issued NQS_SW_STAGES - 1 proposals ahead and the table reads in the loop are LDS.128
// instead of L1-missing LDG (29 % long-scoreboard stalls before, profiles/r1e_sampler_full_summary.md).
#define NQS_SW_STAGES 4
#define NQS_SW_XCHG_BYTES 512   // WPC = 2: per-proposal (mantissa, exponent) exchange [2 parities][4 pairs][2] + final log cosh sums
inline size_t fast_sweep_smem_bytes(int N, int C, int warps, int Mpad)
{
  const size_t npad = (size_t)((N+15)/16)*16;
  const size_t head = ((size_t)warps*C*npad+(size_t)warps*C*N+(size_t)N*sizeof(int)+15)/16*16;
  return head+(size_t)NQS_SW_STAGES*2*Mpad*sizeof(FlipTab)+(size_t)2*NQS_SW_STAGES*sizeof(uint64_t)+NQS_SW_XCHG_BYTES;
}

// (mantissa, exponent) products of C chains reduced over the 32 lanes with a TRANSPOSING butterfly: after the first
// log2(C) stages every lane keeps only the chain of its own lane group (32/C consecutive lanes), so a C-chain reduction
// costs log2(C) + ... + 5 value shuffles instead of 5*C.  On return lane l holds the total of chain l / (32/C).
template <int C>
__device__ __forceinline__ void reduce_me_transposed(double (&m)[C], int (&e)[C], const int lane, double & mo, int & eo)
{
  static_assert(C == 1 || C == 2 || C == 4, "C must be 1, 2 or 4");
  double km[2]; int ke[2];
  if (C == 4)
  {
    const bool up = (lane&16) != 0;
#pragma unroll
    for (int q = 0; q < 2; ++q)
    {
      const double sm = up ? m[q] : m[2+q];
      const int se = up ? e[q] : e[2+q];
      km[q] = (up ? m[2+q] : m[q])*__shfl_xor_sync(0xffffffffu, sm, 16);
      ke[q] = (up ? e[2+q] : e[q])+__shfl_xor_sync(0xffffffffu, se, 16);
    }
    const bool up8 = (lane&8) != 0;
    const double sm = up8 ? km[0] : km[1];
    const int se = up8 ? ke[0] : ke[1];
    mo = (up8 ? km[1] : km[0])*__shfl_xor_sync(0xffffffffu, sm, 8);
    eo = (up8 ? ke[1] : ke[0])+__shfl_xor_sync(0xffffffffu, se, 8);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
    {
      mo *= __shfl_xor_sync(0xffffffffu, mo, o);
      eo += __shfl_xor_sync(0xffffffffu, eo, o);
    }
  }
  else if (C == 2)
  {
    const bool up = (lane&16) != 0;
    const double sm = up ? m[0] : m[C-1];
    const int se = up ? e[0] : e[C-1];
    mo = (up ? m[C-1] : m[0])*__shfl_xor_sync(0xffffffffu, sm, 16);
    eo = (up ? e[C-1] : e[0])+__shfl_xor_sync(0xffffffffu, se, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1)
    {
      mo *= __shfl_xor_sync(0xffffffffu, mo, o);
      eo += __shfl_xor_sync(0xffffffffu, eo, o);
    }
  }
  else
  {
    mo = m[0]; eo = e[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      mo *= __shfl_xor_sync(0xffffffffu, mo, o);
      eo += __shfl_xor_sync(0xffffffffu, eo, o);
    }
  }
}

// State per (chain, hidden unit), in registers: the DOUBLE angles (cosh 2x, sinh 2x, cos 2y, sin 2y) of theta = x + i y, because
//   2 |cosh theta|^2 = cosh 2x + cos 2y .
// Table per (site, hidden unit): ftab_a = (cosh 4ReW, sinh 4ReW), ftab_b = (cos 4ImW, sin 4ImW).  A flip of a spin sigma maps
// theta -> theta - 2 sigma W, i.e. 2x -> 2x - 4 sigma ReW, 2y -> 2y - 4 sigma ImW:
//   2 |cosh theta'|^2 = [cosh 2x cosh 4ReW + cos 2y cos 4ImW] + sigma [sin 2y sin 4ImW - sinh 2x sinh 4ReW] = A + sigma B
// sigma enters as ONE fp64 register per chain in the last FMA: 6 fp64 instructions and no integer work per (proposal, chain,
// unit).  (The earlier (sinh x, cosh x, cos y, sin y) form needed 7 + two sign-bit XORs that each cost a LOP3 and a MOV to
// rebuild the register pair: as many integer as fp64 instructions in the loop -- SASS histogram, profiles/r1h_sweep_sass.md.)
// Every factor carries the constant 2, so the product of a chain carries 2^Mpad: the tracked reference product R0 starts with
// the same factor and inherits it at every accept, and the ratio never sees it.
// CTA shapes: one chain per warp with at most 8 hidden-unit slots per lane fits 128 registers and runs 8 warps per CTA (16 per SM);
// everything else runs 4 warps per CTA, 2 CTAs per SM, at up to 255 registers.
// WPC = 2 (M up to 1024): TWO warps share one chain, each holding 16 hidden-unit slots per lane (512 units); per proposal
// they exchange their (mantissa, exponent) partial products through shared memory behind a 64-thread named barrier, take the
// same accept decision from the same numbers, and update their own half of the state.  8 warps = 4 chains per CTA, 1 CTA per SM.
template <int JPL, int C, int WPC> struct SweepShape
{
  static constexpr int warps = (WPC > 1 || (C == 1 && JPL <= 8)) ? 8 : 4;
  static constexpr int min_ctas = (WPC > 1) ? 1 : 2;
};
template <int JPL, int C, int WPC = 1>
__global__ void __launch_bounds__(32*SweepShape<JPL, C, WPC>::warps, SweepShape<JPL, C, WPC>::min_ctas) rbm_sweep_fast_kernel(const FastSweepArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  static_assert(WPC == 1 || C == 1, "warp pairs hold one chain");
  constexpr int G = 32/C;                      // lanes per chain group (owner lanes of a chain's accept decision)
  const int warps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  const int wsub = (WPC > 1) ? (w%WPC) : 0, wgrp = w/WPC;     // which part of the hidden layer / which chain group of the CTA
  const int joff = 32*JPL*wsub;                // first hidden unit of this warp's part
  const int N = a.N, M = a.M, Mpad = a.Mpad;
  const int npad = ((N+15)/16)*16;
  int8_t * sp = reinterpret_cast<int8_t*>(smem_raw)+(size_t)w*C*npad;                         // [C][npad]
  int8_t * rec = reinterpret_cast<int8_t*>(smem_raw)+(size_t)warps*C*npad+(size_t)w*C*N;      // [N][C]: 0 rejected, +-1 = accepted flip of a spin that was +-1
  int * ord = reinterpret_cast<int*>(smem_raw+(size_t)warps*C*npad+(size_t)warps*C*N);
  const size_t head = ((size_t)warps*C*npad+(size_t)warps*C*N+(size_t)N*sizeof(int)+15)/16*16;
  FlipTab * stage0 = reinterpret_cast<FlipTab*>(smem_raw+head);                      // [NQS_SW_STAGES][2][Mpad]
  uint64_t * full = reinterpret_cast<uint64_t*>(stage0+(size_t)NQS_SW_STAGES*2*Mpad); // [NQS_SW_STAGES] TMA arrival
  uint64_t * empty = full+NQS_SW_STAGES;                                              // [NQS_SW_STAGES] every active warp is done
  double * xm = reinterpret_cast<double*>(empty+NQS_SW_STAGES);                       // [2][4][2] partial mantissas (WPC = 2)
  int * xe = reinterpret_cast<int*>(xm+16);                                           // [2][4][2] partial exponents
  cd * xl = reinterpret_cast<cd*>(xe+16);                                             // [4][2] partial log cosh sums
  for (int i = threadIdx.x; i < N; i += blockDim.x) ord[i] = a.order[i];
  const long long kblock = (long long)blockIdx.x*(warps/WPC)*C;
  if (threadIdx.x == 0)
  {
    long long nact = ((a.K-kblock+C-1)/C)*WPC;            // warps of this CTA that own at least one chain
    if (nact > warps) nact = warps;
    if (nact < 1) nact = 1;
    for (int q = 0; q < NQS_SW_STAGES; ++q) { mbar_init(full+q, 1); mbar_init(empty+q, (uint32_t)nact); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long kbase = ((long long)blockIdx.x*(warps/WPC)+wgrp)*C;
  if (kbase >= a.K) return;
  const cd * avis = a.params+(size_t)N*M;
  const int myc = lane/G;                      // the chain this lane decides for
  const bool my_valid = (kbase+myc < a.K);
  const long long my_k = my_valid ? kbase+myc : kbase;
  bool valid[C];
  cd ln0[C], sa[C];
  bool any_acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    valid[c] = (kbase+c < a.K);
    const long long k = valid[c] ? kbase+c : kbase;
    for (int i = lane; i < N; i += 32) sp[c*npad+i] = a.spins[k*N+i];
    ln0[c] = a.lnpsi0[k]; sa[c] = a.sa[k];
    any_acc[c] = false;
  }
  // R0 = exp(2 (Re lnpsi0 - Re sa)) = prod_j |cosh theta_j|^2 of the TRACKED amplitude, as m * 2^e, kept by the owner lanes
  double r0m; int r0e;
  {
    const cd l0 = a.lnpsi0[my_k], s0 = a.sa[my_k];
    const double v = 2.0*(l0.x-s0.x)*1.4426950408889634;
    const double fl = floor(v);
    r0m = exp2(v-fl);
    r0e = (int)fmax(fmin(fl, 100000.0), -100000.0)+Mpad;   // every factor of the products below carries a 2
  }
  __syncwarp();
  int pos = a.pos0;
  long long t_glob = 0;
  const long long t_end = (long long)a.nsweeps*N;
  const uint32_t row_bytes = (uint32_t)(Mpad*sizeof(FlipTab));
  // producer (lane 0 of warp 0): table rows of proposal q -> slot q % NQS_SW_STAGES
  auto issue_rows = [&](const long long q)
  {
    const int sq = ord[(int)(((long long)a.pos0+q)%N)];
    const int slot = (int)(q%NQS_SW_STAGES);
    FlipTab * dst = stage0+(size_t)slot*2*Mpad;
    mbar_expect_tx(full+slot, 2*row_bytes);
    tma_load_1d(dst, a.ftab_a+(size_t)sq*Mpad, row_bytes, full+slot);
    tma_load_1d(dst+Mpad, a.ftab_b+(size_t)sq*Mpad, row_bytes, full+slot);
  };
  if (w == 0 && lane == 0)
    for (long long q = 0; q < NQS_SW_STAGES-1 && q < t_end; ++q) issue_rows(q);
  double ubuf = 0.0, unext = 0.0;              // uniform of proposal (t_glob rounded down to G) + lane%G of chain myc; next group's

  for (int sweep = 0; sweep < a.nsweeps; ++sweep)
  {
    // ---- (1) rebuild the multiplicative state from the exact theta
    double S[C][JPL], Ch[C][JPL], cy[C][JPL], sy[C][JPL];     // sinh 2x, cosh 2x, cos 2y, sin 2y
#pragma unroll
    for (int c = 0; c < C; ++c)
    {
      const long long k = valid[c] ? kbase+c : kbase;
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj)
      {
        const int j = joff+lane+32*jj;
        cd th = cmake(0.0, 0.0);
        if (j < M) th = a.theta[k*M+j];
        const double ex = exp(2.0*th.x), emx = 1.0/ex;
        S[c][jj] = 0.5*(ex-emx); Ch[c][jj] = 0.5*(ex+emx);
        sincos(2.0*th.y, &sy[c][jj], &cy[c][jj]);
      }
    }
    const int pos_sweep0 = pos;
    // ---- (2) N proposals
    for (int t = 0; t < N; ++t, ++t_glob)
    {
      if ((t_glob&(G-1)) == 0)
      { // the uniforms of this group of G proposals were fetched one group ahead (a pre-drawn feed may live in pinned HOST
        // memory: its latency must not sit on the accept decision); fetch the next group's now
        if (a.uniforms)
        {
          if (t_glob == 0)
          {
            const long long tt = (long long)(lane&(G-1));
            unext = (tt < t_end) ? a.uniforms[tt*a.K+my_k] : 0.0;
          }
          ubuf = unext;
          const long long tn = t_glob+G+(lane&(G-1));
          unext = (tn < t_end) ? a.uniforms[tn*a.K+my_k] : 0.0;
        }
        else
        {
          const long long tt = t_glob+(lane&(G-1));
          if (tt < t_end) ubuf = philox_uniform(a.seed, (unsigned long long)(a.chain_offset+my_k), a.step0+(unsigned long long)tt);
        }
      }
      const int site = ord[pos];
      pos = (pos+1 == N) ? 0 : pos+1;
      const int slot = (int)(t_glob&(NQS_SW_STAGES-1));
      if (w == 0 && lane == 0)
      { // keep NQS_SW_STAGES - 1 proposals in flight: the slot of proposal t_glob - 1 is refilled once every warp released it
        const long long q = t_glob+NQS_SW_STAGES-1;
        if (q < t_end)
        {
          if (t_glob > 0) mbar_wait(empty+(int)(q&(NQS_SW_STAGES-1)), (uint32_t)(((t_glob-1)/NQS_SW_STAGES)&1));
          issue_rows(q);
        }
      }
      mbar_wait(full+slot, (uint32_t)((t_glob/NQS_SW_STAGES)&1));
      const FlipTab * trow_a = stage0+(size_t)slot*2*Mpad+joff+lane;
      const FlipTab * trow_b = trow_a+Mpad;
      int smask[C];                              // sign bit of sigma
      double sg[C];                              // sigma
      double prod[C];
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        smask[c] = (sp[c*npad+site] < 0) ? (int)0x80000000 : 0;
        sg[c] = smask[c] ? -1.0 : 1.0;
        prod[c] = 1.0;
      }
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj)
      {
        const double2 Ta = trow_a[32*jj], Tb = trow_b[32*jj];
#pragma unroll
        for (int c = 0; c < C; ++c)
        {
          const double A = fma(cy[c][jj], Tb.x, Ch[c][jj]*Ta.x);
          const double B = fma(sy[c][jj], Tb.y, -(S[c][jj]*Ta.y));
          prod[c] *= fma(sg[c], B, A);
        }
      }
      double pm[C]; int pe[C];
#pragma unroll
      for (int c = 0; c < C; ++c) split_me(fmax(prod[c], 1e-300), pm[c], pe[c]);
      double m; int e;
      reduce_me_transposed<C>(pm, pe, lane, m, e);
      if (WPC > 1)
      { // the two halves of the hidden layer: both warps read both partials and combine them in the same order
        const int xq = ((int)(t_glob&1)*4+wgrp)*2;
        if (lane == 0) { xm[xq+wsub] = m; xe[xq+wsub] = e; }
        asm volatile("bar.sync %0, %1;" :: "r"(1+wgrp), "r"(32*WPC) : "memory");
        m = xm[xq]*xm[xq+1];
        e = xe[xq]+xe[xq+1];
      }
      // accept  <=>  u < P' A / R0   (== u < exp(2 (Re lnpsi' - Re lnpsi0)), ref impl_mcmc_sampler.cuh:75-99), decided by the owner lanes
      const double u = __shfl_sync(0xffffffffu, ubuf, (lane&~(G-1))|(int)(t_glob&(G-1)));
      const bool my_up = sp[myc*npad+site] > 0;
      const double A = a.afac[2*site+(my_up ? 0 : 1)];
      int de = e-r0e;
      de = max(-2000, min(2000, de));
      const bool my_acc = my_valid && (u*r0m < scalbn(m*A, de));
      const unsigned int bal = __ballot_sync(0xffffffffu, my_acc);
      if (my_acc)
      { // renormalise (m in [1, 2^32)) and adopt as the new reference product
        double m2; int e2;
        split_me(m, m2, e2);
        r0m = m2; r0e = e+e2;
      }
      bool acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        acc[c] = ((bal>>(c*G))&1u) != 0;
        if (lane == 0)
        {
          if (a.acc_log && valid[c] && wsub == 0) a.acc_log[t_glob*a.K+kbase+c] = acc[c] ? 1 : 0;
          rec[t*C+c] = acc[c] ? (int8_t)(smask[c] ? -1 : 1) : (int8_t)0;
        }
      }
      if (bal != 0u)
      { // (with the table rows in shared memory a warp-uniform branch per accepted chain beats predication: no select / move per state register)
        const cd ai = avis[site];
#pragma unroll
        for (int c = 0; c < C; ++c)
          if (acc[c])
          {
            const double two_s = smask[c] ? -2.0 : 2.0;
            sa[c] = cmake(sa[c].x-two_s*ai.x, sa[c].y-two_s*ai.y);
            any_acc[c] = true;
          }
        // warp-uniform branch per accepted chain: in-place state update without predicated moves (tables re-read from smem)
#pragma unroll
        for (int c = 0; c < C; ++c)
        {
          if (acc[c])
          {
#pragma unroll
            for (int jj = 0; jj < JPL; ++jj)
            {
              const double2 Ta = trow_a[32*jj], Tb = trow_b[32*jj];
              const double s0 = S[c][jj], c0 = Ch[c][jj], y0 = cy[c][jj], y1 = sy[c][jj];
              const double tys = sg[c]*Ta.y, tbs = sg[c]*Tb.y;       // sigma folded into the table entry: 10 fp64 per unit
              S[c][jj] = fma(-c0, tys, s0*Ta.x);
              Ch[c][jj] = fma(-s0, tys, c0*Ta.x);
              cy[c][jj] = fma(y1, tbs, y0*Tb.x);
              sy[c][jj] = fma(-y0, tbs, y1*Tb.x);
            }
          }
        }
        __syncwarp();
        if (lane == 0)
        {
#pragma unroll
          for (int c = 0; c < C; ++c)
            if (acc[c]) sp[c*npad+site] = (int8_t)(smask[c] ? 1 : -1);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty+slot);      // this warp is done with the table rows of the proposal
    }
    // ---- (3) replay the accepted flips of this sweep on the exact theta, in order (bit-identical to the generic kernel)
    {
      cd th[C][JPL];
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        const long long k = valid[c] ? kbase+c : kbase;
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const int j = joff+lane+32*jj;
          th[c][jj] = (j < M) ? a.theta[k*M+j] : cmake(0.0, 0.0);
        }
      }
      int rp = pos_sweep0;
      for (int t = 0; t < N; ++t)
      {
        const int site = ord[rp];
        rp = (rp+1 == N) ? 0 : rp+1;
        int r[C];
        bool any = false;
#pragma unroll
        for (int c = 0; c < C; ++c) { r[c] = rec[t*C+c]; any = any || (r[c] != 0); }
        if (!any) continue;
        const cd * wrow = a.w2+(size_t)site*Mpad;
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const cd wv = ld_tab(wrow+joff+lane+32*jj);
#pragma unroll
          for (int c = 0; c < C; ++c)
          {
            if (r[c] != 0)
            { // theta -= W * (2 sigma): w2 = 2W is exact, so this equals the reference's y - w*(2*s) bit for bit
              const double s = (double)r[c];
              th[c][jj].x -= wv.x*s; th[c][jj].y -= wv.y*s;
            }
          }
        }
      }
      const bool last = (sweep+1 == a.nsweeps);
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        if (!valid[c]) continue;
        const long long k = kbase+c;
        cd lsum = cmake(0.0, 0.0);
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const int j = joff+lane+32*jj;
          if (j < M)
          {
            a.theta[k*M+j] = th[c][jj];
            if (last && any_acc[c]) lsum = cadd(lsum, c_logcosh(th[c][jj]));
          }
        }
        if (last && any_acc[c])
        {
          lsum = warp_sum(lsum);
          if (WPC > 1)
          { // any_acc is the same in both warps of the pair (same decisions): both reach the barrier
            if (lane == 0) xl[wgrp*2+wsub] = lsum;
            asm volatile("bar.sync %0, %1;" :: "r"(1+wgrp), "r"(32*WPC) : "memory");
            lsum = cadd(xl[wgrp*2], xl[wgrp*2+1]);
          }
          ln0[c] = cadd(lsum, sa[c]);
        }
      }
      __syncwarp();
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    if (!valid[c] || wsub != 0) continue;
    const long long k = kbase+c;
    for (int i = lane; i < N; i += 32) a.spins[k*N+i] = sp[c*npad+i];
    if (lane == 0)
    {
      a.lnpsi0[k] = ln0[c];
      a.sa[k] = sa[c];
      if (any_acc[c]) a.fresh[k] = 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Local energy, RBM:  psi(s^(i))/psi(s) = prod_j [cosh(2W_ij) - s_i tanh(theta_j) sinh(2W_ij)] * exp(-2 s_i a_i)
// (cosh(t - 2sW) = cosh t cosh 2W - s sinh t sinh 2W).  The N flips of a configuration are independent, so LANES RUN OVER
// SITES: lane l of a warp owns site i = 32*block + l and walks all hidden units j serially, multiplying its own complex
// product -- no cross-lane reduction per flip at all.  tanh(theta_j) of the CTA's C chains sits in shared memory (one
// broadcast LDS per (j, chain)); the table is stored TRANSPOSED ([j][i], i contiguous) so that each (j, 32 sites) access is a
// coalesced LDG.128 shared by the C chains.  8 fp64 FMAs per (site, hidden unit, chain).  The reference evaluates N full
// forward(i) passes (k3+k4+c1) + k11 per call (impl_hamiltonians.cuh:233-238).  The tracked lnpsi0 enters as in the
// reference: exp(lnpsi' - lnpsi0) = ratio * exp(lnpsi(theta) - lnpsi0); the second factor is 1 for chains flagged fresh and
// is evaluated explicitly (M log cosh) only for stale ones.
// ---------------------------------------------------------------------------------------------------------------------
struct FastElocArgs
{
  int N, M, Npad;            // Npad = N rounded up to 32 (neutral table padding)
  long long K;
  const CoshTab * ctabT_a;   // [M][Npad] cosh(2 W_ij)
  const CoshTab * ctabT_b;   // [M][Npad] sinh(2 W_ij)
  const cd * aexp;           // [N][2] exp(-+2 a_i)
  const int8_t * spins;
  const cd * theta;
  const cd * lnpsi0;
  const cd * sa;
  const unsigned char * fresh;
  const double * Jmat;
  const double * sjs;        // [K] sum_ij s_i J_ij s_j from the tensor-core GEMM (sv_struct.cuh), or null: computed here
  double hfield;
  cd * htilda;
};

inline size_t fast_eloc_smem_bytes(int N, int M, int C)
{
  const size_t npad = (size_t)((N+15)/16)*16;
  return (size_t)C*M*sizeof(cd)+(size_t)C*npad+(size_t)C*8*6*sizeof(double)+16;
}

template <int C>
__global__ void __launch_bounds__(256) rbm_eloc_sites_kernel(const FastElocArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, M = a.M, Npad = a.Npad;
  const int npad = ((N+15)/16)*16;
  const int nwarps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  cd * Tsh = reinterpret_cast<cd*>(smem_raw);                                   // [C][M] tanh(theta)
  int8_t * sp = reinterpret_cast<int8_t*>(Tsh+(size_t)C*M);                     // [C][npad]
  double * red = reinterpret_cast<double*>(smem_raw+(size_t)C*M*sizeof(cd)+(size_t)C*npad); // [C][8 warps][4]
  const long long kbase = (long long)blockIdx.x*C;
  // ---- phase 0: tanh(theta), spins, staleness
  bool stale = false;
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    const long long k = (kbase+c < a.K) ? kbase+c : kbase;
    stale = stale || (a.fresh[k] == 0);
  }
  double ls_x[C], ls_y[C];
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    const long long k = (kbase+c < a.K) ? kbase+c : kbase;
    ls_x[c] = 0.0; ls_y[c] = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x)
    {
      const cd th = a.theta[k*M+j];
      Tsh[c*M+j] = c_tanh(th);
      if (stale) { const cd lc = c_logcosh(th); ls_x[c] += lc.x; ls_y[c] += lc.y; }
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) sp[c*npad+i] = a.spins[k*N+i];
  }
  __syncthreads();
  // ---- phase 1: per-thread partial sums: [0] diag, [1..2] sum_i ratio_i, and (stale only) [3..] log cosh sums
  double part_d[C], part_x[C], part_y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { part_d[c] = 0.0; part_x[c] = 0.0; part_y[c] = 0.0; }
  // 1/2 sum_ij s_i J_ij s_j  (ref c5 + k10, impl_hamiltonians.cuh:226-231,871-887): threads over i
  for (int i = threadIdx.x; i < N && a.sjs == nullptr; i += blockDim.x)
  {
    const double * Jrow = a.Jmat+(size_t)i*N;
    double sj[C];
#pragma unroll
    for (int c = 0; c < C; ++c) sj[c] = 0.0;
    for (int j = 0; j < N; ++j)
    {
      const double Jv = __ldg(Jrow+j);
#pragma unroll
      for (int c = 0; c < C; ++c) sj[c] = fma(Jv, (double)sp[c*npad+j], sj[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) part_d[c] = fma(sj[c], (double)sp[c*npad+i], part_d[c]);
  }
  // N single-flip ratios: warp w takes site blocks w, w+nwarps, ...
  for (int sb = w; sb*32 < N; sb += nwarps)
  {
    const int i = sb*32+lane;
    const bool ok = (i < N);
    int smask[C];
#pragma unroll
    for (int c = 0; c < C; ++c) smask[c] = (ok && sp[c*npad+i] < 0) ? (int)0x80000000 : 0;
    double sg[C];                                // sigma of this lane's site: one fp64 register instead of sign-bit XORs
#pragma unroll
    for (int c = 0; c < C; ++c) sg[c] = smask[c] ? -1.0 : 1.0;
    double p0x[C], p0y[C], p1x[C], p1y[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { p0x[c] = 1.0; p0y[c] = 0.0; p1x[c] = 1.0; p1y[c] = 0.0; }
    const CoshTab * ca = a.ctabT_a+sb*32+lane;
    const CoshTab * cb = a.ctabT_b+sb*32+lane;
    int j = 0;
    for (; j+1 < M; j += 2)
    {
      const double2 c0 = ld_tab(ca+(size_t)j*Npad), s0 = ld_tab(cb+(size_t)j*Npad);
      const double2 c1 = ld_tab(ca+(size_t)(j+1)*Npad), s1 = ld_tab(cb+(size_t)(j+1)*Npad);
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        const cd t0 = Tsh[c*M+j], t1 = Tsh[c*M+j+1];
        { // f = cosh2W - sigma (tanh sinh2W) ; p0 *= f
          const double gr = fma(-t0.y, s0.y, t0.x*s0.x), gi = fma(t0.y, s0.x, t0.x*s0.y);
          const double fr = fma(-sg[c], gr, c0.x), fi = fma(-sg[c], gi, c0.y);
          const double nr = fma(p0x[c], fr, -p0y[c]*fi), ni = fma(p0x[c], fi, p0y[c]*fr);
          p0x[c] = nr; p0y[c] = ni;
        }
        {
          const double gr = fma(-t1.y, s1.y, t1.x*s1.x), gi = fma(t1.y, s1.x, t1.x*s1.y);
          const double fr = fma(-sg[c], gr, c1.x), fi = fma(-sg[c], gi, c1.y);
          const double nr = fma(p1x[c], fr, -p1y[c]*fi), ni = fma(p1x[c], fi, p1y[c]*fr);
          p1x[c] = nr; p1y[c] = ni;
        }
      }
    }
    if (j < M)
    {
      const double2 c0 = ld_tab(ca+(size_t)j*Npad), s0 = ld_tab(cb+(size_t)j*Npad);
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        const cd t0 = Tsh[c*M+j];
        const double gr = fma(-t0.y, s0.y, t0.x*s0.x), gi = fma(t0.y, s0.x, t0.x*s0.y);
        const double fr = fma(-sg[c], gr, c0.x), fi = fma(-sg[c], gi, c0.y);
        const double nr = fma(p0x[c], fr, -p0y[c]*fi), ni = fma(p0x[c], fi, p0y[c]*fr);
        p0x[c] = nr; p0y[c] = ni;
      }
    }
    if (ok)
    {
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        const cd pr = cmul(cmake(p0x[c], p0y[c]), cmake(p1x[c], p1y[c]));
        const cd ae = a.aexp[2*i+(smask[c] ? 1 : 0)];
        const cd ratio = cmul(pr, ae);
        part_x[c] += ratio.x; part_y[c] += ratio.y;
      }
    }
  }
  // ---- phase 2: block reduction (fixed order) and output
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    const double d = warp_sum(part_d[c]), x = warp_sum(part_x[c]), y = warp_sum(part_y[c]);
    const double lx = stale ? warp_sum(ls_x[c]) : 0.0, ly = stale ? warp_sum(ls_y[c]) : 0.0;
    if (lane == 0)
    {
      double * r = red+((size_t)c*8+w)*4;
      r[0] = d; r[1] = x; r[2] = y;
      red[(size_t)C*8*4+((size_t)c*8+w)*2] = lx; red[(size_t)C*8*4+((size_t)c*8+w)*2+1] = ly;
    }
  }
  __syncthreads();
  if (threadIdx.x < C && kbase+threadIdx.x < a.K)
  {
    const int c = threadIdx.x;
    double d = 0, x = 0, y = 0, lx = 0, ly = 0;
    for (int ww = 0; ww < nwarps; ++ww)
    {
      const double * r = red+((size_t)c*8+ww)*4;
      d += r[0]; x += r[1]; y += r[2];
      lx += red[(size_t)C*8*4+((size_t)c*8+ww)*2]; ly += red[(size_t)C*8*4+((size_t)c*8+ww)*2+1];
    }
    if (a.sjs != nullptr) d = a.sjs[kbase+c];
    cd off = cmake(x, y);
    if (stale)
    { // exp(lnpsi(theta) - lnpsi0_tracked): != 1 only right after warm_up's quirk flip or a parameter update
      const long long k = kbase+c;
      const cd l0 = a.lnpsi0[k], s0 = a.sa[k];
      off = cmul(off, c_exp(cmake(lx+s0.x-l0.x, ly+s0.y-l0.y)));
    }
    a.htilda[kbase+c] = cmake((0.5*d+a.hfield*off.x)/(double)N, (a.hfield*off.y)/(double)N);
  }
}
} // namespace nqs
