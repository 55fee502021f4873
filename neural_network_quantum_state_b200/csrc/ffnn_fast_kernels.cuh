// Sampler-side kernels for the one-hidden-layer complex FNN with the hidden-unit state RESIDENT on chip.
//
// ref: FFNN::forward(int) (gpu/include/impl_neural_quantum_state.cuh:820-830: k3 logcosh over K*M + Zgemm with w1o) called once
// per proposal by BaseParallelSampler::do_mcmc_steps (impl_mcmc_sampler.cuh:28-39) and N times per chain by
// LITFIChain::get_htilda_ (impl_hamiltonians.cuh:233-238); FFNN::spin_flip (:880-890) keeps y = theta by rank-1 updates.
//
// ln psi = sum_j w_j log cosh theta_j has no product form (complex weights), but the CHANGE of one term under a flip of spin
// sigma at site i needs no exp / sincos:
//   log cosh(theta_j - 2 sigma W_ij) - log cosh theta_j = log f_j ,   f_j = cosh 2W_ij - sigma tanh(theta_j) sinh 2W_ij ,
// one complex log per hidden unit -- a branch-free series, because f_j is close to 1 (c_log_near_one) -- instead of the four
// transcendentals of a fresh log cosh, with
// tanh(theta_j) resident per chain and (cosh 2W, sinh 2W) tabulated once per parameter update (build_fast_tables_kernel).
// Complex weights make the BRANCH of Im log cosh matter (2 pi i w_j is not a multiple of 2 pi i): the reference takes the
// principal atan2 of each log cosh separately, so the resident state also carries A_j = Im log cosh theta_j and the new value is
// the principal value of A_j + arg f_j.
// On an accepted flip the resident state moves multiplicatively, tanh' = (tanh cosh 2W - sigma sinh 2W) / f, A' as above, while
// the exact theta (registers) takes the reference's rank-1 update theta -= 2 sigma W_i (ref conditional_y_update :1314-1329) with
// the same operands in the same order, so it stays bit-identical to the generic kernel's; tanh / A / ln psi are rebuilt from it
// at every sweep, so nothing drifts.
#pragma once
#include "device_math.cuh"
#include "fast_kernels.cuh"

namespace nqs
{
#define NQS_FF_STAGES 2
#define NQS_FF_MAX_WARPS 12
#define NQS_FF_PI 3.14159265358979323846

struct FfnnSweepArgs
{
  int N, M, Mpad;
  long long K;
  const cd * params;         // [W (i*M+j) | b1 | w1o]
  const CoshTab * ctab_a;    // [N][Mpad] cosh 2W
  const CoshTab * ctab_b;    // [N][Mpad] sinh 2W
  const cd * w2;             // [N][Mpad] 2W (exact)
  int8_t * spins;
  cd * theta;
  cd * lnpsi0;
  unsigned char * fresh;
  const int * order;
  int pos0, nsweeps;
  const double * uniforms;
  unsigned long long seed, step0;
  long long chain_offset;
  unsigned char * acc_log;
};

inline size_t ffnn_sweep_smem_bytes(int N, int warps, int Mpad)
{
  const size_t npad = (size_t)((N+15)/16)*16;
  size_t b = (size_t)warps*Mpad*(sizeof(cd)+sizeof(double));          // tanh theta, Im log cosh theta
  b += (size_t)Mpad*sizeof(cd);                                      // w1o
  b += (size_t)NQS_FF_STAGES*3*Mpad*sizeof(cd);                      // table rows (cosh 2W, sinh 2W, 2W) of the proposals in flight
  b += (size_t)warps*npad+(size_t)N*sizeof(int);                     // spins, site order
  return (b+15)/16*16+2*NQS_FF_STAGES*sizeof(uint64_t)+16;
}

// log f for f = 1 + z.  The flip factors sit close to 1 (|2W| is small for any trained or freshly initialised network), where
//   log(1 + z) = 2 atanh(u),  u = z / (2 + z),  atanh(u) = u (1 + u^2/3 + u^4/5 + ...)
// converges fast: for |z|^2 < 0.09 (|u| < 0.177) ten terms leave < 1e-17, and the whole complex log is ~70 branch-free fp64
// instructions -- against ~200 with branches for log + atan2 (fp64 atan2 alone is ~150), which is what made the first version
// of these kernels no faster than a fresh log cosh.  The principal branch is the series' own (|Im| < pi/2 for |z| < 1).
// Returns false when |z|^2 >= 0.09: the caller takes log / atan2 then.
__device__ __forceinline__ bool c_log_near_one(const double fx, const double fy, double & lr, double & li)
{
  const double zx = fx-1.0, zy = fy;
  const double dx = 2.0+zx;
  const double inv = 1.0/fma(dx, dx, zy*zy);
  // u = z conj(2 + z) / |2 + z|^2
  const double ux = fma(zx, dx, zy*zy)*inv, uy = (zy*dx-zx*zy)*inv;
  const double wx = fma(ux, ux, -uy*uy), wy = 2.0*ux*uy;        // u^2
  double sx = 1.0/19.0, sy = 0.0;
#pragma unroll
  for (int n = 17; n >= 1; n -= 2)
  {
    const double tx = fma(sx, wx, fma(-sy, wy, 1.0/(double)n)), ty = fma(sx, wy, sy*wx);
    sx = tx; sy = ty;
  }
  lr = 2.0*fma(ux, sx, -uy*sy); li = 2.0*fma(ux, sy, uy*sx);
  return fma(zx, zx, zy*zy) < 0.09;
}
// (Keeping the series branch-free and redoing a whole proposal on the rare out-of-range factor was tried: the compiler then
// interleaves all 16 factors of a lane, spills ~650 bytes and the sweep gets 20 % slower.  The per-factor branch stays.)
__device__ __forceinline__ void c_log_factor(const double fx, const double fy, double & lr, double & li)
{
  if (!c_log_near_one(fx, fy, lr, li))
  {
    lr = 0.5*log(fma(fx, fx, fy*fy));
    li = atan2(fy, fx);
  }
}

// principal value of a + d for a in [-pi, pi], d in [-pi, pi]
__device__ __forceinline__ double wrap_pi(const double v)
{
  return (v > NQS_FF_PI) ? v-2.0*NQS_FF_PI : ((v < -NQS_FF_PI) ? v+2.0*NQS_FF_PI : v);
}

// One warp per chain; lane l owns hidden units l, l+32, ... (JPL = Mpad/32 of them).  The exact theta of those units stays in
// REGISTERS for the whole launch and takes the reference's rank-1 update at every accepted flip (same operands, same order:
// bit-identical to the generic kernel); tanh theta and Im log cosh theta sit in shared memory, rebuilt from theta at every
// sweep.  The three table rows of a proposal's site arrive by TMA through a ring shared by the warps of the CTA.
template <int JPL>
__global__ void __launch_bounds__(32*NQS_FF_MAX_WARPS) ffnn_sweep_fast_kernel(const FfnnSweepArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  const int N = a.N, M = a.M;
  constexpr int Mpad = 32*JPL;
  const int npad = ((N+15)/16)*16;
  cd * Tall = reinterpret_cast<cd*>(smem_raw);                                   // [warps][Mpad]
  cd * wsh = Tall+(size_t)warps*Mpad;                                            // [Mpad]
  cd * stage0 = wsh+Mpad;                                                        // [STAGES][3][Mpad]
  double * Aall = reinterpret_cast<double*>(stage0+(size_t)NQS_FF_STAGES*3*Mpad); // [warps][Mpad]
  int * ord = reinterpret_cast<int*>(Aall+(size_t)warps*Mpad);                   // [N]
  int8_t * spall = reinterpret_cast<int8_t*>(ord+N);                             // [warps][npad]
  uint64_t * full = reinterpret_cast<uint64_t*>(smem_raw+(((size_t)(reinterpret_cast<unsigned char*>(spall)+(size_t)warps*npad-smem_raw)+15)/16)*16);
  uint64_t * empty = full+NQS_FF_STAGES;
  cd * T = Tall+(size_t)w*Mpad;
  double * A = Aall+(size_t)w*Mpad;
  int8_t * sp = spall+(size_t)w*npad;
  const cd * w1o = a.params+(size_t)N*M+M;
  for (int i = threadIdx.x; i < N; i += blockDim.x) ord[i] = a.order[i];
  for (int j = threadIdx.x; j < Mpad; j += blockDim.x) wsh[j] = (j < M) ? w1o[j] : cmake(0.0, 0.0);
  const long long kblock = (long long)blockIdx.x*warps;
  if (threadIdx.x == 0)
  {
    long long nact = a.K-kblock;
    if (nact > warps) nact = warps;
    if (nact < 1) nact = 1;
    for (int q = 0; q < NQS_FF_STAGES; ++q) { mbar_init(full+q, 1); mbar_init(empty+q, (uint32_t)nact); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long k = kblock+w;
  if (k >= a.K) return;
  for (int i = lane; i < N; i += 32) sp[i] = a.spins[k*N+i];
  cd th[JPL];
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj)
  {
    const int j = lane+32*jj;
    th[jj] = (j < M) ? a.theta[k*M+j] : cmake(0.0, 0.0);
  }
  cd ln0 = a.lnpsi0[k];
  bool any_acc = false;
  __syncwarp();
  int pos = a.pos0;
  long long t_glob = 0;
  const long long t_end = (long long)a.nsweeps*N;
  const uint32_t row_bytes = (uint32_t)(Mpad*sizeof(cd));
  auto issue_rows = [&](const long long q)
  {
    const int sq = ord[(int)(((long long)a.pos0+q)%N)];
    const int slot = (int)(q%NQS_FF_STAGES);
    cd * dst = stage0+(size_t)slot*3*Mpad;
    mbar_expect_tx(full+slot, 3*row_bytes);
    tma_load_1d(dst, a.ctab_a+(size_t)sq*Mpad, row_bytes, full+slot);
    tma_load_1d(dst+Mpad, a.ctab_b+(size_t)sq*Mpad, row_bytes, full+slot);
    tma_load_1d(dst+2*Mpad, a.w2+(size_t)sq*Mpad, row_bytes, full+slot);
  };
  if (w == 0 && lane == 0)
    for (long long q = 0; q < NQS_FF_STAGES-1 && q < t_end; ++q) issue_rows(q);
  double ubuf = 0.0;
  cd lncur = cmake(0.0, 0.0);

  for (int sweep = 0; sweep < a.nsweeps; ++sweep)
  {
    // ---- (1) resident state from the exact theta: tanh, Im log cosh, and ln psi(theta) = sum_j w_j log cosh theta_j
    lncur = cmake(0.0, 0.0);
#pragma unroll
    for (int jj = 0; jj < JPL; ++jj)
    {
      const int j = lane+32*jj;
      cd t = cmake(0.0, 0.0);
      double im = 0.0;
      if (j < M)
      {
        t = c_tanh(th[jj]);
        const cd lc = c_logcosh(th[jj]);
        im = lc.y;
        lncur = cadd(lncur, cmul(wsh[j], lc));
      }
      T[j] = t; A[j] = im;
    }
    lncur = warp_sum(lncur);
    __syncwarp();
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
      const int slot = (int)(t_glob%NQS_FF_STAGES);
      if (w == 0 && lane == 0)
      {
        const long long q = t_glob+NQS_FF_STAGES-1;
        if (q < t_end)
        {
          if (t_glob > 0) mbar_wait(empty+(int)(q%NQS_FF_STAGES), (uint32_t)(((t_glob-1)/NQS_FF_STAGES)&1));
          issue_rows(q);
        }
      }
      mbar_wait(full+slot, (uint32_t)((t_glob/NQS_FF_STAGES)&1));
      const cd * ca = stage0+(size_t)slot*3*Mpad+lane;
      const cd * cb = ca+Mpad;
      const cd * cw = ca+2*Mpad;
      const double sg = (double)sp[site];
      cd dsum = cmake(0.0, 0.0);
      double Anew[JPL];                          // Im log cosh of the proposed state: adopted as it is on an accept
#pragma unroll
      for (int jj = 0; jj < JPL; ++jj)
      {
        const int j = lane+32*jj;
        const cd tj = T[j], c2 = ca[32*jj], s2 = cb[32*jj], wj = wsh[j];
        // f = cosh 2W - sigma tanh(theta) sinh 2W
        const double gx = fma(-tj.y, s2.y, tj.x*s2.x), gy = fma(tj.y, s2.x, tj.x*s2.y);
        const double fx = fma(-sg, gx, c2.x), fy = fma(-sg, gy, c2.y);
        double lr, la;
        c_log_factor(fx, fy, lr, la);
        const double aj = A[j];
        Anew[jj] = wrap_pi(aj+la);
        const double li = Anew[jj]-aj;
        dsum.x += wj.x*lr-wj.y*li; dsum.y += wj.x*li+wj.y*lr;
      }
      dsum = warp_sum(dsum);
      const cd ln1 = cadd(lncur, dsum);
      // ref k6 Sampler__ParallelMetropolisUpdate__ (impl_mcmc_sampler.cuh:75-102)
      const double d = ln1.x-ln0.x;
      const double ratio = exp(2.0*((d < 0) ? 1.0 : 0.0)*d);
      const bool acc = (u < ratio);
      const double delta = acc ? 1.0 : 0.0;
      ln0 = cmake(ln0.x+delta*(ln1.x-ln0.x), ln0.y+delta*(ln1.y-ln0.y));
      if (a.acc_log && lane == 0) a.acc_log[t_glob*a.K+k] = acc ? 1 : 0;
      if (acc)
      {
        lncur = ln1;
        any_acc = true;
#pragma unroll
        for (int jj = 0; jj < JPL; ++jj)
        {
          const int j = lane+32*jj;
          const cd tj = T[j], c2 = ca[32*jj], s2 = cb[32*jj], wv = cw[32*jj];
          // ref k7 conditional_y_update (:1314-1329): theta -= W * (2 sigma); the table row holds 2W (exact), so (2W) * sigma has
          // the reference's bits
          th[jj].x -= wv.x*sg; th[jj].y -= wv.y*sg;
          const double gx = fma(-tj.y, s2.y, tj.x*s2.x), gy = fma(tj.y, s2.x, tj.x*s2.y);
          const double fx = fma(-sg, gx, c2.x), fy = fma(-sg, gy, c2.y);
          A[j] = Anew[jj];
          // tanh(theta - 2 sigma W) = (tanh cosh2W - sigma sinh2W) / (cosh2W - sigma tanh sinh2W) = num / f
          const double nx = fma(-tj.y, c2.y, tj.x*c2.x)-sg*s2.x, ny = fma(tj.y, c2.x, tj.x*c2.y)-sg*s2.y;
          const double inv = 1.0/fma(fx, fx, fy*fy);
          T[j] = cmake((nx*fx+ny*fy)*inv, (ny*fx-nx*fy)*inv);
        }
        if (lane == 0) sp[site] = (int8_t)(-sp[site]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty+slot);
    }
  }
  // ---- (3) write back; chains that accepted get ln psi0 from the final exact theta (the tracked value is the same up to the
  // rounding of the accumulated differences)
  cd lsum = cmake(0.0, 0.0);
#pragma unroll
  for (int jj = 0; jj < JPL; ++jj)
  {
    const int j = lane+32*jj;
    if (j < M)
    {
      a.theta[k*M+j] = th[jj];
      if (any_acc) lsum = cadd(lsum, cmul(wsh[j], c_logcosh(th[jj])));
    }
  }
  if (any_acc) ln0 = warp_sum(lsum);
  for (int i = lane; i < N; i += 32) a.spins[k*N+i] = sp[i];
  if (lane == 0)
  {
    a.lnpsi0[k] = ln0;
    if (any_acc) a.fresh[k] = 1;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Local energy, FNN: lanes run over SITES (as rbm_eloc_sites_kernel): lane l owns site i = 32*block + l and walks the hidden
// units serially, accumulating sum_j w_j log f_j; tanh(theta_j), Im log cosh theta_j of the CTA's C chains and w sit in shared
// memory (broadcast reads), the tables are read TRANSPOSED ([j][i], i contiguous: coalesced, shared by the C chains).
// ---------------------------------------------------------------------------------------------------------------------
struct FfnnElocArgs
{
  int N, M, Npad;
  long long K;
  const cd * params;
  const CoshTab * ctabT_a;   // [M][Npad] cosh 2W
  const CoshTab * ctabT_b;   // [M][Npad] sinh 2W
  const int8_t * spins;
  const cd * theta;
  const cd * lnpsi0;
  const double * Jmat;
  const double * sjs;        // [K] sum_ij s_i J_ij s_j (tensor-core GEMM) or null
  double hfield;
  cd * htilda;
};

inline size_t ffnn_eloc_smem_bytes(int N, int M, int C)
{
  const size_t npad = (size_t)((N+15)/16)*16;
  return (size_t)C*M*(sizeof(cd)+sizeof(double))+(size_t)M*sizeof(cd)+(size_t)C*npad+(size_t)C*8*6*sizeof(double)+32;
}

template <int C>
__global__ void __launch_bounds__(256) ffnn_eloc_fast_kernel(const FfnnElocArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, M = a.M, Npad = a.Npad;
  const int npad = ((N+15)/16)*16;
  const int nwarps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  cd * Tsh = reinterpret_cast<cd*>(smem_raw);                       // [C][M]
  cd * wsh = Tsh+(size_t)C*M;                                       // [M]
  double * Ash = reinterpret_cast<double*>(wsh+M);                  // [C][M]
  double * red = Ash+(size_t)C*M;                                   // [C][8][6]
  int8_t * sp = reinterpret_cast<int8_t*>(red+(size_t)C*8*6);       // [C][npad]
  const cd * w1o = a.params+(size_t)N*M+M;
  const long long kbase = (long long)blockIdx.x*C;
  double ls_x[C], ls_y[C];
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    const long long k = (kbase+c < a.K) ? kbase+c : kbase;
    ls_x[c] = 0.0; ls_y[c] = 0.0;
    for (int j = threadIdx.x; j < M; j += blockDim.x)
    {
      const cd th = a.theta[k*M+j];
      const cd lc = c_logcosh(th), wj = w1o[j];
      Tsh[c*M+j] = c_tanh(th);
      Ash[c*M+j] = lc.y;
      ls_x[c] += wj.x*lc.x-wj.y*lc.y; ls_y[c] += wj.x*lc.y+wj.y*lc.x;      // ln psi(theta)
      if (c == 0) wsh[j] = wj;
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) sp[c*npad+i] = a.spins[k*N+i];
  }
  __syncthreads();
  double part_d[C], part_x[C], part_y[C];
#pragma unroll
  for (int c = 0; c < C; ++c) { part_d[c] = 0.0; part_x[c] = 0.0; part_y[c] = 0.0; }
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
  // sum_j w_j log f_ij for this lane's site, all chains of the CTA; kept as (re, im) of ln psi' - ln psi(theta)
  for (int sb = w; sb*32 < N; sb += nwarps)
  {
    const int i = sb*32+lane;
    const bool ok = (i < N);
    double sg[C], dx[C], dy[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { sg[c] = (ok && sp[c*npad+i] < 0) ? -1.0 : 1.0; dx[c] = 0.0; dy[c] = 0.0; }
    const CoshTab * ca = a.ctabT_a+sb*32+lane;
    const CoshTab * cb = a.ctabT_b+sb*32+lane;
    for (int j = 0; j < M; ++j)
    {
      const double2 c2 = ld_tab(ca+(size_t)j*Npad), s2 = ld_tab(cb+(size_t)j*Npad);
      const cd wj = wsh[j];
#pragma unroll
      for (int c = 0; c < C; ++c)
      {
        const cd tj = Tsh[c*M+j];
        const double aj = Ash[c*M+j];
        const double gx = fma(-tj.y, s2.y, tj.x*s2.x), gy = fma(tj.y, s2.x, tj.x*s2.y);
        const double fx = fma(-sg[c], gx, c2.x), fy = fma(-sg[c], gy, c2.y);
        double lr, la;
        c_log_factor(fx, fy, lr, la);
        const double li = wrap_pi(aj+la)-aj;
        dx[c] += wj.x*lr-wj.y*li; dy[c] += wj.x*li+wj.y*lr;
      }
    }
    if (ok)
    {
#pragma unroll
      for (int c = 0; c < C; ++c) { const cd e = c_exp(cmake(dx[c], dy[c])); part_x[c] += e.x; part_y[c] += e.y; }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c)
  {
    const double d = warp_sum(part_d[c]), x = warp_sum(part_x[c]), y = warp_sum(part_y[c]);
    const double lx = warp_sum(ls_x[c]), ly = warp_sum(ls_y[c]);
    if (lane == 0)
    {
      double * r = red+((size_t)c*8+w)*6;
      r[0] = d; r[1] = x; r[2] = y; r[3] = lx; r[4] = ly;
    }
  }
  __syncthreads();
  if (threadIdx.x < C && kbase+threadIdx.x < a.K)
  {
    const int c = threadIdx.x;
    double d = 0, x = 0, y = 0, lx = 0, ly = 0;
    for (int ww = 0; ww < nwarps; ++ww)
    {
      const double * r = red+((size_t)c*8+ww)*6;
      d += r[0]; x += r[1]; y += r[2]; lx += r[3]; ly += r[4];
    }
    if (a.sjs != nullptr) d = a.sjs[kbase+c];
    // sum_i exp(ln psi'_i - ln psi0) = exp(ln psi(theta) - ln psi0) sum_i exp(ln psi'_i - ln psi(theta)); the first factor is 1 up to
    // rounding when the tracked ln psi0 is current, and carries the staleness after warm_up's quirk flip or a parameter update
    const cd l0 = a.lnpsi0[kbase+c];
    const cd off = cmul(cmake(x, y), c_exp(cmake(lx-l0.x, ly-l0.y)));
    a.htilda[kbase+c] = cmake((0.5*d+a.hfield*off.x)/(double)N, (a.hfield*off.y)/(double)N);
  }
}
} // namespace nqs
