// One launch per CG iteration for everything that is not the pass over O.
//
// ref: ConjugateGradient::solve (gpu/include/conjugate_gradient.cuh:29-74) + the tails of SMatrixForCG::dot / applyPrecond
// (gpu/include/functor_for_CG.cuh:115-135).  The reference spends 7 small kernels, 3 thrust reductions and 4 host syncs per
// iteration there.  Here ONE kernel of co-resident CTAs (<= 1 per SM) walks the iteration with
// two software grid barriers; every CTA re-derives the scalars from the same per-CTA partial sums in the same fixed order,
// so no broadcast is needed and the result is deterministic.  Scalars stay on the device (CgScalars); the host only polls
// `done` every few iterations, without draining the queue.
//
//   MODE_ITER  t = S p ;  alpha = rho / Re<t,p> ;  x += alpha p ;  r -= alpha t ;  |r|^2 < thr -> done
//              z = M^-1 r ;  rho' = Re<z,r> ;  beta = rho'/rho ;  p = z + beta p ;  <O>.p updated by linearity
//   MODE_INIT  t = S x0 ;  r = F - t ;  |F|^2 == 0 -> x = 0, done ;  |r|^2 < thr -> done ;  p = M^-1 r ;  rho = Re<p,r> ;  <O>.p
//   MODE_DOT   t = S v  (nqs_smatrix_dot)
// with  (S v)_p = traw_p / K - conj(<O>_p) (<O>.v) + lambda diag_p v_p,   traw = sum_k conj(O_kp) (O_k . v)  (all ranks),
//       M = (1 + lambda) diag(S).
// traw comes straight from the cluster / row-block partials of the S*v kernel (folded here in fixed order: saves a launch).
//
// Multi-GPU: the all-reduce of the P complex partial sums that every CG iteration needs is done INSIDE this kernel over
// NVLink peer memory instead of a separate NCCL launch: every rank stores its folded partial into slot [rank] of a receive
// buffer that lives on EVERY peer (plain st.global on cudaIpc-mapped peer pointers), publishes per-(parity, rank, CTA) epoch
// flags with a system-scope release, waits for the flags of all ranks and adds the slots in RANK ORDER -- so the sum is
// bit-identical on all ranks (the replicated CG state never diverges) and costs one NVLink hop (~P*16 B per peer) instead
// of a collective launch.  Receive buffers are double-buffered by epoch parity: a rank can only write epoch n+2 after it
// has seen every peer's flag n+1, which a peer raises after it finished reading epoch n.  (If the peer mapping is not
// available the engine falls back to colsum_reduce_kernel + ncclAllReduce + traw.)
#pragma once
#include "device_math.cuh"
#include "sr_kernels.cuh"

namespace nqs
{
enum { CG_MODE_ITER = 0, CG_MODE_INIT = 1, CG_MODE_DOT = 2 };

#define NQS_CG_THREADS 256
#define NQS_CG_MAX_CTAS 148   // one CTA per SM at most: the software grid barrier needs every CTA resident
#define NQS_CG_NVALS 6
#define NQS_CG_MAX_RANKS 16

struct CgArgs
{
  long long P;
  int mode;
  int nparts;              // > 0: traw = sum over parts of part[q][{re,im}][P]; 0: read traw
  const double * part;
  const double * traw;     // [2][P]
  double inv_ktot, lambda;
  const cd * aO;
  const double * diag;
  const cd * F;            // INIT
  cd * v;                  // ITER: p (updated in place); INIT: x0 (read; zeroed when |F| = 0); DOT: v
  cd * pvec;               // INIT: p out
  cd * x;                  // ITER
  cd * r;                  // ITER (in/out), INIT (out)
  cd * t;                  // scratch / DOT result
  CgScalars * sc;
  double * slots;          // [2][NQS_CG_MAX_CTAS][NQS_CG_NVALS] (successive grid sums alternate between the two halves)
  unsigned int * barrier;  // zero between launches
  // in-kernel all-reduce over peer memory (n_ranks > 1 and the peers are mapped)
  int n_ranks, rank;
  unsigned int epoch;      // increases by one per exchange on every rank
  double * peer_x[NQS_CG_MAX_RANKS];            // peer_x[r]: receive buffer of rank r, [2][n_ranks][2P]
  unsigned int * peer_flag[NQS_CG_MAX_RANKS];   // peer_flag[r]: flags of rank r, [2][NQS_CG_MAX_RANKS][NQS_CG_MAX_CTAS]
  unsigned long long * trace;                   // NQS_CG_TRACE=1: globaltimer stamps of CTA 0, [NQS_CG_TRACE_WORDS] per launch (else null)
  const double * hsums;    // INIT (may be null): all-reduced (sum Re h, ...) -- a non-finite energy ends the solve before it starts
};
#define NQS_CG_TRACE_WORDS 24   // entry, stored, released, arrival of every rank's flag [16], waited, end
__device__ __forceinline__ unsigned long long cg_now()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// All CTAs of the grid are co-resident (grid <= #SMs; the launch is cooperative, so the driver guarantees it): spin barrier on
// a global counter.  Where a cooperative launch is not available the wait is bounded: after NQS_CG_BARRIER_TIMEOUT_NS the CTA
// raises *timeout_flag and moves on (the results of the launch are garbage then, and the host reports the error instead of hanging).
#define NQS_CG_BARRIER_TIMEOUT_NS 20000000000ull
__device__ __forceinline__ void cg_grid_barrier(unsigned int * counter, const unsigned int target, int * timeout_flag)
{
  __syncthreads();
  if (threadIdx.x == 0)
  { // release (cumulative over the CTA's writes ordered by the barrier above) / acquire at gpu scope: no full fences needed
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(counter) : "memory");
    unsigned int seen;
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    for (;;)
    {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen >= target) break;
      if ((++spins&1023u) == 0u)
      { // look at the clock only now and then: the common wait is a few hundred polls
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now-t0 > NQS_CG_BARRIER_TIMEOUT_NS) { *timeout_flag = 1; break; }
      }
    }
  }
  __syncthreads();
}

// CTA partials -> slots; after the grid barrier every CTA folds all slots in the same order.  vals/out: NV doubles.
template <int NV>
__device__ __forceinline__ void cg_grid_sum(double (&vals)[NV], const CgArgs & a, double * sh, unsigned int & epoch)
{
  const int lane = threadIdx.x&31, w = threadIdx.x>>5;
  // a CTA can only write the slots of sum k+2 after every CTA finished reading those of sum k (it passed barrier k+1)
  double * slots = a.slots+(size_t)(epoch&1u)*NQS_CG_MAX_CTAS*NQS_CG_NVALS;
#pragma unroll
  for (int i = 0; i < NV; ++i) vals[i] = warp_sum(vals[i]);
  if (lane == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[w*NV+i] = vals[i];
  }
  __syncthreads();
  if (threadIdx.x < NV)
  {
    double s = 0.0;
    for (int ww = 0; ww < NQS_CG_THREADS/32; ++ww) s += sh[ww*NV+threadIdx.x];
    slots[(size_t)blockIdx.x*NQS_CG_NVALS+threadIdx.x] = s;
  }
  ++epoch;
  cg_grid_barrier(a.barrier, epoch*gridDim.x, &a.sc->barrier_timeout);
  // every warp of every CTA folds the per-CTA partials in the same order: lane-strided, then a fixed butterfly
  double s[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) s[i] = 0.0;
  for (int b = lane; b < (int)gridDim.x; b += 32)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i) s[i] += __ldcg(slots+(size_t)b*NQS_CG_NVALS+i);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) vals[i] = warp_sum(s[i]);
}

// sum_q part[q][{re,im}][p] in a fixed order; the loads of up to 16 partials are issued before the first add (a serial chain
// would expose one L2 round trip per partial: ~9 us for 15 cluster partials)
__device__ __forceinline__ void cg_fold_parts(const double * __restrict__ part, const int nparts, const long long P, const long long p,
  double & rx, double & ry)
{
  const double * base = part+p;
  rx = 0.0; ry = 0.0;
  for (int q0 = 0; q0 < nparts; q0 += 16)
  {
    double vx[16], vy[16];
#pragma unroll
    for (int u = 0; u < 16; ++u)
    {
      const bool ok = (q0+u < nparts);
      vx[u] = ok ? __ldcg(base+(size_t)(q0+u)*2*P) : 0.0;
      vy[u] = ok ? __ldcg(base+(size_t)(q0+u)*2*P+P) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) { rx += vx[u]; ry += vy[u]; }
  }
}

// MODE_ITER with every vector element a thread owns held in REGISTERS across the two grid sums (EPT = ceil(P / threads of the
// grid) <= 4): one round of global loads, issued before the wait for the peers, and one round of stores at the end -- the
// phases between the barriers touch no memory but the reduction slots.  Same arithmetic in the same order as the generic
// loop below (bit-identical results).
template <int EPT>
__device__ __forceinline__ void cg_iter_regs(const CgArgs & a, double * sh, unsigned int & epoch)
{
  const long long P = a.P;
  const long long i0 = (long long)blockIdx.x*blockDim.x+threadIdx.x, stride = (long long)gridDim.x*blockDim.x;
  const double pre = 1.0+a.lambda;
  const bool tracing = (a.trace != nullptr && blockIdx.x == 0);
  const double aovx = a.sc->aov_x, aovy = a.sc->aov_y, rho = a.sc->rho, thr = a.sc->thr;
  const int fixed = a.sc->fixed;
  bool ok[EPT];
  long long pp[EPT];
  double trx[EPT], try_[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e)
  {
    ok[e] = (i0+e*stride < P);
    pp[e] = ok[e] ? i0+e*stride : 0;
    trx[e] = 0.0; try_[e] = 0.0;
    if (ok[e])
    {
      if (a.nparts > 0) cg_fold_parts(a.part, a.nparts, P, pp[e], trx[e], try_[e]);
      else { trx[e] = a.traw[pp[e]]; try_[e] = a.traw[P+pp[e]]; }
    }
  }
  const bool p2p = (a.n_ranks > 1 && a.nparts > 0);
  const int par = (int)(a.epoch&1u);
  if (p2p)
  {
    const size_t slot = ((size_t)par*a.n_ranks+a.rank)*2*(size_t)P;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
      if (ok[e])
        for (int r = 0; r < a.n_ranks; ++r) { a.peer_x[r][slot+pp[e]] = trx[e]; a.peer_x[r][slot+P+pp[e]] = try_[e]; }
  }
  // the vectors do not depend on the exchange: their loads travel while the peers' partials do
  cd ao[EPT], pv[EPT], xv[EPT], rv[EPT];
  double dg[EPT];
#pragma unroll
  for (int e = 0; e < EPT; ++e)
  {
    ao[e] = cmake(0.0, 0.0); pv[e] = ao[e]; xv[e] = ao[e]; rv[e] = ao[e]; dg[e] = 1.0;
    if (ok[e]) { ao[e] = a.aO[pp[e]]; pv[e] = a.v[pp[e]]; xv[e] = a.x[pp[e]]; rv[e] = a.r[pp[e]]; dg[e] = a.diag[pp[e]]; }
  }
  if (p2p)
  {
    __syncthreads();
    if (tracing && threadIdx.x == 0) a.trace[1] = cg_now();
    if (threadIdx.x < a.n_ranks)
      asm volatile("st.release.sys.global.u32 [%0], %1;"
        :: "l"(a.peer_flag[threadIdx.x]+((size_t)par*NQS_CG_MAX_RANKS+a.rank)*NQS_CG_MAX_CTAS+blockIdx.x), "r"(a.epoch) : "memory");
    if (tracing && threadIdx.x == 0) a.trace[2] = cg_now();
    if (threadIdx.x < a.n_ranks)
    {
      const unsigned int * f = a.peer_flag[a.rank]+((size_t)par*NQS_CG_MAX_RANKS+threadIdx.x)*NQS_CG_MAX_CTAS+blockIdx.x;
      unsigned int seen;
      const unsigned long long t0 = cg_now();
      for (;;)
      {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
        if ((int)(seen-a.epoch) >= 0) break;
        if (cg_now()-t0 > 20000000000ull) { a.sc->peer_timeout = 1; break; }
      }
      if (tracing) a.trace[3+threadIdx.x] = cg_now();
    }
    __syncthreads();
    if (tracing && threadIdx.x == 0) a.trace[19] = cg_now();
    const double * xin = a.peer_x[a.rank]+(size_t)par*a.n_ranks*2*(size_t)P;
#pragma unroll
    for (int e = 0; e < EPT; ++e)
    {
      trx[e] = 0.0; try_[e] = 0.0;
      if (ok[e])
        for (int r = 0; r < a.n_ranks; ++r) { trx[e] += __ldcv(xin+(size_t)r*2*P+pp[e]); try_[e] += __ldcv(xin+(size_t)r*2*P+P+pp[e]); }
    }
  }
  // ---- t = S p and Re<t, p>
  cd tv[EPT];
  double s1[1] = {0.0};
#pragma unroll
  for (int e = 0; e < EPT; ++e)
  {
    const double cx = ao[e].x*aovx+ao[e].y*aovy, cy = ao[e].x*aovy-ao[e].y*aovx;
    tv[e] = cmake(trx[e]*a.inv_ktot-cx, try_[e]*a.inv_ktot-cy);
    tv[e].x += a.lambda*dg[e]*pv[e].x; tv[e].y += a.lambda*dg[e]*pv[e].y;
    if (ok[e]) s1[0] += tv[e].x*pv[e].x+tv[e].y*pv[e].y;
  }
  cg_grid_sum<1>(s1, a, sh, epoch);
  const double alpha = rho/s1[0];
  cd zv[EPT];
  double s2[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int e = 0; e < EPT; ++e)
  {
    xv[e].x += alpha*pv[e].x; xv[e].y += alpha*pv[e].y;
    rv[e].x -= alpha*tv[e].x; rv[e].y -= alpha*tv[e].y;
    const double den = pre*dg[e];
    zv[e] = cmake(rv[e].x/den, rv[e].y/den);
    if (ok[e])
    {
      s2[0] += cnorm(rv[e]);
      s2[1] += zv[e].x*rv[e].x+zv[e].y*rv[e].y;
      s2[2] += ao[e].x*zv[e].x-ao[e].y*zv[e].y; s2[3] += ao[e].x*zv[e].y+ao[e].y*zv[e].x;
    }
  }
  cg_grid_sum<4>(s2, a, sh, epoch);
  const double beta = s2[1]/rho;
  const bool done = (!fixed && s2[0] < thr);
#pragma unroll
  for (int e = 0; e < EPT; ++e)
  {
    if (!ok[e]) continue;
    a.x[pp[e]] = xv[e]; a.r[pp[e]] = rv[e];
    if (!done) a.v[pp[e]] = cmake(zv[e].x+beta*pv[e].x, zv[e].y+beta*pv[e].y);        // conjugate_gradient.cuh:71
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    CgScalars * sc = a.sc;
    sc->tp = s1[0]; sc->alpha = alpha; sc->res2 = s2[0]; sc->iters += 1; sc->rho_old = rho; sc->rho = s2[1]; sc->beta = beta;
    sc->aov_x = s2[2]+beta*aovx; sc->aov_y = s2[3]+beta*aovy;   // <O>.(z + beta p) by linearity
    if (done) sc->done = 1;
  }
}

__global__ void __launch_bounds__(NQS_CG_THREADS, 1) cg_fused_kernel(const CgArgs a)
{
  if (a.mode == CG_MODE_ITER && a.sc->done) return;     // uniform over the grid: converged earlier
  if (a.mode == CG_MODE_INIT && a.hsums != nullptr && !isfinite(a.hsums[0]))
  { // ref optimizer.cuh:134-138: <h> not finite -> no solve, no update (uniform over the grid and over the ranks)
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.sc->nonfinite = 1; a.sc->done = 1; a.sc->iters = 0; }
    return;
  }
  __shared__ double sh[(NQS_CG_THREADS/32)*NQS_CG_NVALS];
  const long long P = a.P;
  const long long i0 = (long long)blockIdx.x*blockDim.x+threadIdx.x, stride = (long long)gridDim.x*blockDim.x;
  unsigned int epoch = 0;
  const double pre = 1.0+a.lambda;
  const bool tracing = (a.trace != nullptr && blockIdx.x == 0);
  if (tracing && threadIdx.x == 0) a.trace[0] = cg_now();
  const long long ept = (P+stride-1)/stride;
  if (a.mode == CG_MODE_ITER && ept <= 4)
  {
    if (ept <= 1) cg_iter_regs<1>(a, sh, epoch);
    else if (ept <= 2) cg_iter_regs<2>(a, sh, epoch);
    else cg_iter_regs<4>(a, sh, epoch);
  }
  else
  {
  // <O>.v: carried by recurrence during the iteration, computed explicitly for x0 / an arbitrary v
  double aovx, aovy;
  if (a.mode == CG_MODE_ITER) { aovx = a.sc->aov_x; aovy = a.sc->aov_y; }
  else
  {
    double s[2] = {0.0, 0.0};
    for (long long p = i0; p < P; p += stride)
    {
      const cd ao = a.aO[p], vv = a.v[p];
      s[0] += ao.x*vv.x-ao.y*vv.y; s[1] += ao.x*vv.y+ao.y*vv.x;
    }
    cg_grid_sum<2>(s, a, sh, epoch);
    aovx = s[0]; aovy = s[1];
  }

  // ---- multi-GPU: push this rank's folded partial to every peer, then wait for everybody's
  const bool p2p = (a.n_ranks > 1 && a.nparts > 0);
  const double * xin = nullptr;
  if (p2p)
  {
    const int par = (int)(a.epoch&1u);
    const size_t slot = ((size_t)par*a.n_ranks+a.rank)*2*(size_t)P;
    for (long long p = i0; p < P; p += stride)
    {
      double trx, try_;
      cg_fold_parts(a.part, a.nparts, P, p, trx, try_);
      for (int r = 0; r < a.n_ranks; ++r) { a.peer_x[r][slot+p] = trx; a.peer_x[r][slot+P+p] = try_; }
    }
    // CTA b of every rank owns the same elements p (equal grids), so the hand-shake is per CTA: no grid barrier, and a slow
    // CTA only delays its own counterparts.  The release store at system scope publishes every store the CTA made before
    // the barrier (cumulativity); flags are [parity][source rank][CTA].
    __syncthreads();
    if (tracing && threadIdx.x == 0) a.trace[1] = cg_now();
    if (threadIdx.x < a.n_ranks)
      asm volatile("st.release.sys.global.u32 [%0], %1;"
        :: "l"(a.peer_flag[threadIdx.x]+((size_t)par*NQS_CG_MAX_RANKS+a.rank)*NQS_CG_MAX_CTAS+blockIdx.x), "r"(a.epoch) : "memory");
    if (tracing && threadIdx.x == 0) a.trace[2] = cg_now();
    if (threadIdx.x < a.n_ranks)
    {
      const unsigned int * f = a.peer_flag[a.rank]+((size_t)par*NQS_CG_MAX_RANKS+threadIdx.x)*NQS_CG_MAX_CTAS+blockIdx.x;
      unsigned int seen;
      unsigned long long t0 = 0, now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
      for (;;)
      {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
        if ((int)(seen-a.epoch) >= 0) break;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (now-t0 > 20000000000ull) { a.sc->peer_timeout = 1; break; }   // a peer died: never hang the GPU (host raises NQS_ERR_NCCL)
      }
      if (tracing) a.trace[3+threadIdx.x] = cg_now();
    }
    __syncthreads();
    if (tracing && threadIdx.x == 0) a.trace[19] = cg_now();
    xin = a.peer_x[a.rank]+(size_t)par*a.n_ranks*2*(size_t)P;
  }

  // ---- t = S v, and the sums that decide the step
  double s1[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
  for (long long p = i0; p < P; p += stride)
  {
    double trx = 0.0, try_ = 0.0;
    if (p2p)
    { // rank order: identical bits on every rank
      for (int r = 0; r < a.n_ranks; ++r) { trx += __ldcv(xin+(size_t)r*2*P+p); try_ += __ldcv(xin+(size_t)r*2*P+P+p); }
    }
    else if (a.nparts > 0) cg_fold_parts(a.part, a.nparts, P, p, trx, try_);
    else { trx = a.traw[p]; try_ = a.traw[P+p]; }
    const cd ao = a.aO[p], vv = a.v[p];
    const double dg = a.diag[p];
    // conj(aO) * aov
    const double cx = ao.x*aovx+ao.y*aovy, cy = ao.x*aovy-ao.y*aovx;
    cd tv = cmake(trx*a.inv_ktot-cx, try_*a.inv_ktot-cy);
    tv.x += a.lambda*dg*vv.x; tv.y += a.lambda*dg*vv.y;
    if (a.mode == CG_MODE_INIT)
    {
      const cd f = a.F[p];
      const cd rv = csub(f, tv);
      a.r[p] = rv;
      const double den = pre*dg;
      const cd pv = cmake(rv.x/den, rv.y/den);
      a.pvec[p] = pv;
      s1[0] += cnorm(f); s1[1] += cnorm(rv);
      s1[2] += pv.x*rv.x+pv.y*rv.y;                            // Re(p conj(r))
      s1[3] += ao.x*pv.x-ao.y*pv.y; s1[4] += ao.x*pv.y+ao.y*pv.x;
    }
    else
    {
      a.t[p] = tv;
      s1[0] += tv.x*vv.x+tv.y*vv.y;                            // Re(t conj(p))
    }
  }
  if (a.mode != CG_MODE_DOT) cg_grid_sum<5>(s1, a, sh, epoch);

  if (a.mode == CG_MODE_DOT) {}
  else if (a.mode == CG_MODE_INIT)
  {
    const double rhs2 = s1[0], res2 = s1[1];
    const bool zero_rhs = (rhs2 == 0.0);
    const double thr = fmax(a.sc->tol2*rhs2, 2.2250738585072014e-308);     // std::numeric_limits<double>::min()
    if (zero_rhs)
      for (long long p = i0; p < P; p += stride) a.v[p] = cmake(0.0, 0.0);   // conjugate_gradient.cuh:39-43: x = 0
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
      CgScalars * sc = a.sc;
      sc->rhs2 = rhs2; sc->res2 = res2; sc->rho = s1[2]; sc->aov_x = s1[3]; sc->aov_y = s1[4];
      sc->zero_rhs = zero_rhs ? 1 : 0; sc->thr = thr; sc->iters = 0;
      sc->done = (zero_rhs || (!sc->fixed && res2 < thr)) ? 1 : 0;
    }
  }
  else
  {
    const double rho = a.sc->rho, thr = a.sc->thr;
    const int fixed = a.sc->fixed;
    const double alpha = rho/s1[0];
    double s2[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long p = i0; p < P; p += stride)
    {
      const cd pv = a.v[p], tv = a.t[p], ao = a.aO[p];
      cd xv = a.x[p], rv = a.r[p];
      xv.x += alpha*pv.x; xv.y += alpha*pv.y;
      rv.x -= alpha*tv.x; rv.y -= alpha*tv.y;
      a.x[p] = xv; a.r[p] = rv;
      const double den = pre*a.diag[p];
      const cd zv = cmake(rv.x/den, rv.y/den);
      a.t[p] = zv;                                             // t is dead now: reuse it for z
      s2[0] += cnorm(rv);
      s2[1] += zv.x*rv.x+zv.y*rv.y;
      s2[2] += ao.x*zv.x-ao.y*zv.y; s2[3] += ao.x*zv.y+ao.y*zv.x;
    }
    cg_grid_sum<4>(s2, a, sh, epoch);
    const double beta = s2[1]/rho;
    const bool done = (!fixed && s2[0] < thr);
    if (!done)
      for (long long p = i0; p < P; p += stride)
      {
        const cd zv = a.t[p], pv = a.v[p];
        a.v[p] = cmake(zv.x+beta*pv.x, zv.y+beta*pv.y);        // conjugate_gradient.cuh:71
      }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
      CgScalars * sc = a.sc;
      sc->tp = s1[0]; sc->alpha = alpha; sc->res2 = s2[0]; sc->iters += 1; sc->rho_old = rho; sc->rho = s2[1]; sc->beta = beta;
      sc->aov_x = s2[2]+beta*aovx; sc->aov_y = s2[3]+beta*aovy;   // <O>.(z + beta p) by linearity
      if (done) sc->done = 1;
    }
  }
  }
  if (tracing && threadIdx.x == 0) a.trace[20] = cg_now();
  // leave the barrier counter at zero for the next launch: the last CTA to get here resets it
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const unsigned int n = atomicAdd(a.barrier, 1u);
    if (n == (epoch+1)*gridDim.x-1) *a.barrier = 0u;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// SR setup across ranks without a collective launch: the all-reduce of the 5P+3 local sums (sum O, sum O conj(h), sum |O|^2,
// sum h, sum |h|^2) and setup_finalize_kernel in ONE kernel over the NVLink peer mapping of the CG exchange.  Every rank stores
// its sums into slot [rank] of a receive buffer on every peer, raises per-(parity, rank, CTA) epoch flags (system-scope
// release), waits for the flags of CTA b of every peer and adds the slots in RANK ORDER -- bit-identical <O>, F, diag S on all
// ranks, as ncclAllReduce gives, at one NVLink hop instead of a ~75 us collective (8 GPUs, 1.3 MB).  Grid <= #SMs (all CTAs
// resident: a CTA only ever waits for its counterparts on the peers).  The all-reduced h sums go to hsall[3] for the CG kernel's
// finite-energy check and the host's statistics.
// ---------------------------------------------------------------------------------------------------------------------
struct SetupXArgs
{
  long long P;
  double inv_ktot;
  const double * sums;      // [5P+3] local sums
  double * hsall;           // [3] out: all-reduced (sum Re h, sum Im h, sum |h|^2)
  cd * aO;
  cd * F;                   // may be null
  double * diag;
  int n_ranks, rank;
  unsigned int epoch;
  double * peer_x[NQS_CG_MAX_RANKS];            // receive buffers [2][n_ranks][5P+4]
  unsigned int * peer_flag[NQS_CG_MAX_RANKS];   // flags [2][NQS_CG_MAX_RANKS][NQS_CG_MAX_CTAS]
  int * timeout_flag;
};

__global__ void __launch_bounds__(256) setup_exchange_finalize_kernel(const SetupXArgs a)
{
  const long long P = a.P, stride5 = 5*P+4;
  const int par = (int)(a.epoch&1u);
  const long long i0 = (long long)blockIdx.x*blockDim.x+threadIdx.x, gs = (long long)gridDim.x*blockDim.x;
  const size_t myslot = ((size_t)par*a.n_ranks+a.rank)*(size_t)stride5;
  for (long long p = i0; p < P; p += gs)
  {
    double v[5];
#pragma unroll
    for (int c = 0; c < 5; ++c) v[c] = a.sums[c*P+p];
    for (int r = 0; r < a.n_ranks; ++r)
    {
      double * dst = a.peer_x[r]+myslot;
#pragma unroll
      for (int c = 0; c < 5; ++c) dst[c*P+p] = v[c];
    }
  }
  if (threadIdx.x < 3)
  { // every CTA needs <h>: each CTA publishes the three h sums under its own flag (its counterparts read them from their own copy)
    const double hv = a.sums[5*P+threadIdx.x];
    for (int r = 0; r < a.n_ranks; ++r) a.peer_x[r][myslot+5*P+threadIdx.x] = hv;
  }
  __syncthreads();
  if (threadIdx.x < a.n_ranks)
  {
    asm volatile("st.release.sys.global.u32 [%0], %1;"
      :: "l"(a.peer_flag[threadIdx.x]+((size_t)par*NQS_CG_MAX_RANKS+a.rank)*NQS_CG_MAX_CTAS+blockIdx.x), "r"(a.epoch) : "memory");
    const unsigned int * f = a.peer_flag[a.rank]+((size_t)par*NQS_CG_MAX_RANKS+threadIdx.x)*NQS_CG_MAX_CTAS+blockIdx.x;
    unsigned int seen;
    const unsigned long long t0 = cg_now();
    for (;;)
    {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(f) : "memory");
      if ((int)(seen-a.epoch) >= 0) break;
      if (cg_now()-t0 > NQS_CG_BARRIER_TIMEOUT_NS) { *a.timeout_flag = 1; break; }
    }
  }
  __syncthreads();
  const double * xin = a.peer_x[a.rank]+(size_t)par*a.n_ranks*(size_t)stride5;
  double hs[3] = {0.0, 0.0, 0.0};
  for (int r = 0; r < a.n_ranks; ++r)
  {
#pragma unroll
    for (int c = 0; c < 3; ++c) hs[c] += __ldcv(xin+(size_t)r*stride5+5*P+c);
  }
  // NOTE: the h sums of rank r were published by EVERY CTA of rank r with identical values; this CTA has seen the flag of its
  // counterpart, whose copy is complete
  const cd conj_havg = cmake(hs[0]*a.inv_ktot, -hs[1]*a.inv_ktot);
  for (long long p = i0; p < P; p += gs)
  {
    double v[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int r = 0; r < a.n_ranks; ++r)
    {
#pragma unroll
      for (int c = 0; c < 5; ++c) v[c] += __ldcv(xin+(size_t)r*stride5+c*P+p);
    }
    const cd ao = cmake(v[0]*a.inv_ktot, v[1]*a.inv_ktot);
    a.aO[p] = ao;
    if (a.F)
    {
      const cd fr = cmake(v[2]*a.inv_ktot, v[3]*a.inv_ktot);
      a.F[p] = cconj(csub(fr, cmul(conj_havg, ao)));
    }
    a.diag[p] = v[4]*a.inv_ktot-cnorm(ao);
  }
  if (blockIdx.x == 0 && threadIdx.x < 3) a.hsall[threadIdx.x] = hs[threadIdx.x];
}
} // namespace nqs
