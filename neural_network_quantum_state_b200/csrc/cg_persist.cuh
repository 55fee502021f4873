// The whole preconditioned CG solve of one SR step as ONE persistent kernel launch.
//
// ref: ConjugateGradient::solve (gpu/include/conjugate_gradient.cuh:29-74) driving SMatrixForCG::dot / applyPrecond
// (gpu/include/functor_for_CG.cuh:104-135): per iteration 2 passes over O (Zgemm + Zgemv), 7 small kernels, 3 thrust
// reductions and 4 host synchronisations.  cg_fused.cuh + sv_fused.cuh brought that down to two launches per iteration
// (one pass over O, one vector kernel) with a host poll every few iterations; what was left -- two launch gaps, the prologue
// of the cluster kernel (shared-memory fill, barrier init, cluster sync, cold TMA pipeline) and the vector kernel's own
// launch -- is a FIXED cost of ~26 us per iteration on one GPU and ~46 us on eight, which does not shrink when the chains
// are sharded and was the strong-scaling limiter (8 GPUs: 0.78 efficiency).
//
// Here the clusters of sv_fused_kernel stay resident for the whole solve.  Per product (product 0 is S x0, then one per
// iteration):
//   rows    the row loop of sv_fused.cuh, unchanged in structure: a producer warp streams this CTA's column slice of the
//           cluster's rows by TMA, the consumer warps form O_k . v, exchange the CTA partials over DSMEM and accumulate
//           conj(O_kp) z_k in registers.  The mbarrier phases simply keep counting across products, and the producer
//           PREFETCHES the first rows of the next product while the vector phase runs (O does not change during a solve),
//           so the HBM stream never drains between products.
//   A       cluster partials -> part[cluster][2][P]; software grid barrier.
//   vector  every consumer thread of the grid owns <= EPT vector elements: folds the cluster partials (fixed order),
//           multi-GPU: pushes them to every peer over NVLink, raises / awaits the per-CTA epoch flags and adds the ranks'
//           slots in rank order (cg_fused.cuh's exchange, now without a launch of its own); then the iteration body of
//           cg_fused.cuh with x, r, p, z held in registers across its two grid sums.  Same arithmetic per element as
//           cg_fused_kernel; the grid sums fold the elements in a different (equally fixed) order, so the iterates agree
//           with the launch-per-iteration path to rounding and are run-to-run deterministic (tests).
//   (Merging the two grid sums of an iteration into one 11-value sum -- alpha, the predicted rho' and |r'|^2 from scalar products
//           known before alpha -- was measured: ~15 us against 7.5 + 5 us for the two small sums.  Not kept.)
//   (no barrier B) the next direction d = z + beta d is NOT waited for: z is stored before the second grid sum, beta is known to
//           every thread after it, and the previous direction was stored one product earlier -- so every CTA forms its slice
//           of the new direction itself at the top of the next row pass (same fma as the owner: same bits).
// Three grid barriers (~3 us each) replace two launches, and the host neither polls nor synchronises during the solve: it
// reads the scalars back with the step's final read-back.  All CTAs take the same decisions from the same numbers (every
// grid sum is folded in the same order by every warp), so `converged` is uniform without a broadcast.
#pragma once
#include "sv_fused.cuh"
#include "cg_fused.cuh"

namespace nqs
{
struct CgpArgs
{
  // ---- the pass over O (see SvArgs)
  long long K, P;
  const cd * O;
  double * part;            // [n_clusters][2][P]
  long long pc, rows_per_cluster;
  int nslot;
  unsigned int slot_bytes;
  int depth;
  // ---- the solve
  double inv_ktot, lambda, tol2;
  int fixed_iters;          // > 0: exactly this many iterations (no convergence test)
  int max_iter;
  const cd * aO;
  const double * diag;
  const cd * F;
  cd * x;                   // in: warm start, out: solution
  cd * r;
  cd * pb[2];               // direction d_k of product k lives in pb[k & 1] (written by its owner thread during product k-1)
  cd * zv;                  // z_k = M^-1 r_k, written before the second grid sum of product k
  CgScalars * sc;
  double * slots;           // [2][NQS_CGP_MAX_CTAS][NQS_CG_NVALS]
  unsigned int * barrier;   // zero between launches
  const double * hsums;     // sums + 5P: (sum Re h, sum Im h, sum |h|^2) after the all-reduce -- non-finite energy skips the solve
  // ---- in-kernel all-reduce over peer memory (n_ranks > 1)
  int n_ranks, rank;
  unsigned int epoch0;      // exchange epochs used by this launch: epoch0 + 1, epoch0 + 2, ... (one per product)
  // Reduce-scatter + all-gather in "LL" packets (NCCL's low-latency protocol): every double travels as a 16-byte packet
  // {lo32, epoch, hi32, epoch}, so the data carries its own arrival flag -- 8-byte stores are atomic, the receiver polls the
  // packet itself, and no release fence / flag round trip follows the stores.
  //   phase 1: rank r sends the elements of slice q (P/n_ranks of them) of its folded partial to rank q        -> ll1[q][r][slice]
  //   phase 2: rank q adds the n_ranks copies of its slice in RANK ORDER and sends the sums to every rank     -> ll2[*][P]
  // (the earlier all-to-all push of whole vectors moved n_ranks times the bytes: 3.7 MB per rank and product at 8 GPUs, ~12 us
  // with its system-scope release, trace of round 2)
  uint4 * peer_ll1[NQS_CG_MAX_RANKS];     // [n_ranks][S][2] packets, S = ceil(P / n_ranks)
  uint4 * peer_ll2[NQS_CG_MAX_RANKS];     // [P][2] packets
  unsigned long long * trace;   // NQS_CG_TRACE=1: [NQS_CGP_TRACE_WORDS] globaltimer stamps of CTA 0 per product (else null)
  int trace_max;
};
#define NQS_CGP_NVALS 8          // values per grid sum (at most 5 are used)
#define NQS_CGP_MAX_CTAS 2048    // small problems run several narrow CTAs per SM; slots / exchange flags are sized for this many
#define NQS_CGP_TRACE_WORDS 8   // product start, rows done, barrier A passed, pushed + flags raised, peers seen, sum 1, sum 2, end
// shared memory after sv_fused's core tail (part of NQS_SV_TAIL_BYTES): reduction scratch [NQS_SV_MAX_WARPS][NQS_CG_NVALS] doubles | control words
static_assert(NQS_SV_TAIL_BYTES-NQS_SV_TAIL_CORE >= NQS_SV_MAX_WARPS*NQS_CGP_NVALS*8+64+64, "NQS_SV_TAIL_BYTES reserves the scratch");
template <int CPT> struct CgpEpt { static const int value = (CPT <= 3) ? 1 : 2; };

// mbarrier primitives on precomputed shared-memory addresses
__device__ __forceinline__ void mbar_wait_a(const uint32_t addr, const uint32_t parity)
{
  uint32_t ok;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_a(const uint32_t addr)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(const uint32_t addr, const uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(addr), "r"(bytes) : "memory");
}

// ---- LL packets
__device__ __forceinline__ void ll_store(uint4 * dst, const double v, const unsigned int flag)
{
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};"
    :: "l"(dst), "r"((unsigned int)__double2loint(v)), "r"(flag), "r"((unsigned int)__double2hiint(v)), "r"(flag) : "memory");
}
// waits until both halves of the packet carry `flag`; false after NQS_CG_BARRIER_TIMEOUT_NS (a peer died)
__device__ __forceinline__ bool ll_wait(const uint4 * src, const unsigned int flag, double & v)
{
  unsigned int lo, f1, hi, f2, spins = 0;
  unsigned long long t0 = 0;
  for (;;)
  {
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(lo), "=r"(f1), "=r"(hi), "=r"(f2) : "l"(src) : "memory");
    if (f1 == flag && f2 == flag) break;
    if ((++spins&4095u) == 0u)
    {
      const unsigned long long now = cg_now();
      if (t0 == 0) t0 = now;
      else if (now-t0 > NQS_CG_BARRIER_TIMEOUT_NS) { v = 0.0; return false; }
    }
  }
  v = __hiloint2double((int)hi, (int)lo);
  return true;
}

// CTA-level barrier of the consumer warps only (the producer warp never joins): named barrier 1
__device__ __forceinline__ void cgp_cta_sync(const int NT) { asm volatile("bar.sync 1, %0;" :: "r"(NT) : "memory"); }

// Software grid barrier over the consumer threads of all CTAs; false = gave up (see cg_grid_barrier).  One fence on either
// side of a RELAXED arrive / poll: an acquire load per poll would invalidate the L1 every time round the loop (CCTL.IVALL),
// and a release reduction would fence per arrive -- the form cooperative_groups' grid.sync() uses.
__device__ __forceinline__ bool cgp_grid_barrier(const CgpArgs & a, const unsigned int target, const int NT, volatile int * ctl)
{
  cgp_cta_sync(NT);
  if (threadIdx.x == 0)
  {
    __threadfence();          // release: everything the CTA wrote before the barrier above (cumulativity)
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" :: "l"(a.barrier) : "memory");
    unsigned int seen, spins = 0;
    unsigned long long t0 = 0;
    for (;;)
    {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.barrier) : "memory");
      if (seen >= target) break;
      if ((++spins&1023u) == 0u)
      {
        int gave_up;
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(gave_up) : "l"(&a.sc->barrier_timeout) : "memory");
        const unsigned long long now = cg_now();
        if (t0 == 0) t0 = now;
        if (gave_up || now-t0 > NQS_CG_BARRIER_TIMEOUT_NS) { a.sc->barrier_timeout = 1; ctl[2] = 1; break; }
      }
    }
    __threadfence();          // acquire
  }
  cgp_cta_sync(NT);
  return ctl[2] == 0;
}

// grid sum of NV values per thread over the consumer threads of the grid (cg_grid_sum's fold order: warps of a CTA in warp
// order, then the CTAs lane-strided + butterfly, identical in every warp)
template <int NV>
__device__ __forceinline__ bool cgp_grid_sum(double (&vals)[NV], const CgpArgs & a, double * sh, unsigned int & epoch, unsigned int & nsum,
  const int NT, volatile int * ctl)
{
  const int lane = threadIdx.x&31, w = threadIdx.x>>5, NW = NT>>5;
  double * slots = a.slots+(size_t)(nsum&1u)*NQS_CGP_MAX_CTAS*NQS_CGP_NVALS;
#pragma unroll
  for (int i = 0; i < NV; ++i) vals[i] = warp_sum(vals[i]);
  if (lane == 0)
  {
#pragma unroll
    for (int i = 0; i < NV; ++i) sh[w*NV+i] = vals[i];
  }
  cgp_cta_sync(NT);
  if (threadIdx.x < NV)
  {
    double s = 0.0;
    for (int ww = 0; ww < NW; ++ww) s += sh[ww*NV+threadIdx.x];
    slots[(size_t)blockIdx.x*NQS_CGP_NVALS+threadIdx.x] = s;
  }
  ++epoch; ++nsum;
  const bool ok = cgp_grid_barrier(a, epoch*gridDim.x, NT, ctl);
  // ONE warp per CTA folds the per-CTA partials -- lane-strided with up to 8 loads per value in flight (a serial loop costs an
  // L2 round trip per 32 CTAs), then a fixed butterfly -- and hands the totals to the other warps through shared memory: the
  // same order in every CTA, and 1/NW of the L2 reads of a fold by every warp
  if (w == 0)
  {
    double s[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) s[i] = 0.0;
    for (int b0 = 0; b0 < (int)gridDim.x; b0 += 256)
    {
      double v[8][NV];
#pragma unroll
      for (int q = 0; q < 8; ++q)
      {
        const int b = b0+32*q+lane;
#pragma unroll
        for (int i = 0; i < NV; ++i) v[q][i] = (b < (int)gridDim.x) ? __ldcg(slots+(size_t)b*NQS_CGP_NVALS+i) : 0.0;
      }
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int i = 0; i < NV; ++i) s[i] += v[q][i];
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) s[i] = warp_sum(s[i]);
    if (lane == 0)
    {
#pragma unroll
      for (int i = 0; i < NV; ++i) sh[i] = s[i];
    }
  }
  cgp_cta_sync(NT);
#pragma unroll
  for (int i = 0; i < NV; ++i) vals[i] = sh[i];
  cgp_cta_sync(NT);      // sh is rewritten by the next grid sum
  return ok;
}

// Everything of one product that is not the pass over O: barrier A, fold / exchange of the partials, the CG recurrence with its
// two grid sums, and (product 0 only) barrier B.  Returns 1 when the solve is over (converged, iteration limit, or a barrier /
// peer gave up), else 0.
struct CgpCtx
{
  double * sh;              // reduction scratch
  double * fsc;             // recurrence scalars
  volatile int * ctl;       // control words
  unsigned int epoch, nsum; // grid barriers passed / grid sums done
  int NT, CS;
  bool tracing;
};
template <int EPT>
__device__ __noinline__ int cgp_vector_phase(const CgpArgs & a, const int prod, CgpCtx & cx)
{
  double * const sh = cx.sh;
  double * const fsc = cx.fsc;
  volatile int * const ctl = cx.ctl;
  unsigned int & epoch = cx.epoch, & nsum = cx.nsum;
  const int NT = cx.NT, tid = threadIdx.x;
  const unsigned int CS = (unsigned int)cx.CS;
  const bool tracing = cx.tracing;
  const long long P = a.P;
  const long long gtid = (long long)blockIdx.x*NT+tid, gstride = (long long)gridDim.x*NT;
  const double pre = 1.0+a.lambda;
  const bool p2p = (a.n_ranks > 1);
  bool alive = true;
  {
      // ================================================================ A: every cluster's partials are visible
      ++epoch;
      if (!cgp_grid_barrier(a, epoch*gridDim.x, NT, ctl)) return 1;
      if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+2] = cg_now();

      // ================================================================ vector phase
      bool ok[EPT];
      long long pp[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e) { ok[e] = (gtid+e*gstride < P); pp[e] = ok[e] ? gtid+e*gstride : 0; }
      // traw_p = sum over clusters (fixed order) [and ranks, in rank order]
      double trx[EPT], try_[EPT];
      const int ncl = (int)(gridDim.x/CS);
#pragma unroll
      for (int e = 0; e < EPT; ++e)
      {
        trx[e] = 0.0; try_[e] = 0.0;
        if (ok[e]) cg_fold_parts(a.part, ncl, P, pp[e], trx[e], try_[e]);
      }
      const unsigned int xepoch = a.epoch0+1u+(unsigned int)prod;      // LL flag of this product: never 0, never repeated
      bool peer_ok = true;
      const long long S = (P+a.n_ranks-1)/a.n_ranks;          // elements per slice
      if (p2p)
      { // phase 1: each element of the folded partial goes to the rank that owns its slice
#pragma unroll
        for (int e = 0; e < EPT; ++e)
          if (ok[e])
          {
            const int q = (int)(pp[e]/S);
            uint4 * dst = a.peer_ll1[q]+(((size_t)a.rank*S+(size_t)(pp[e]-q*S))<<1);
            ll_store(dst, trx[e], xepoch); ll_store(dst+1, try_[e], xepoch);
          }
      }
      // the vectors do not depend on the exchange: their loads travel while the peers' partials do
      cd ao[EPT], pv[EPT], xv[EPT], rv[EPT];
      double dg[EPT];
#pragma unroll
      for (int e = 0; e < EPT; ++e)
      {
        ao[e] = cmake(0.0, 0.0); pv[e] = ao[e]; xv[e] = ao[e]; rv[e] = ao[e]; dg[e] = 1.0;
        if (ok[e])
        {
          ao[e] = a.aO[pp[e]]; dg[e] = a.diag[pp[e]]; xv[e] = __ldcg(a.x+pp[e]);
          if (prod > 0) { pv[e] = __ldcg(a.pb[prod&1]+pp[e]); rv[e] = __ldcg(a.r+pp[e]); }
          else { pv[e] = xv[e]; rv[e] = a.F[pp[e]]; }     // product 0: v = x0, and r starts from F
        }
      }
      if (p2p)
      {
        if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+3] = cg_now();
        // phase 1 receive / phase 2 send: the elements of THIS rank's slice, spread over all CTAs (thread t of CTA b takes slice
        // element t*gridDim + b, so every SM injects its share into NVLink)
        long long s_mine = P-(long long)a.rank*S;
        if (s_mine > S) s_mine = S;
        for (long long i = (long long)tid*gridDim.x+blockIdx.x; i < s_mine; i += (long long)NT*gridDim.x)
        {
          double fx = 0.0, fy = 0.0;
          for (int r = 0; r < a.n_ranks; ++r)
          { // rank order: the same bits on every rank
            const uint4 * src = a.peer_ll1[a.rank]+(((size_t)r*S+(size_t)i)<<1);
            double vx, vy;
            peer_ok = ll_wait(src, xepoch, vx) && peer_ok;
            peer_ok = ll_wait(src+1, xepoch, vy) && peer_ok;
            fx += vx; fy += vy;
          }
          const size_t pe = ((size_t)a.rank*S+(size_t)i)<<1;
          for (int r = 0; r < a.n_ranks; ++r) { ll_store(a.peer_ll2[r]+pe, fx, xepoch); ll_store(a.peer_ll2[r]+pe+1, fy, xepoch); }
        }
        // phase 2 receive: this thread's own elements, summed over all ranks
#pragma unroll
        for (int e = 0; e < EPT; ++e)
        {
          trx[e] = 0.0; try_[e] = 0.0;
          if (ok[e])
          {
            const uint4 * src = a.peer_ll2[a.rank]+((size_t)pp[e]<<1);
            peer_ok = ll_wait(src, xepoch, trx[e]) && peer_ok;
            peer_ok = ll_wait(src+1, xepoch, try_[e]) && peer_ok;
          }
        }
        if (!peer_ok) { a.sc->peer_timeout = 1; a.sc->barrier_timeout = 1; ctl[2] = 1; }
        if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+4] = cg_now();
      }

      if (prod == 0)
      { // ---- MODE_INIT of cg_fused.cuh: <O>.x0, t = S x0, r = F - t, p = M^-1 r
        double s0[2] = {0.0, 0.0};
#pragma unroll
        for (int e = 0; e < EPT; ++e)
          if (ok[e]) { s0[0] += ao[e].x*xv[e].x-ao[e].y*xv[e].y; s0[1] += ao[e].x*xv[e].y+ao[e].y*xv[e].x; }
        if (!cgp_grid_sum<2>(s0, a, sh, epoch, nsum, NT, ctl)) return 1;
        double s1[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int e = 0; e < EPT; ++e)
        {
          const double cx = ao[e].x*s0[0]+ao[e].y*s0[1], cy = ao[e].x*s0[1]-ao[e].y*s0[0];
          cd tv = cmake(trx[e]*a.inv_ktot-cx, try_[e]*a.inv_ktot-cy);
          tv.x += a.lambda*dg[e]*xv[e].x; tv.y += a.lambda*dg[e]*xv[e].y;
          const cd f = rv[e];
          rv[e] = csub(f, tv);
          const double den = pre*dg[e];
          pv[e] = cmake(rv[e].x/den, rv[e].y/den);
          if (ok[e])
          {
            s1[0] += cnorm(f); s1[1] += cnorm(rv[e]);
            s1[2] += pv[e].x*rv[e].x+pv[e].y*rv[e].y;
            s1[3] += ao[e].x*pv[e].x-ao[e].y*pv[e].y; s1[4] += ao[e].x*pv[e].y+ao[e].y*pv[e].x;
          }
        }
        if (!cgp_grid_sum<5>(s1, a, sh, epoch, nsum, NT, ctl)) return 1;
        const double rhs2 = s1[0], res2 = s1[1];
        const bool zero_rhs = (rhs2 == 0.0);
        const double thr = fmax(a.tol2*rhs2, 2.2250738585072014e-308);
        const double rho = s1[2], aovx = s1[3], aovy = s1[4];
        const bool done = zero_rhs || (a.fixed_iters <= 0 && res2 < thr);
        if (tid == 0) { fsc[0] = rho; fsc[1] = thr; fsc[2] = aovx; fsc[3] = aovy; fsc[4] = 0.0; ctl[3] = 0; }
#pragma unroll
        for (int e = 0; e < EPT; ++e)
        {
          if (!ok[e]) continue;
          a.r[pp[e]] = rv[e]; a.pb[1][pp[e]] = pv[e];                          // d_1
          if (zero_rhs) a.x[pp[e]] = cmake(0.0, 0.0);                          // conjugate_gradient.cuh:39-43
        }
        if (blockIdx.x == 0 && tid == 0)
        {
          CgScalars * sc = a.sc;
          sc->rhs2 = rhs2; sc->res2 = res2; sc->rho = rho; sc->aov_x = aovx; sc->aov_y = aovy;
          sc->zero_rhs = zero_rhs ? 1 : 0; sc->thr = thr; sc->iters = 0; sc->done = done ? 1 : 0;
        }
        if (done) alive = false;
      }
      else
      { // ---- MODE_ITER of cg_fused.cuh (cg_iter_regs), same arithmetic in the same order
        const double rho = fsc[0], thr = fsc[1], aovx = fsc[2], aovy = fsc[3];
        const int iters = ctl[3]+1;
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
        if (!cgp_grid_sum<1>(s1, a, sh, epoch, nsum, NT, ctl)) return 1;
        if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+5] = cg_now();
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
            a.zv[pp[e]] = zv[e];             // visible to every CTA behind the barrier of the grid sum below
            s2[0] += cnorm(rv[e]);
            s2[1] += zv[e].x*rv[e].x+zv[e].y*rv[e].y;
            s2[2] += ao[e].x*zv[e].x-ao[e].y*zv[e].y; s2[3] += ao[e].x*zv[e].y+ao[e].y*zv[e].x;
          }
        }
        if (!cgp_grid_sum<4>(s2, a, sh, epoch, nsum, NT, ctl)) return 1;
        if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+6] = cg_now();
        const double beta = s2[1]/rho;
        const bool conv = (a.fixed_iters <= 0 && s2[0] < thr);
        const bool done = conv || (a.fixed_iters > 0 ? iters >= a.fixed_iters : iters >= a.max_iter);
#pragma unroll
        for (int e = 0; e < EPT; ++e)
        {
          if (!ok[e]) continue;
          a.x[pp[e]] = xv[e]; a.r[pp[e]] = rv[e];
          // d_{k+1} = z_k + beta_k d_k (conjugate_gradient.cuh:71) for this thread's own use in product k+1 and for the row passes of k+2
          if (!conv) a.pb[(prod+1)&1][pp[e]] = cmake(fma(beta, pv[e].x, zv[e].x), fma(beta, pv[e].y, zv[e].y));
        }
        if (blockIdx.x == 0 && tid == 0)
        {
          CgScalars * sc = a.sc;
          sc->tp = s1[0]; sc->alpha = alpha; sc->res2 = s2[0]; sc->iters = iters; sc->rho_old = rho; sc->rho = s2[1]; sc->beta = beta;
          sc->aov_x = s2[2]+beta*aovx; sc->aov_y = s2[3]+beta*aovy;
          if (conv) sc->done = 1;
        }
        // every thread has read the old scalars before the two grid sums above (CTA barriers inside), so thread 0 may overwrite them
        if (tid == 0) { fsc[0] = s2[1]; fsc[2] = s2[2]+beta*aovx; fsc[3] = s2[3]+beta*aovy; fsc[4] = beta; ctl[3] = iters; }   // <O>.(z + beta p) by linearity
        if (done) alive = false;
      }
      if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+7] = cg_now();
      if (!alive) return 1;
      if (prod == 0)
      { // ============================================================== B (once per solve): d_1 is visible to every CTA
        ++epoch;
        if (!cgp_grid_barrier(a, epoch*gridDim.x, NT, ctl)) return 1;
      }
      else cgp_cta_sync(NT);   // beta of this product (shared memory, thread 0) before the next row pass reads it
  }
  return 0;
}

// blockDim.x = 32*(NW+1) as in sv_fused_kernel; grid = clusters x cluster size, every CTA resident (cooperative launch)
template <int CPT, int DEFER>
__global__ void __maxnreg__(SvMaxRegs<CPT>::value) cg_persist_kernel(const CgpArgs a)
{
  constexpr int EPT = CgpEpt<CPT>::value;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int CS = cluster.num_blocks(), crank = cluster.block_rank();
  const int NT = blockDim.x-32, tid = threadIdx.x, lane = tid&31, w = tid>>5, NW = NT>>5;
  const long long cid = blockIdx.x/CS;
  unsigned char * tail = smem_raw+(size_t)a.nslot*a.slot_bytes;
  cd * red = reinterpret_cast<cd*>(tail);
  cd * zbuf = red+NQS_SV_RBUFS*NQS_SV_MAX_WARPS;
  uint64_t * full = reinterpret_cast<uint64_t*>(zbuf+NQS_SV_ZBUFS*NQS_SV_MAX_CLUSTER);
  uint64_t * empty = full+NQS_SV_MAX_SLOTS;
  uint64_t * wfull = empty+NQS_SV_MAX_SLOTS;
  uint64_t * zfull = wfull+NQS_SV_RBUFS;
  double * sh = reinterpret_cast<double*>(tail+NQS_SV_TAIL_CORE);                    // [NW][NQS_CGP_NVALS] reduction scratch
  volatile int * ctl = reinterpret_cast<volatile int*>(sh+NQS_SV_MAX_WARPS*NQS_CGP_NVALS+8); // [0] stop, [1] rows consumed, [2] barrier gave up, [3] iterations done

  const long long P = a.P;
  const long long c0 = (long long)crank*a.pc;
  long long nr_ll = P-c0;
  if (nr_ll > a.pc) nr_ll = a.pc;
  if (nr_ll < 0) nr_ll = 0;
  const int n_r = (int)nr_ll;
  const long long k0 = cid*a.rows_per_cluster;
  long long k1 = k0+a.rows_per_cluster;
  if (k1 > a.K) k1 = a.K;
  const int nrows = (k1 > k0) ? (int)(k1-k0) : 0;
  const uint32_t row_bytes = (uint32_t)n_r*(uint32_t)sizeof(cd);
  const int n_prod_max = (a.fixed_iters > 0 ? a.fixed_iters : a.max_iter)+1;

  { // zero the slots once (TMA only writes the first row_bytes of a slot: the padding reads as 0, no bounds predicate in the loop)
    double2 * z = reinterpret_cast<double2*>(smem_raw);
    const int nz16 = (int)(((size_t)a.nslot*a.slot_bytes)/sizeof(double2));
    for (int i = tid; i < nz16; i += blockDim.x) z[i] = make_double2(0.0, 0.0);
  }
  if (tid == 0)
  {
    for (int s = 0; s < a.nslot; ++s) { mbar_init(full+s, 1); mbar_init(empty+s, (uint32_t)NW); }
    for (int q = 0; q < NQS_SV_RBUFS; ++q) mbar_init(wfull+q, (uint32_t)NW);
    for (int q = 0; q < NQS_SV_ZBUFS; ++q) mbar_init(zfull+q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    ctl[0] = 0; ctl[1] = 0; ctl[2] = 0; ctl[3] = 0;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  cluster.sync();

  // <h> not finite (ref optimizer.cuh:134-138: the reference stops there): nothing is solved, x keeps its value and the update
  // kernel skips itself on the same flag.  Uniform over the grid (all ranks hold the same all-reduced sums).
  const bool finite_h = isfinite(a.hsums[0]);

  if (w == NW)
  { // ---- producer warp: one lane keeps NSLOT rows of this CTA's column slice in flight, across product boundaries
    if (lane == 0 && n_r > 0 && finite_h)
    {
      const cd * Oslice = a.O+c0;
      long long g = 0;                 // rows issued so far
      int slot = 0;
      uint32_t par = 0;
      bool stopped = false;
      for (int prod = 0; prod < n_prod_max && !stopped; ++prod)
        for (int it = 0; it < nrows; ++it)
        {
          if (g >= a.nslot)
          { // consumers released the row that used this slot before -- or the solve is over
            const uint32_t addr = smem_u32(empty+slot);
            uint32_t ok = 0;
            for (;;)
            {
              asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(ok) : "r"(addr), "r"(par^1u), "r"(2000u) : "memory");
              if (ok) break;
              if (ctl[0]) { stopped = true; break; }
            }
            if (stopped) break;
          }
          mbar_expect_tx(full+slot, row_bytes);
          tma_load_1d(smem_raw+(size_t)slot*a.slot_bytes, Oslice+(k0+it)*P, row_bytes, full+slot);
          ++g;
          if (++slot == a.nslot) { slot = 0; par ^= 1u; }
        }
      while (ctl[0] == 0) __nanosleep(200);
      // rows that were prefetched for a product that never ran: let them land before the CTA may exit
      for (long long q = (long long)ctl[1]; q < g; ++q)
        mbar_wait(full+(int)(q%a.nslot), (uint32_t)((q/a.nslot)&1));
    }
  }
  else
  {
    unsigned int epoch = 0, nsum = 0;           // grid barriers passed / grid sums done (slot parity)
    const long long gtid = (long long)blockIdx.x*NT+tid, gstride = (long long)gridDim.x*NT;
    const bool tracing = (a.trace != nullptr && blockIdx.x == 0 && tid == 0);
    const double pre = 1.0+a.lambda;
    const bool p2p = (a.n_ranks > 1);
    // scalars of the recurrence live in shared memory between products (identical in every thread; thread 0 stores them), so the
    // row loop -- which is at the register limit -- carries nothing of the vector phase
    double * fsc = sh+NQS_SV_MAX_WARPS*NQS_CGP_NVALS;           // [0] rho [1] thr [2] aov_x [3] aov_y [4] beta [5] alpha
    const cd * const sbase = reinterpret_cast<const cd*>(smem_raw)+tid;
    const size_t slot_elems = a.slot_bytes/sizeof(cd);
    const int depth = a.depth;
    // Ring positions / mbarrier phase parities keep counting across products: the TMA slots and the reducer warp roll, the
    // power-of-two rings (warp partials, z buffers) are bit fields of the running row counters
    int slot = 0, tail_slot = 0, redw = 0, gi = 0, gt = 0;   // gi / gt: rows that went through pass (2) / pass (3) so far
    uint32_t full_par = 0;
    const uint32_t full_a = smem_u32(full), empty_a = smem_u32(empty), wfull_a = smem_u32(wfull), zfull_a = smem_u32(zfull);
    // where this CTA's partial lands in CTA `lane` of the cluster (shared::cluster addresses are linear inside a CTA's window)
    const uint32_t rz_data = (lane < (int)CS) ? map_to_rank(zbuf+crank, (uint32_t)lane) : 0u;
    const uint32_t rz_bar = (lane < (int)CS) ? map_to_rank(zfull, (uint32_t)lane) : 0u;
    bool alive = finite_h;

    for (int prod = 0; alive; ++prod)
    {
      if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+0] = cg_now();
      // ================================================================ rows: part[cid] = sum_k conj(O_kp) (O_k . v)
      {
        // product 0: v = x0.  product 1: d_1 = M^-1 r_0 from pb[1] (written by product 0, grid barrier B behind it).  product k >= 2:
        // d_k = z_{k-1} + beta_{k-1} d_{k-1} is formed HERE from z (visible since the second grid sum of product k-1) and d_{k-1}
        // (written during product k-2): nobody has to wait for the owners to store d_k, so iterations need no barrier B.
        // The owner forms its own copy with the same fma: both are the same bits.
        cd vr[CPT], acc[CPT];
        {
          const cd * v = (prod == 0) ? a.x : a.pb[(prod == 1) ? 1 : ((prod-1)&1)];
          const double beta_prev = fsc[4];
#pragma unroll
          for (int c = 0; c < CPT; ++c)
          {
            const int idx = c*NT+tid;
            vr[c] = (idx < n_r) ? __ldcg(v+c0+idx) : cmake(0.0, 0.0);
            acc[c] = cmake(0.0, 0.0);
          }
          if (prod >= 2)
          {
#pragma unroll
            for (int c = 0; c < CPT; ++c)
            {
              const int idx = c*NT+tid;
              const cd zz = (idx < n_r) ? __ldcg(a.zv+c0+idx) : cmake(0.0, 0.0);
              vr[c] = cmake(fma(beta_prev, vr[c].x, zz.x), fma(beta_prev, vr[c].y, zz.y));
            }
          }
        }
        // (3) for the oldest row still waiting for it: z_k = sum of the CS partials in rank order, acc += conj(O_kp) z_k
        auto wait_z = [&](double & zx, double & zy)
        {
          const int tzq = gt&(NQS_SV_ZBUFS-1);
          mbar_wait_a(zfull_a+8u*(uint32_t)tzq, (uint32_t)((gt>>3)&1));
          const cd * zrow = zbuf+tzq*NQS_SV_MAX_CLUSTER;
          zx = 0.0; zy = 0.0;
          for (unsigned int r = 0; r < CS; ++r)
          {
            const cd t = zrow[r];
            zx += t.x; zy += t.y;
          }
          ++gt;
        };
        auto pass3_from_slot = [&]()
        {
          double zx, zy;
          wait_z(zx, zy);
          const cd * prow = sbase+(size_t)tail_slot*slot_elems;
#pragma unroll
          for (int c = 0; c < CPT; ++c)
          {
            const cd q = prow[c*NT];
            acc[c].x = fma(q.x, zx, acc[c].x); acc[c].x = fma(q.y, zy, acc[c].x);
            acc[c].y = fma(q.x, zy, acc[c].y); acc[c].y = fma(-q.y, zx, acc[c].y);
          }
          __syncwarp();
          if (lane == 0 && n_r > 0) mbar_arrive_a(empty_a+8u*(uint32_t)tail_slot);
          if (++tail_slot == a.nslot) tail_slot = 0;
        };
        for (int it = 0; it < nrows; ++it)
        {
          cd o[CPT];
          const cd * srow = sbase+(size_t)slot*slot_elems;
          if (n_r > 0) mbar_wait_a(full_a+8u*(uint32_t)slot, full_par);
#pragma unroll
          for (int c = 0; c < CPT; ++c) o[c] = srow[c*NT];
          if (DEFER == 0)
          {
            __syncwarp();
            if (lane == 0 && n_r > 0) mbar_arrive_a(empty_a+8u*(uint32_t)slot);
          }
          double pa = 0.0, pb = 0.0, pc_ = 0.0, pd = 0.0;
#pragma unroll
          for (int c = 0; c < CPT; ++c)
          {
            pa = fma(o[c].x, vr[c].x, pa); pb = fma(o[c].y, vr[c].y, pb);
            pc_ = fma(o[c].x, vr[c].y, pc_); pd = fma(o[c].y, vr[c].x, pd);
          }
          const cd wp = warp_sum(cmake(pa-pb, pc_+pd));
          const int rq = gi&(NQS_SV_RBUFS-1), zq = gi&(NQS_SV_ZBUFS-1);
          cd * redrow = red+rq*NQS_SV_MAX_WARPS;
          if (lane == 0) { redrow[w] = wp; mbar_arrive_a(wfull_a+8u*(uint32_t)rq); }
          if (w == redw)
          { // this row's reducer warp: CTA partial = fixed-order fold of the warp partials, sent to every CTA of the cluster
            mbar_wait_a(wfull_a+8u*(uint32_t)rq, (uint32_t)((gi>>2)&1));
            const cd sred = warp_sum((lane < NW) ? redrow[lane] : cmake(0.0, 0.0));
            if (lane == 0) mbar_expect_tx_a(zfull_a+8u*(uint32_t)zq, CS*(uint32_t)sizeof(cd));
            if (lane < (int)CS)
              st_async_remote_cd(rz_data+(uint32_t)zq*(uint32_t)(NQS_SV_MAX_CLUSTER*sizeof(cd)), sred, rz_bar+8u*(uint32_t)zq);
          }
          ++gi;
          if (++redw == NW) redw = 0;
          if (DEFER == 0)
          {
            double zx, zy;
            wait_z(zx, zy);
#pragma unroll
            for (int c = 0; c < CPT; ++c)
            {
              acc[c].x = fma(o[c].x, zx, acc[c].x); acc[c].x = fma(o[c].y, zy, acc[c].x);
              acc[c].y = fma(o[c].x, zy, acc[c].y); acc[c].y = fma(-o[c].y, zx, acc[c].y);
            }
          }
          else if (it >= depth) pass3_from_slot();   // row it-depth: its exchange travelled while this warp worked on the rows after it
          if (++slot == a.nslot) { slot = 0; full_par ^= 1u; }
        }
        if (DEFER != 0)
          for (int jt = (nrows > depth ? nrows-depth : 0); jt < nrows; ++jt) pass3_from_slot();
        double * base = a.part+(size_t)cid*2*(size_t)P;
#pragma unroll
        for (int c = 0; c < CPT; ++c)
        {
          const int idx = c*NT+tid;
          if (idx < n_r) { base[c0+idx] = acc[c].x; base[P+c0+idx] = acc[c].y; }
        }
      }
      if (tracing && prod < a.trace_max) a.trace[(size_t)prod*NQS_CGP_TRACE_WORDS+1] = cg_now();
      // ================================================================ A, vector phase, (B): kept OUT of line on purpose -- inlined, its
      // live ranges pushed the row loop above (which sits at the register limit) into spilling: 27 local-memory accesses per row
      // and 15 % on the pass over O
      {
        CgpCtx cx;
        cx.sh = sh; cx.fsc = fsc; cx.ctl = ctl; cx.epoch = epoch; cx.nsum = nsum; cx.NT = NT; cx.CS = (int)CS; cx.tracing = tracing;
        const int stop = cgp_vector_phase<EPT>(a, prod, cx);
        epoch = cx.epoch; nsum = cx.nsum;
        if (stop) break;
      }
    }
    if (!finite_h && blockIdx.x == 0 && tid == 0) { a.sc->nonfinite = 1; a.sc->done = 1; a.sc->iters = 0; }
    // tell the producer how far the consumers got, then leave the grid-barrier counter at zero for the next launch
    cgp_cta_sync(NT);
    if (tid == 0)
    {
      ctl[1] = gi;
      __threadfence_block();
      ctl[0] = 1;
      const unsigned int n = atomicAdd(a.barrier, 1u);
      if (n == (epoch+1)*gridDim.x-1) *a.barrier = 0u;
    }
  }
  cluster.sync();   // no CTA may exit while a peer can still write into its shared memory
}
} // namespace nqs
