// One-pass S*v:  part[q][p] = sum_{k in rows of cluster q} conj(O_kp) (sum_p' O_kp' v_p')   with O read from HBM ONCE.
//
// ref: SMatrixForCG::dot, gpu/include/functor_for_CG.cuh:104-127 = c8 Zgemm(1xKxP) z = O v, then c9 Zgemv(PxK) O^H z: two
// full passes over O (2*K*P*16 B per CG iteration).  z_k needs the WHOLE row k before O^H z can use it, and the column
// accumulators of a full row (P*16 B = 530 KB at N=128, M=256) do not fit one SM, so a single CTA cannot fuse the two.
// A thread-block CLUSTER can: the CS CTAs of a cluster split the columns, CTA r owns the slice [r*pc, (r+1)*pc) and keeps
// its v slice and its column accumulators in REGISTERS (CPT columns per thread) for the whole launch.  For every row of the
// cluster's row block:
//   (1) a dedicated PRODUCER warp streams the slice of row k into shared memory by TMA (cp.async.bulk + mbarrier
//       complete_tx), NSLOT rows deep, re-arming a slot as soon as the consumer warps released it (per-slot "empty"
//       mbarrier), so the HBM stream never waits for the math;
//   (2) every consumer thread pulls its CPT elements of the row into registers and forms its part of O_k . v; a shuffle
//       butterfly gives the warp partial, the warp partials meet on a CTA-local mbarrier, and one (rotating) reducer warp
//       folds them in a fixed order and sends the CTA partial into every CTA of the cluster over DSMEM with st.async (the
//       store itself completes bytes on the receiver's mbarrier).  No __syncthreads and no cluster barrier in the loop;
//   (3) after the mbarrier wait every warp adds the CS CTA partials in rank order (bit-identical z_k in all CTAs,
//       run-to-run deterministic) and accumulates conj(O_kp) z_k.
// The reduction + DSMEM round trip of (2) costs ~0.7 us per row (measured, profiles/r1d_sv_fused_experiments.md), more than
// half of the 1.2 us the HBM stream needs per row slice.  DEFER = 1 therefore software-pipelines the loop by one row: the
// z exchange of row k is in flight while the warps already run (2) on rows k+1 .. k+depth, and (3) for row k re-reads the
// slice from its shared-memory slot (released only then) instead of holding it in registers; depth = 1 when only three row
// slots fit (wide slices), up to 3 for narrow ones.  DEFER = 0 keeps the row in registers and releases the slot right after (2).
// HBM traffic: K*P*16 B per S*v instead of 2*K*P*16 B.  Cluster partials go to part[q][{re,im}][P] and are folded in fixed
// order by colsum_reduce_kernel exactly like the two-pass kernels' row-block partials.
#pragma once
#include <cooperative_groups.h>
#include "device_math.cuh"

namespace nqs
{
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t map_to_rank(const void * local, const uint32_t rank)
{
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
  return r;
}
// 16-byte asynchronous store into a peer CTA's shared memory that completes 16 bytes of transaction count on the PEER's
// mbarrier (st.async): data and signal travel together, so the consumer needs only the ordinary CTA-scope mbarrier wait.
// (A release-arrive + acquire.cluster wait works too, but every acquire at cluster scope costs an L1 invalidate,
// CCTL.IVALL, per polling thread -- 46 % of all stall samples in profiles/r1c_sv_fused_v1_summary.md.)
__device__ __forceinline__ void st_async_remote_cd(const uint32_t remote_addr, const cd v, const uint32_t remote_bar)
{
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
    :: "r"(remote_addr), "d"(v.x), "d"(v.y), "r"(remote_bar) : "memory");
}
struct SvArgs
{
  long long K, P;
  const cd * O;            // [K][P]
  const cd * v;            // [P]
  double * part;           // [n_clusters][2][P]
  const int * done;        // device flag: converged CG -> return immediately (may be nullptr)
  long long pc;            // columns per CTA slice = ceil(P / cluster size)
  long long rows_per_cluster;
  int nslot;               // shared-memory row slots (TMA pipeline depth)
  unsigned int slot_bytes; // bytes per slot: >= CPT * consumer threads * 16 (the tail past the slice stays zero)
  int depth;               // DEFER = 1: pass (3) of row k runs after pass (2) of row k+depth (1..NQS_SV_MAX_DEPTH, <= nslot-2)
  // GEN = 1 (RBM): the rows of O do not exist yet.  The producer stages the FACTORS of row k -- T_k = tanh(theta_k) [M] and the
  // spins [N] int8 -- and every consumer thread forms its elements O_kp = s_ki T_kj itself, uses them for the product and
  // WRITES them to O (coalesced 16-byte stores): the O writer (ref RBM::backward, k13 :1426-1449) and the first S*v of the
  // CG (S x0) cost one HBM write pass instead of a write pass plus a read pass.  When the consumer thread count is a multiple
  // of M (gen_q = NT / M > 0) every column of a thread inside the W block has the SAME hidden unit j and sites gen_q apart:
  // one T load and a sign flip per element (s = +-1) instead of two loads and two multiplications.
  const cd * T;            // [K][M]
  const int8_t * spins8;   // [K][N], N a multiple of 16
  cd * Ow;                 // [K][P] out
  int N, M;
  int gen_q;
};
// GEN slot layout: T row [M] | (1,0) | spins [N] int8 | +1 | 0   (the three constants serve the a / b blocks and the padding)
inline size_t sv_gen_slot_bytes(int N, int M) { return (size_t)(M+1)*16+(((size_t)(N+2)+15)/16)*16; }

#define NQS_SV_MAX_CLUSTER 16
#define NQS_SV_MAX_SLOTS 8
#define NQS_SV_MAX_WARPS 32
#define NQS_SV_ZBUFS 8     // z exchange buffers, >= 2*depth+2: a peer may run (2) up to depth+1 rows ahead of this CTA's (3)
#define NQS_SV_MAX_DEPTH 3
#define NQS_SV_RBUFS 4     // CTA-level reduction buffers, >= depth+1: a warp may run (2) up to depth rows ahead of a reducer
// shared memory after the slots: red[RBUFS][warps] | zbuf[ZBUFS][cluster] | full[8] | empty[8] | wfull[RBUFS] | zfull[ZBUFS]
#define NQS_SV_TAIL_CORE (NQS_SV_RBUFS*NQS_SV_MAX_WARPS*16+NQS_SV_ZBUFS*NQS_SV_MAX_CLUSTER*16+2*NQS_SV_MAX_SLOTS*8+NQS_SV_RBUFS*8+NQS_SV_ZBUFS*8)
// + the reduction scratch [warps][8], the recurrence scalars and the control words of the persistent CG kernel (cg_persist.cuh), so both kernels share one launch plan
#define NQS_SV_TAIL_BYTES (NQS_SV_TAIL_CORE+NQS_SV_MAX_WARPS*8*8+64+64)

// register budget (16384 registers per SM sub-partition): CPT <= 3 runs up to 992+32 threads (8 warps per sub-partition x 64
// registers), larger CPT up to 480+32 threads (4 warps per sub-partition x 128 registers)
#define NQS_SV_MAX_CPT 10
template <int CPT, int GEN = 0> struct SvMaxRegs { static const int value = (CPT <= 3) ? 64 : 128; };

// blockDim.x = 32*(NW+1): warps 0..NW-1 consume (NT = 32*NW threads own the columns), warp NW is the TMA producer.
template <int CPT, int DEFER, int GEN>
__global__ void __maxnreg__((SvMaxRegs<CPT, GEN>::value)) sv_fused_kernel(const SvArgs a)
{
  if (a.done != nullptr && *a.done) return;   // uniform over the grid
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const unsigned int CS = cluster.num_blocks(), crank = cluster.block_rank();
  const int NT = blockDim.x-32, tid = threadIdx.x, lane = tid&31, w = tid>>5, NW = NT>>5;
  const long long cid = blockIdx.x/CS;
  unsigned char * tail = smem_raw+(size_t)a.nslot*a.slot_bytes;
  cd * red = reinterpret_cast<cd*>(tail);                                   // [RBUFS][NQS_SV_MAX_WARPS] warp partials of this CTA
  cd * zbuf = red+NQS_SV_RBUFS*NQS_SV_MAX_WARPS;                                       // [ZBUFS][NQS_SV_MAX_CLUSTER] CTA partials of the cluster
  uint64_t * full = reinterpret_cast<uint64_t*>(zbuf+NQS_SV_ZBUFS*NQS_SV_MAX_CLUSTER); // [slots] TMA arrival
  uint64_t * empty = full+NQS_SV_MAX_SLOTS;                                 // [slots] all consumer warps released the slot
  uint64_t * wfull = empty+NQS_SV_MAX_SLOTS;                                // [RBUFS] all NW warp partials of a row are in red[]
  uint64_t * zfull = wfull+NQS_SV_RBUFS;                                               // [ZBUFS] all CS CTA partials of a row arrived (tx bytes)

  const long long c0 = (long long)crank*a.pc;
  long long nr_ll = a.P-c0;
  if (nr_ll > a.pc) nr_ll = a.pc;
  if (nr_ll < 0) nr_ll = 0;
  const int n_r = (int)nr_ll;                                               // columns of this CTA's slice
  const long long k0 = cid*a.rows_per_cluster;
  long long k1 = k0+a.rows_per_cluster;
  if (k1 > a.K) k1 = a.K;
  const int nrows = (k1 > k0) ? (int)(k1-k0) : 0;
  const uint32_t row_bytes = (uint32_t)n_r*(uint32_t)sizeof(cd);

  // zero the slots once: TMA only ever writes the first row_bytes of a slot, so elements past the slice read as 0 and the
  // row loop needs no bounds predicate at all
  {
    double2 * z = reinterpret_cast<double2*>(smem_raw);
    const int nz16 = (int)(((size_t)a.nslot*a.slot_bytes)/sizeof(double2));
    for (int i = tid; i < nz16; i += blockDim.x) z[i] = make_double2(0.0, 0.0);
  }
  if (GEN)
  {
    __syncthreads();
    for (int s = tid; s < a.nslot; s += blockDim.x)
    {
      cd * Ts = reinterpret_cast<cd*>(smem_raw+(size_t)s*a.slot_bytes);
      Ts[a.M] = cmake(1.0, 0.0);
      reinterpret_cast<int8_t*>(Ts+a.M+1)[a.N] = (int8_t)1;
    }
  }
  if (tid == 0)
  {
    for (int s = 0; s < a.nslot; ++s) { mbar_init(full+s, 1); mbar_init(empty+s, (uint32_t)NW); }
    for (int q = 0; q < NQS_SV_RBUFS; ++q) mbar_init(wfull+q, (uint32_t)NW);
    for (int q = 0; q < NQS_SV_ZBUFS; ++q) mbar_init(zfull+q, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zero fill before async-proxy (TMA) writes
  cluster.sync();   // every CTA's barriers exist before anybody signals them

  if (w == NW)
  { // ---- (1) producer warp: one lane keeps NSLOT rows of this CTA's column slice in flight
    if (lane == 0 && n_r > 0)
    {
      const cd * Oslice = a.O+c0;
      int slot = 0;
      uint32_t par = 0;
      for (int it = 0; it < nrows; ++it)
      {
        if (it >= a.nslot) mbar_wait(empty+slot, par^1u);   // consumers released the row that used this slot before
        if (GEN)
        {
          unsigned char * dst = smem_raw+(size_t)slot*a.slot_bytes;
          mbar_expect_tx(full+slot, (uint32_t)a.M*16u+(uint32_t)a.N);
          tma_load_1d(dst, a.T+(k0+it)*a.M, (uint32_t)a.M*16u, full+slot);
          tma_load_1d(dst+(size_t)(a.M+1)*16, a.spins8+(k0+it)*a.N, (uint32_t)a.N, full+slot);
        }
        else
        {
          mbar_expect_tx(full+slot, row_bytes);
          tma_load_1d(smem_raw+(size_t)slot*a.slot_bytes, Oslice+(k0+it)*a.P, row_bytes, full+slot);
        }
        if (++slot == a.nslot) { slot = 0; par ^= 1u; }
      }
    }
  }
  else
  {
    cd vr[CPT], acc[CPT];
#pragma unroll
    for (int c = 0; c < CPT; ++c)
    {
      const int idx = c*NT+tid;
      vr[c] = (idx < n_r) ? a.v[c0+idx] : cmake(0.0, 0.0);
      acc[c] = cmake(0.0, 0.0);
    }
    const cd * const sbase = reinterpret_cast<const cd*>(smem_raw)+tid;
    const size_t slot_elems = a.slot_bytes/sizeof(cd);
    // GEN: where the two factors of column c sit in a slot, packed (T index | spin index << 16)
    int fac[GEN ? CPT : 1];
    bool fastp = false;      // all CPT columns of this thread lie in the W block and share their hidden unit
    int jf = 0, i0 = 0;
    if (GEN)
    {
      const long long NM = (long long)a.N*a.M;
      if (a.gen_q > 0 && (CPT-1)*NT+tid < n_r && c0+(long long)(CPT-1)*NT+tid < NM)
      {
        fastp = true;
        jf = (int)((c0+tid)%a.M); i0 = (int)((c0+tid)/a.M);
      }
#pragma unroll
      for (int c = 0; c < CPT; ++c)
      {
        const int idx = c*NT+tid;
        const long long p = c0+idx;
        int tj, si;
        if (idx >= n_r) { tj = a.M; si = a.N+1; }                              // padding: (1,0) * 0
        else if (p < NM) { tj = (int)(p%a.M); si = (int)(p/a.M); }             // s_ki T_kj
        else if (p < NM+a.N) { tj = a.M; si = (int)(p-NM); }                   // s_ki
        else { tj = (int)(p-NM-a.N); si = a.N; }                               // T_kj
        fac[GEN ? c : 0] = tj|(si<<16);
      }
    }
    // element c of the row staged in slot gslot; trow = T of the thread's hidden unit in that row (fast path)
    auto gen_o = [&](const int c, const int gslot, const cd trow) -> cd
    {
      const cd * Ts = reinterpret_cast<const cd*>(smem_raw+(size_t)gslot*a.slot_bytes);
      const int8_t * sb = reinterpret_cast<const int8_t*>(Ts+a.M+1);
      if (fastp)
      { // s = +-1: flip the sign bits
        const int m = ((int)sb[i0+c*a.gen_q])&0x80000000;
        return cmake(__hiloint2double(__double2hiint(trow.x)^m, __double2loint(trow.x)), __hiloint2double(__double2hiint(trow.y)^m, __double2loint(trow.y)));
      }
      const int f = fac[GEN ? c : 0];
      const cd t = Ts[f&0xffff];
      const double sp = (double)sb[f>>16];
      return cmake(t.x*sp, t.y*sp);
    };
    auto gen_trow = [&](const int gslot) -> cd
    {
      return fastp ? reinterpret_cast<const cd*>(smem_raw+(size_t)gslot*a.slot_bytes)[jf] : cmake(0.0, 0.0);
    };

    // (3) for row jt: z_k = sum of the CS partials in rank order, then acc += conj(O_kp) z_k
    auto wait_z = [&](const int jt, double & zx, double & zy)
    {
      const int zq = jt&(NQS_SV_ZBUFS-1);
      mbar_wait(zfull+zq, (uint32_t)((jt/NQS_SV_ZBUFS)&1));
      const cd * zrow = zbuf+zq*NQS_SV_MAX_CLUSTER;
      zx = 0.0; zy = 0.0;
      for (unsigned int r = 0; r < CS; ++r)
      {
        const cd t = zrow[r];
        zx += t.x; zy += t.y;
      }
    };

    // (3) for row jt held in slot jslot: re-read the slice from shared memory, then release the slot
    auto pass3_from_slot = [&](const int jt, const int jslot)
    {
      double zx, zy;
      wait_z(jt, zx, zy);
      const cd * prow = sbase+(size_t)jslot*slot_elems;
      const cd trow = GEN ? gen_trow(jslot) : cmake(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < CPT; ++c)
      {
        const cd q = GEN ? gen_o(c, jslot, trow) : prow[c*NT];
        acc[c].x = fma(q.x, zx, acc[c].x); acc[c].x = fma(q.y, zy, acc[c].x);
        acc[c].y = fma(q.x, zy, acc[c].y); acc[c].y = fma(-q.y, zx, acc[c].y);
      }
      __syncwarp();
      if (lane == 0 && n_r > 0) mbar_arrive(empty+jslot);
    };

    const int depth = a.depth;
    int slot = 0, tail_slot = 0;       // tail_slot: slot of row it-depth (the oldest row still waiting for its pass (3))
    uint32_t full_par = 0;
    for (int it = 0; it < nrows; ++it)
    {
      // ---- (2) row slice -> registers, partial O_k . v
      cd o[CPT];
      const cd * srow = sbase+(size_t)slot*slot_elems;
      if (n_r > 0) mbar_wait(full+slot, full_par);
      if (GEN)
      {
        cd * orow = a.Ow+(k0+it)*a.P+c0+tid;
        const cd trow = gen_trow(slot);
#pragma unroll
        for (int c = 0; c < CPT; ++c)
        {
          o[c] = gen_o(c, slot, trow);
          if (fastp || c*NT+tid < n_r) orow[c*NT] = o[c];
        }
      }
      else
      {
#pragma unroll
        for (int c = 0; c < CPT; ++c) o[c] = srow[c*NT];
      }
      if (DEFER == 0)
      {
        __syncwarp();
        if (lane == 0 && n_r > 0) mbar_arrive(empty+slot);
      }
      double pa = 0.0, pb = 0.0, pc_ = 0.0, pd = 0.0;   // four independent chains
#pragma unroll
      for (int c = 0; c < CPT; ++c)
      {
        pa = fma(o[c].x, vr[c].x, pa); pb = fma(o[c].y, vr[c].y, pb);
        pc_ = fma(o[c].x, vr[c].y, pc_); pd = fma(o[c].y, vr[c].x, pd);
      }
      const cd wp = warp_sum(cmake(pa-pb, pc_+pd));
      const int rq = it&(NQS_SV_RBUFS-1), zq = it&(NQS_SV_ZBUFS-1);
      cd * redrow = red+rq*NQS_SV_MAX_WARPS;
      if (lane == 0) { redrow[w] = wp; mbar_arrive(wfull+rq); }
      if (w == it%NW)
      { // this row's reducer warp: CTA partial = fixed-order fold of the warp partials, sent to every CTA of the cluster
        mbar_wait(wfull+rq, (uint32_t)((it/NQS_SV_RBUFS)&1));
        const cd s = warp_sum((lane < NW) ? redrow[lane] : cmake(0.0, 0.0));
        if (lane == 0) mbar_expect_tx(zfull+zq, CS*(uint32_t)sizeof(cd));   // this row's CS partials land here
        if (lane < (int)CS)
          st_async_remote_cd(map_to_rank(zbuf+zq*NQS_SV_MAX_CLUSTER+crank, (uint32_t)lane), s, map_to_rank(zfull+zq, (uint32_t)lane));
      }
      if (DEFER == 0)
      {
        double zx, zy;
        wait_z(it, zx, zy);
#pragma unroll
        for (int c = 0; c < CPT; ++c)
        {
          acc[c].x = fma(o[c].x, zx, acc[c].x); acc[c].x = fma(o[c].y, zy, acc[c].x);
          acc[c].y = fma(o[c].x, zy, acc[c].y); acc[c].y = fma(-o[c].y, zx, acc[c].y);
        }
      }
      else if (it >= depth)
      { // row it-depth: its exchange travelled while this warp worked on the rows after it; the slice is still in its slot
        pass3_from_slot(it-depth, tail_slot);
        if (++tail_slot == a.nslot) tail_slot = 0;
      }
      if (++slot == a.nslot) { slot = 0; full_par ^= 1u; }
    }
    if (DEFER != 0)
    {
      for (int jt = (nrows > depth ? nrows-depth : 0); jt < nrows; ++jt)
      {
        pass3_from_slot(jt, tail_slot);
        if (++tail_slot == a.nslot) tail_slot = 0;
      }
    }
    double * base = a.part+(size_t)cid*2*(size_t)a.P;
#pragma unroll
    for (int c = 0; c < CPT; ++c)
    {
      const int idx = c*NT+tid;
      if (idx < n_r) { base[c0+idx] = acc[c].x; base[a.P+c0+idx] = acc[c].y; }
    }
  }
  cluster.sync();   // no CTA may exit while a peer can still write into its shared memory
}
} // namespace nqs
