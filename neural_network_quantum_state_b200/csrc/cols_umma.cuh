// part[chunk][{re,im}][p] = sum_{k in chunk} conj(O_kp) z_k from the factors of O (the O^H z half of the structured S*v, and the
// SR setup sums with z = 1 / z = htilda) on the 5th-generation tensor cores: S^T C with C_kj = conj(T_kj) z_k, as an int8 UMMA
// (tcgen05.mma kind::i8, accumulators in TMEM) with fp64-grade results.  Same outputs as spin_cols_dmma_kernel (sv_struct.cuh);
// ref: Zgemv pair of SMatrixForCG::dot, gpu/include/functor_for_CG.cuh:104-127, and the sums of optimizer.cuh:140-143.
//
// The spins are +-1 (exact in int8).  C is formed on the fly in fp64 and split error-free into 7 balanced base-256 digits
//   Q = rint(C 2^(54-e_j)),  |Q| < 2^54,  Q + 0x80808080808080 = the 7 unsigned bytes whose XOR with 0x80 are the digits
// with one power-of-two scale per hidden unit: 2^e_j >= 2 max_k|T_kj|_inf max_k|z_k|_inf >= |Re C|, |Im C| (max|T| over the
// rank's chains comes from colmax_abs_kernel, once per SR step; max|z| over the CTA's chains is taken here).  The int32
// accumulators are exact for chunks of up to 2^24 chains and recombined like in rows_umma.cuh; the only rounding is Q's:
// <= 2^(e_j-55) per element, i.e. 2^-54 of the bound, below the rounding of the fp64 sum it replaces.
//
// Geometry: a CTA owns 64 real columns (32 hidden units) and a chunk of chains; UMMA M = 128 sites (TMEM lane = site; N <= 128),
// UMMA N = 7 planes x 64 columns = 448 = two instructions of 224, UMMA K = chains, 64 per block (2 K steps).  Both operands are
// MN-major (the natural layouts: sites contiguous in spins[k][.], columns contiguous in C[k][.]), no swizzle: core matrix =
// 8 chains x 16 bytes; blocks of 16 sites / 16 columns NQS_CU_SBO bytes apart (a multiple of 16 that is not a multiple of 128:
// conflict-free stores), groups of 8 chains 128 bytes apart.  A thread produces one (chain, 8 columns) task per block: the
// columns of a tile are ordered (lane-of-8 c, u, re/im) <-> hidden unit j0 + 8 u + c, so that its 4 loads of T are 16-byte
// pieces of 128-byte runs across 8 lanes and its 8 digits per plane are one 8-byte store.  The tiles are double-buffered: the
// MMAs of block b (one elected thread, tcgen05.commit -> mbarrier) run under the production of block b+1.
#pragma once
#include "rows_umma.cuh"

namespace nqs
{

#define NQS_CU_THREADS 512
#define NQS_CU_NCC 64                         // real columns per CTA
#define NQS_CU_KB 64                          // chains per block
#define NQS_CU_SBO 1056                       // (NQS_CU_KB/8)*128 + 32: consecutive 16-column blocks start 8 banks apart, which makes the
                                              // 8-byte digit stores of a half-warp (2 chains x 8 lanes x 4 blocks) conflict-free
#define NQS_CU_A_BYTES (8*NQS_CU_SBO)         // 128 sites
#define NQS_CU_NBLK (7*NQS_CU_NCC/16)         // 28 column blocks of 16
#define NQS_CU_B_BYTES (NQS_CU_NBLK*NQS_CU_SBO)
#define NQS_CU_BUF_BYTES (NQS_CU_A_BYTES+NQS_CU_B_BYTES)
#define NQS_CU_OUT_PITCH 33                   // doubles per site in the staged output tile
#define NQS_CU_TMEM_COLS 512

// tiles [2] (reused as the staged output tile [2][128][33] doubles) | z of the block [2][64] | reductions [16][64] x 2 | site sums [4][128] cd | barriers
inline size_t cols_umma_smem()
{
  const size_t tiles = (size_t)2*NQS_CU_BUF_BYTES, outt = (size_t)2*128*NQS_CU_OUT_PITCH*sizeof(double);
  const size_t need = (tiles > outt ? tiles : outt)+(size_t)2*NQS_CU_KB*sizeof(cd)+(size_t)2*16*64*sizeof(double)+64+(size_t)64*sizeof(double);
  return need > (size_t)120*1024 ? need : (size_t)120*1024;     // more than half an SM: one CTA per SM, each owns the 512 TMEM columns
}

// max over the chains of max(|Re T_kj|, |Im T_kj|) per hidden unit, as the bit pattern of a non-negative double (ordered like
// the integers; +Inf marks a column holding a NaN or an Inf).  out must be zeroed first.
__global__ void __launch_bounds__(256) colmax_abs_kernel(const long long K, const int M, const cd * __restrict__ T, unsigned long long * __restrict__ out,
                                                         const long long rows_per_block)
{
  __shared__ double sm[8][32];
  const int lane = threadIdx.x&31, w = threadIdx.x>>5, j = blockIdx.x*32+lane;
  const long long k0 = (long long)blockIdx.y*rows_per_block, k1 = (k0+rows_per_block < K) ? k0+rows_per_block : K;
  double m = 0.0;
  if (j < M)
    for (long long k = k0+w; k < k1; k += 8)
    {
      const cd t = T[k*M+j];
      const double v = fmax(fabs(t.x), fabs(t.y));
      m = (fabs(t.x) <= DBL_MAX && fabs(t.y) <= DBL_MAX) ? fmax(m, v) : INFINITY;
    }
  sm[w][lane] = m;
  __syncthreads();
  if (w == 0 && j < M)
  {
#pragma unroll
    for (int ww = 1; ww < 8; ++ww) m = fmax(m, sm[ww][lane]);
    atomicMax(out+j, (unsigned long long)__double_as_longlong(m));
  }
}

struct ColsUmmaArgs
{
  ColsArgs c;
  const unsigned long long * tmax;   // [M] from colmax_abs_kernel
};

__device__ __forceinline__ void cu_tmem_ld8(const uint32_t taddr, uint32_t (&v)[8]) { ru_tmem_ld8(taddr, v); }

template <int MODEL>
__global__ void __launch_bounds__(NQS_CU_THREADS, 1) spin_cols_umma_kernel(const ColsUmmaArgs args)
{
  const ColsArgs & a = args.c;
  if (a.done != nullptr && *a.done) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int KB = NQS_CU_KB, SBO = NQS_CU_SBO;
  const int N = a.N, M = a.M;
  const size_t tiles = (size_t)2*NQS_CU_BUF_BYTES, outt = (size_t)2*128*NQS_CU_OUT_PITCH*sizeof(double);
  unsigned char * tile = smem_raw;                                             // [2][A | B]
  double * outs = reinterpret_cast<double*>(smem_raw);                         // [2][128][33] after the last MMA
  cd * zs = reinterpret_cast<cd*>(smem_raw+(tiles > outt ? tiles : outt));     // [2][KB]
  double * red = reinterpret_cast<double*>(zs+2*KB);                           // [2][16 warps][64]
  uint64_t * mdone = reinterpret_cast<uint64_t*>(red+2*16*64);                 // [2]
  uint32_t * tptr = reinterpret_cast<uint32_t*>(mdone+2);
  double * zred = reinterpret_cast<double*>(tptr+2);                           // [1] max |z|
  double * upt = zred+1;                                                       // [32] 2^(54-e_j): C -> Q
  double * sct = upt+32;                                                       // [32] 2^(e_j-54) (NaN for a non-finite column)
  const int tid = threadIdx.x, lane = tid&31, w = tid>>5;
  const int j0 = blockIdx.x*(NQS_CU_NCC/2);                                    // first hidden unit of this column group
  const long long k0 = (long long)blockIdx.y*a.rows_per_chunk;
  const long long k1 = (k0+a.rows_per_chunk < a.K) ? k0+a.rows_per_chunk : a.K;
  const int nblocks = (int)((k1-k0+KB-1)/KB);
  const bool do_a = (MODEL == MODEL_RBM && blockIdx.x == 0);
  const bool ones = (a.zmode == 1 && blockIdx.z == 0);

  if (w == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tptr)), "r"((uint32_t)NQS_CU_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32)
  {
    mbar_init(mdone, 1); mbar_init(mdone+1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // max |z| over this CTA's chains (red[] as scratch)
  {
    double m = ones ? 1.0 : 0.0;
    if (!ones)
      for (long long k = k0+tid; k < k1; k += NQS_CU_THREADS)
      {
        const cd z = a.zk[k];
        m = (fabs(z.x) <= DBL_MAX && fabs(z.y) <= DBL_MAX) ? fmax(m, fmax(fabs(z.x), fabs(z.y))) : INFINITY;
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[w] = m;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tptr;
  if (tid < 32)
  { // lane = hidden unit j0 + tid of this column group: its power-of-two scales
    double zmax = red[0];
    for (int ww = 1; ww < NQS_CU_THREADS/32; ++ww) zmax = fmax(zmax, red[ww]);
    const int j = j0+tid;
    const double tm = (j < M) ? __longlong_as_double((long long)args.tmax[j]) : 0.0;
    double upv = 0.0, scv = 0.0;      // all-zero column, or |C| below 1e-290 / above 1e+290: counts as zero
    if (tm > 0.0 && zmax > 0.0)
    {
      if (tm <= DBL_MAX && zmax <= DBL_MAX)
      {
        const int e = ilogb(tm)+ilogb(zmax)+3, bu = 1023+54-e, bs = 1023-54+e;
        if (bu >= 1 && bu <= 2046 && bs >= 1 && bs <= 2046) { upv = __hiloint2double(bu<<20, 0); scv = __hiloint2double(bs<<20, 0); }
      }
      else scv = nan("");             // a NaN / Inf among the factors: the column's results are NaN, like the fp64 GEMM's
    }
    upt[tid] = upv; sct[tid] = scv;
  }
  __syncthreads();

  // producer task of this thread: chain slot kk of the block, hidden units j0 + 8 u + c (u < 4)
  const int c = tid&7, kk = tid>>3;
  double up[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) up[u] = upt[8*u+c];
  // instruction descriptor: D = s32, A = B = s8, both MN-major, N = 224, M = 128
  constexpr uint32_t idesc = (2u<<4)|(1u<<7)|(1u<<10)|(1u<<15)|(1u<<16)|((uint32_t)(224>>3)<<17)|((128u>>4)<<24);

  cd tv[4];
  cd zv;
  uint4 sp;
  auto prefetch = [&](const int b)
  {
    const long long k = k0+(long long)b*KB+kk;
    const bool ok = (k < k1);
    zv = ok ? (ones ? cmake(1.0, 0.0) : a.zk[k]) : cmake(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      const int j = j0+8*u+c;
      tv[u] = (ok && j < M) ? a.T[k*M+j] : cmake(0.0, 0.0);
    }
    // spin piece: chain kk, sites 16 c .. 16 c + 15
    sp = make_uint4(0u, 0u, 0u, 0u);
    if (ok && 16*c < N)
    {
      const int8_t * src = a.spins+k*N+16*c;
      if ((N&15) == 0 && (reinterpret_cast<size_t>(a.spins)&15) == 0) sp = *reinterpret_cast<const uint4*>(src);
      else
      {
        uint32_t wv[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int t = 0; t < 16; ++t)
          if (16*c+t < N) wv[t>>2] |= (uint32_t)(unsigned char)src[t]<<(8*(t&3));
        sp = make_uint4(wv[0], wv[1], wv[2], wv[3]);
      }
    }
  };
  double bsum[8], b2sum[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) { bsum[t] = 0.0; b2sum[t] = 0.0; }
  double asx = 0.0, asy = 0.0;

#ifdef NQS_RU_TRACE_BUILD
  const bool trace = (blockIdx.x == 1 && blockIdx.y == 2 && blockIdx.z == 0 && tid == 96);
  long long ts[40];
  int nts = 0;
#define NQS_CU_STAMP() do { if (trace && b >= 4 && b < 8 && nts < 40) ts[nts++] = clock64(); } while (0)
#else
#define NQS_CU_STAMP() do { } while (0)
#endif
  if (nblocks > 0) prefetch(0);
  for (int b = 0; b < nblocks; ++b)
  {
    const int bb = b&1;
    NQS_CU_STAMP();
    unsigned char * At = tile+(size_t)bb*NQS_CU_BUF_BYTES;
    unsigned char * Bt = At+NQS_CU_A_BYTES;
    // the MMAs that read this buffer (block b-2) must have completed
    if (b >= 2) ru_mbar_wait(mdone+bb, (uint32_t)(((b>>1)-1)&1));
    NQS_CU_STAMP();
    // C = conj(T) z of the task, 8 reals in tile order (u, re/im)
    double cv[8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
    {
      cv[2*u] = tv[u].x*zv.x+tv[u].y*zv.y;
      cv[2*u+1] = tv[u].x*zv.y-tv[u].y*zv.x;
    }
    const cd zcur = zv;
    const uint4 spcur = sp;
    const double upc[4] = {up[0], up[1], up[2], up[3]};
    if (b+1 < nblocks) prefetch(b+1);
    NQS_CU_STAMP();
    // digits: bytes 0..6 of Q + 0x80..80, flipped to signed
    uint32_t qlo[8], qhi[8];
#pragma unroll
    for (int t = 0; t < 8; ++t)
    {
      bsum[t] += cv[t];
      b2sum[t] = fma(cv[t], cv[t], b2sum[t]);
      const long long Q = __double2ll_rn(cv[t]*upc[t>>1]);
      const unsigned long long d = (unsigned long long)(Q+0x0080808080808080ll)^0x0080808080808080ull;
      qlo[t] = (uint32_t)d; qhi[t] = (uint32_t)(d>>32);
    }
    NQS_CU_STAMP();
    // byte transpose: plane s takes byte s of the 8 values
#pragma unroll
    for (int s = 0; s < 7; ++s)
    {
      const uint32_t * q = (s < 4) ? qlo : qhi;
      const int sh = s&3;
      const uint32_t sel = 0x0040u+(uint32_t)sh*0x0011u;          // bytes: a[sh], b[sh]
      const uint32_t p01 = __byte_perm(q[0], q[1], sel), p23 = __byte_perm(q[2], q[3], sel);
      const uint32_t p45 = __byte_perm(q[4], q[5], sel), p67 = __byte_perm(q[6], q[7], sel);
      const uint32_t w0 = __byte_perm(p01, p23, 0x5410), w1 = __byte_perm(p45, p67, 0x5410);
      const int nb = s*4+(c>>1);                                  // 16-column block: n = s 64 + c 8 + (u, ri)
      *reinterpret_cast<uint2*>(Bt+(size_t)nb*SBO+(kk>>3)*128+(kk&7)*16+(c&1)*8) = make_uint2(w0, w1);
    }
    *reinterpret_cast<uint4*>(At+(size_t)c*SBO+(kk>>3)*128+(kk&7)*16) = spcur;
    if (c == 0) zs[bb*KB+kk] = zcur;
    NQS_CU_STAMP();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    NQS_CU_STAMP();
    if (tid == 0)
    {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t adesc = ru_desc(smem_u32(At), 128u, (uint32_t)SBO), bdesc = ru_desc(smem_u32(Bt), 128u, (uint32_t)SBO);
#pragma unroll
      for (int ks = 0; ks < KB/32; ++ks)          // 32 chains = 4 groups of 8 = 512 bytes further along both operands
#pragma unroll
        for (int hf = 0; hf < 2; ++hf)            // columns 0..223 / 224..447: 14 blocks of 16
          ru_mma_i8(tbase+(uint32_t)hf*224u, adesc+(uint64_t)(ks*32), bdesc+(uint64_t)(ks*32+hf*14*(SBO/16)), idesc, (b > 0 || ks > 0) ? 1u : 0u);
      ru_commit(mdone+bb);
    }
    if (do_a && tid < 128)
    { // RBM visible-bias block: sum_k s_ki z_k for site tid, from the tile just written
      if (tid < N)
      {
        const unsigned char * ap = At+(size_t)(tid>>4)*SBO+(tid&15);
#pragma unroll 8
        for (int q = 0; q < KB; ++q)
        {
          const double s = (double)(int8_t)ap[(q>>3)*128+(q&7)*16];
          const cd z = zs[bb*KB+q];
          asx = fma(s, z.x, asx); asy = fma(s, z.y, asy);
        }
      }
    }
    __syncwarp();
  }
#ifdef NQS_RU_TRACE_BUILD
  if (trace)
  {
    printf("cu trace nblocks %d:", nblocks);
    for (int i = 1; i < nts; ++i) printf("%s%lld", (i%6 == 0) ? " | top " : " ", ts[i]-ts[i-1]);
    printf("   (per block: mbar-wait, cv+prefetch, digits, stores, fence+sync, then top = MMA issue / do_a / loop)\n");
  }
#endif
#undef NQS_CU_STAMP
  // all MMAs complete when the last commit fires (commits complete in order)
  if (nblocks > 0) ru_mbar_wait(mdone+((nblocks-1)&1), (uint32_t)(((nblocks-1)>>1)&1));
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  __syncthreads();      // every thread is past its last tile access: the tile memory becomes the staged output

  const long long P = a.P, NM = (long long)N*M;
  double * base = a.part+(size_t)blockIdx.z*(size_t)a.part_stride+(size_t)blockIdx.y*2*(size_t)P;
  // ---- W block: TMEM lane = site, 4 column groups of warps take 2 of the 8 c each
  {
    const int q = w&3, g = w>>2, site = 32*q+lane;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
    {
      const int c2 = 2*g+cc;
      uint32_t v[7][8];
      if (nblocks > 0)
      {
#pragma unroll
        for (int s = 0; s < 7; ++s) cu_tmem_ld8(tbase+((uint32_t)(32*q)<<16)+(uint32_t)(s*64+c2*8), v[s]);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      else
      {
#pragma unroll
        for (int s = 0; s < 7; ++s)
#pragma unroll
          for (int t = 0; t < 8; ++t) v[s][t] = 0u;
      }
#pragma unroll
      for (int t = 0; t < 8; ++t)
      {
        const int u = t>>1, ri = t&1, j = j0+8*u+c2;
        // sum_s 256^s acc_s (|acc_s| <= 128 * chains per chunk): pairs of planes in int64, then hi 2^32 + lo in one FMA
        const long long p0 = (long long)(int)v[0][t]+256ll*(int)v[1][t], p1 = (long long)(int)v[2][t]+256ll*(int)v[3][t];
        const long long p2 = (long long)(int)v[4][t]+256ll*(int)v[5][t];
        const long long lo = p1*65536+p0, hi = (long long)(int)v[6][t]*65536+p2;     // exact in int64 for any chunk the int32 accumulators can hold
        const double sc = sct[8*u+c2];
        double val = fma((double)hi, 4294967296.0, (double)lo);
        val = (sc == 0.0) ? 0.0 : val*sc;
        if (MODEL == MODEL_RBM) outs[((size_t)ri*128+site)*NQS_CU_OUT_PITCH+8*u+c2] = val;
        else if (site < N && j < M) base[(size_t)ri*P+(size_t)j*N+site] = val;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((uint32_t)NQS_CU_TMEM_COLS) : "memory");
  if (MODEL == MODEL_RBM)
  { // p = i M + j: a warp writes the 32 hidden units of one site, both planes
    const int j = j0+lane;
    for (int site = w; site < N; site += NQS_CU_THREADS/32)
      if (j < M)
      {
        base[(size_t)site*M+j] = outs[((size_t)site)*NQS_CU_OUT_PITCH+lane];
        base[(size_t)P+(size_t)site*M+j] = outs[((size_t)128+site)*NQS_CU_OUT_PITCH+lane];
      }
  }
  // ---- short blocks: column sums of C (b block / FFNN b1 block) and of C^2 (setup), over the 4 lanes of a warp that share c,
  // then over the warps in fixed order
#pragma unroll
  for (int t = 0; t < 8; ++t)
  {
    bsum[t] += __shfl_xor_sync(0xffffffffu, bsum[t], 8);   bsum[t] += __shfl_xor_sync(0xffffffffu, bsum[t], 16);
    b2sum[t] += __shfl_xor_sync(0xffffffffu, b2sum[t], 8); b2sum[t] += __shfl_xor_sync(0xffffffffu, b2sum[t], 16);
  }
  if (lane < 8)
  {
#pragma unroll
    for (int t = 0; t < 8; ++t) { red[w*64+lane*8+t] = bsum[t]; red[(16+w)*64+lane*8+t] = b2sum[t]; }
  }
  __syncthreads();
  if (tid < 64)
  { // tid = c 8 + u 2 + ri
    const int c2 = tid>>3, u = (tid>>1)&3, ri = tid&1, j = j0+8*u+c2;
    double s1 = 0.0, s2 = 0.0;
    for (int ww = 0; ww < 16; ++ww) { s1 += red[ww*64+tid]; s2 += red[(16+ww)*64+tid]; }
    if (j < M)
    {
      const long long p = (MODEL == MODEL_RBM) ? NM+N+j : NM+j;
      base[(size_t)ri*P+p] = s1;
      if (ones) a.abs2[(size_t)blockIdx.y*3*M+2*j+ri] = s2;
    }
  }
  if (do_a && tid < N) { base[NM+tid] = asx; base[P+NM+tid] = asy; }
}

} // namespace nqs
