// C[k][c] = sum_i s_ki B[i][c] on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM) with fp64-grade
// results: the spins are +-1 (exact in int8) and B is split ERROR-FREE into 7 signed base-256 digits per column (an Ozaki
// split with one power-of-two scale per column), so every int32 accumulator is exact and so is their recombination up to one final rounding:
//
//   B[i][c] ~ Q[i][c] 2^(e_c-53),  Q = rint(B 2^(53-e_c)),  |Q| <= 2^53,  2^(e_c-1) <= max_i |B[i][c]| < 2^e_c
//   Q = sum_{s<7} d_s 256^s, d_s in [-128, 127]           (balanced digits: int8)
//   sum_i s_ki B[i][c] ~ 2^(e_c-53) sum_s 256^s (sum_i s_ki d_s[i][c])   -- 7 int8 GEMMs sharing the same A: ONE GEMM whose
//                                                                          B operand carries the 7 digit planes side by side
//
// The only rounding is Q's: |error| <= N 2^(e_c-54) <= N 2^-53 max_i|B[i][c]| per output, below the rounding error bound of the
// fp64 dot product it replaces (N 2^-53 sum_i |B[i][c]|); the recombination of the planes adds one more rounding.
// N <= 512 keeps every partial recombination exact (|acc_s| <= 128 N = 2^16).
//
// Same operands, outputs and epilogues as spin_rows_dmma_kernel (sv_struct.cuh): theta = S W + b with the log cosh row sum
// (ref: Zgemm + logcosh pass, gpu/include/impl_neural_quantum_state.cuh:78,114) and z = O v from the factors of O.
//
// Layout: one CTA per 128 chains (UMMA M = 128: chain r of the tile = TMEM lane r).  A = the spin tile [128][Npad] int8,
// K-major, no swizzle: 8 x 16 B core matrices, (r/8) SBO + (i/16) 128 + (r%8) 16 + i%16.  B is consumed in chunks of 32 real
// columns: tile [7 digit planes x 32 columns = 224 rows][Npad] in the same core-matrix order, prepared once per product by
// ozaki_split_kernel and fetched with one 1-D TMA bulk copy per chunk (ring of up to 6 tiles).  One MMA of 128 x 224 x 32 per
// 32 sites; D = 224 TMEM columns, double-buffered (448 of 512), so the MMAs of chunk c+1 run under the epilogue of chunk c.
// Epilogue: 16 warps = 4 lane quadrants x 4 column groups; a thread owns one chain and 8 real columns of the chunk:
// 7 tcgen05.ld (32x32b.x8), the planes recombined (pairs in int32, pairs of pairs in int64, one FMA), scale, then the complex
// epilogue of the DMMA kernel.
// The chunks run in lockstep (one __syncthreads each), so nothing with a global-memory latency may sit inside the chunk loop:
// scales, biases and output weights are staged in shared memory once, the visible-bias sum is taken while the first tile is
// still in flight, and T (in) / theta (out) travel through cp.async-staged shared-memory tiles, fetched chunks ahead, so that
// global memory sees whole 256-byte runs instead of one cache line per lane.
// In-kernel clocks per chunk (cfg3, Z epilogue; profiles/r2_umma.md): 4500 cycles at first -- 2000 of them LSU wavefronts of
// thread-per-chain T loads, 1500 two dependent L2 round trips for scales and biases -- and 3100 now: 1230 recombination +
// complex epilogue (fp64 pipe), 800 issuing the cp.async of the next T tile, 200 tcgen05.ld, 150 waiting for the MMAs; the
// tensor core itself needs 450.  Next: a dedicated issue warp and mbarrier hand-offs instead of the per-chunk __syncthreads.
// The kernel is launched with programmatic stream serialization: TMEM allocation, barrier setup and the spin tile overlap the
// tail of ozaki_split_kernel.
#pragma once
#include <cfloat>
#include "sv_struct.cuh"

namespace nqs
{

#define NQS_RU_THREADS 512
#define NQS_RU_NC 32                          // real columns of B per chunk
#define NQS_RU_NS 7                           // digit planes
#define NQS_RU_NB (NQS_RU_NC*NQS_RU_NS)       // rows of the B tile = UMMA N
#define NQS_RU_TMEM_COLS 512
#define NQS_RU_TBUF 256                       // TMEM column stride between the two accumulator buffers
#define NQS_RU_MAXBUF 6                       // tile buffers in flight
#define NQS_RU_TPITCH 272                     // bytes per chain in the staged tile of T / theta: 16 hidden units + 16 (LDS.128 conflict-free)
#define NQS_RU_TTILE (128*NQS_RU_TPITCH)

inline int ru_npad(const int N) { return (N+31)/32*32; }
inline size_t ru_chunk_bytes(const int N) { return (size_t)NQS_RU_NB*ru_npad(N); }
inline int ru_nchunks(const int M2) { return (M2+NQS_RU_NC-1)/NQS_RU_NC; }
// spin tile | nbuf chunk tiles | scales [chunks*32] | bias [M] | w1o [M] | 2 staged tiles of T (in) or theta (out), reused for
// the row sums [8][128] after the chunk loop | barriers
inline size_t rows_umma_smem(const int N, const int M, const int M2, const int nbuf, const int nt)
{
  return (size_t)128*ru_npad(N)+(size_t)nbuf*ru_chunk_bytes(N)+(size_t)ru_nchunks(M2)*NQS_RU_NC*sizeof(double)
        +(size_t)2*M*sizeof(cd)+(size_t)nt*NQS_RU_TTILE+128;
}
// Buffers in flight.  Both streams are consumed in lockstep with the chunks, so each must be issued far enough ahead to cover
// its latency (about two epilogues for the L2-resident digit planes, three for T from HBM): as many as fit, T first.
inline void rows_umma_plan(const int N, const int M, const int M2, const bool stream_T, const size_t smem_limit, int & nbuf, int & nt)
{
  const int nch = ru_nchunks(M2);
  nt = 2; nbuf = 2;
  if (stream_T)
    while (nt < 4 && nt < nch && rows_umma_smem(N, M, M2, nbuf, nt+1) <= smem_limit) ++nt;
  while (nbuf < NQS_RU_MAXBUF && nbuf < nch && rows_umma_smem(N, M, M2, nbuf+1, nt) <= smem_limit) ++nbuf;
  if (stream_T && nbuf < 3 && nt > 3 && rows_umma_smem(N, M, M2, 3, nt-1) <= smem_limit) { nbuf = 3; --nt; }
}

// B [N][M2] -> digit planes in the UMMA tile order + scale[c] = 2^(e_c-53) (0 for an all-zero column, NaN for a column holding
// a non-finite value: the product then propagates NaN like the fp64 GEMM).  One block per chunk of 32 columns; a thread owns one
// column and 16 consecutive sites, i.e. exactly one 16-byte row of a core matrix in each of the 7 planes: 7 vector stores.
__global__ void __launch_bounds__(1024) ozaki_split_kernel(const int N, const int M2, const double * __restrict__ B,
                                                           int8_t * __restrict__ Bq, double * __restrict__ scale, const int * __restrict__ done)
{
  // the consumer (spin_rows_umma_kernel, launched with programmatic stream serialization) may start its prologue right away;
  // it waits (griddepcontrol.wait) for this grid to complete before it touches Bq / scale
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (done != nullptr && *done) return;
  __shared__ double smax[32][32];
  const int lane = threadIdx.x&31, kc = threadIdx.x>>5, c = blockIdx.x*NQS_RU_NC+lane;     // blockDim = 32 x (npad/16)
  const int npad = (N+31)/32*32, kch = npad/16;
  double val[16];
  double m = 0.0;
#pragma unroll
  for (int t = 0; t < 16; ++t)
  {
    const int i = kc*16+t;
    val[t] = (c < M2 && i < N) ? B[(size_t)i*M2+c] : 0.0;
    const double v = fabs(val[t]);
    m = (v <= DBL_MAX) ? fmax(m, v) : INFINITY;     // NaN and Inf both land on Inf
  }
  smax[kc][lane] = m;
  __syncthreads();
  double cm = smax[0][lane];
  for (int ww = 1; ww < kch; ++ww) cm = fmax(cm, smax[ww][lane]);
  if (c >= M2) return;
  const bool usable = (cm > 0.0 && cm <= DBL_MAX);
  const int e = usable ? ilogb(cm)+1 : 0;
  if (kc == 0) scale[c] = usable ? scalbn(1.0, e-53) : (cm == 0.0 ? 0.0 : nan(""));
  // 2^(53-e) as a bit pattern when it is a normal number (always, short of columns below 2^-960 or above 2^1020)
  const int be = 1023+53-e;
  const bool direct = (be >= 1 && be <= 2046);
  const double up = direct ? __hiloint2double(be<<20, 0) : 0.0;
  uint32_t pk[NQS_RU_NS][4];
#pragma unroll
  for (int s = 0; s < NQS_RU_NS; ++s) { pk[s][0] = 0u; pk[s][1] = 0u; pk[s][2] = 0u; pk[s][3] = 0u; }
#pragma unroll
  for (int t = 0; t < 16; ++t)
  {
    long long Q = !usable ? 0ll : (direct ? __double2ll_rn(val[t]*up) : llrint(scalbn(val[t], 53-e)));
#pragma unroll
    for (int s = 0; s < NQS_RU_NS; ++s)
    {
      const int d = (int)((Q+128)&255)-128;
      Q = (Q-d)>>8;
      pk[s][t>>2] |= (uint32_t)(d&255)<<(8*(t&3));
    }
  }
  int8_t * tile = Bq+(size_t)blockIdx.x*NQS_RU_NB*npad;
#pragma unroll
  for (int s = 0; s < NQS_RU_NS; ++s)
  {
    const int n = s*NQS_RU_NC+lane;
    *reinterpret_cast<uint4*>(tile+(size_t)(n>>3)*(kch*128)+kc*128+(n&7)*16) = make_uint4(pk[s][0], pk[s][1], pk[s][2], pk[s][3]);
  }
}

// bounded mbarrier wait: a descriptor mistake must end in a trap, not in a hung GPU
__device__ __forceinline__ void ru_mbar_wait(uint64_t * bar, const uint32_t parity)
{
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  const long long t0 = clock64();
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(addr), "r"(parity), "r"(20000u) : "memory");
    if (!ok && clock64()-t0 > 4000000000ll) __trap();
  } while (!ok);
}
// shared-memory matrix descriptor (sm_100 version 1), no swizzle: start address, leading (K) and stride (M/N) byte offsets / 16
__device__ __forceinline__ uint64_t ru_desc(const uint32_t saddr, const uint32_t lbo, const uint32_t sbo)
{
  return (uint64_t)((saddr&0x3FFFFu)>>4)|((uint64_t)(lbo>>4)<<16)|((uint64_t)(sbo>>4)<<32)|(1ull<<46);
}
__device__ __forceinline__ void ru_mma_i8(const uint32_t tmem_d, const uint64_t adesc, const uint64_t bdesc, const uint32_t idesc, const uint32_t accumulate)
{
  const uint32_t zero = 0u;
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
    :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(zero) : "memory");
}
__device__ __forceinline__ void ru_commit(uint64_t * bar)
{
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ru_tmem_ld8(const uint32_t taddr, uint32_t (&v)[8])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}

template <int MODEL, int EPI>
__global__ void __launch_bounds__(NQS_RU_THREADS, 1) spin_rows_umma_kernel(const RowsArgs a, const int8_t * __restrict__ Bq, const double * __restrict__ scale, const int nbuf, const int nt, const int trace_cta)
{
  if (EPI == ROWS_EPI_Z && a.done != nullptr && *a.done) return;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int NC = NQS_RU_NC, NS = NQS_RU_NS, NB = NQS_RU_NB;
  constexpr bool ZF = (EPI == ROWS_EPI_Z);
  const int N = a.N, M = a.M, M2 = (EPI == ROWS_EPI_SJS) ? N : 2*M, npad = (N+31)/32*32, kch = npad/16, nch = (M2+NC-1)/NC;
  const uint32_t sbo = (uint32_t)kch*128u, chunk_bytes = (uint32_t)NB*(uint32_t)npad;
  unsigned char * As = smem_raw;                                  // [128][npad] spins, core-matrix order
  unsigned char * Bs = As+(size_t)128*npad;                       // [nbuf] chunk tiles
  double * scs = reinterpret_cast<double*>(Bs+(size_t)nbuf*chunk_bytes);     // [nch*NC] column scales (0 beyond M2)
  cd * bias_s = reinterpret_cast<cd*>(scs+(size_t)nch*NC);        // [M]
  cd * w1o_s = bias_s+M;                                          // [M] (FFNN)
  unsigned char * Ts = reinterpret_cast<unsigned char*>(w1o_s+M); // [nt][128][NQS_RU_TPITCH]: T of a chunk (Z; nt-1 chunks ahead) / theta of a chunk (THETA, LNPSI; 2)
  cd * red = reinterpret_cast<cd*>(Ts);                           // [2][4 column groups][128 chains] row sums, visible-bias sums: after the loop
  uint64_t * bfull = reinterpret_cast<uint64_t*>(Ts+(size_t)nt*NQS_RU_TTILE);   // [NQS_RU_MAXBUF] chunk tile landed
  uint64_t * mdone = bfull+NQS_RU_MAXBUF;                         // [2] MMAs of the chunk complete
  uint32_t * tptr = reinterpret_cast<uint32_t*>(mdone+2);
  const int tid = threadIdx.x, lane = tid&31, w = tid>>5, q = w&3, g = w>>2, row = 32*q+lane;
  const long long kbase = (long long)blockIdx.x*128, k = kbase+row;
  // Built with -DNQS_RU_TRACE_BUILD and run with NQS_RU_TRACE=<cta>: thread 96 of that CTA prints the clocks of the prologue and
  // of the phases of chunks 4..7 (how the numbers in the header were measured; off by default: the stamps cost a stack frame)
#ifdef NQS_RU_TRACE_BUILD
  const bool trace = (trace_cta >= 0 && (int)blockIdx.x == trace_cta && tid == 96);
  long long ts[36];
  int nts = 0;
#define NQS_RU_STAMP() do { if (trace && nts < 36) ts[nts++] = clock64(); } while (0)
#else
#define NQS_RU_STAMP() do { } while (0)
#endif
  NQS_RU_STAMP();

  if (w == 0)
  {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tptr)), "r"((uint32_t)NQS_RU_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32)
  {
    for (int b = 0; b < NQS_RU_MAXBUF; ++b) mbar_init(bfull+b, 1);
    mbar_init(mdone, 1); mbar_init(mdone+1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // T of a chunk ([128 chains][16 hidden units]) / theta of a chunk travel through shared memory so that global memory sees
  // 256-byte runs per chain (a thread-per-chain access touches 32 cache lines per warp instruction: measured 2000 cycles per
  // chunk of LSU time).  idx -> (chain, 16-byte piece): 16 consecutive lanes cover one chain's 256 bytes.
  int tb_next = 0, tb_cur = 0;          // staging buffers of the next chunk to fetch / of the chunk being consumed (c % nt)
  auto stage_T = [&](const int c)       // always commits a group (empty beyond the last chunk): the waits count groups
  {
    if (c < nch)
    {
#pragma unroll
      for (int t = 0; t < 4; ++t)
      {
        const int idx = tid+NQS_RU_THREADS*t, r = idx>>4, j = c*(NC/2)+(idx&15);
        const bool ok = (kbase+r < a.K && j < M);
        cp_async16(Ts+(size_t)tb_next*NQS_RU_TTILE+r*NQS_RU_TPITCH+(idx&15)*16, ok ? a.T+(kbase+r)*M+j : a.T, ok ? 16 : 0);
      }
      tb_next = (tb_next+1 == nt) ? 0 : tb_next+1;
    }
    cp_async_commit();
  };
  auto wait_T = [&]()                   // all but the newest nt-2 groups have landed
  {
    if (nt == 2) cp_async_wait<0>(); else if (nt == 3) cp_async_wait<1>(); else cp_async_wait<2>();
  };
  auto flush_theta = [&](const int c)
  {
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
      const int idx = tid+NQS_RU_THREADS*t, r = idx>>4, j = c*(NC/2)+(idx&15);
      if (kbase+r < a.K && j < M)
        a.theta[(kbase+r)*M+j] = *reinterpret_cast<const cd*>(Ts+(size_t)(c&1)*NQS_RU_TTILE+r*NQS_RU_TPITCH+(idx&15)*16);
    }
  };
  const bool store_theta = (!ZF && EPI != ROWS_EPI_SJS && a.theta != nullptr);
  if (ZF)
    for (int c = 0; c < nt-1; ++c) stage_T(c);
  { // spin tile: 16-byte pieces, consecutive threads along a chain's row
    const bool vec = (N%16 == 0) && ((reinterpret_cast<size_t>(a.spins)&15) == 0);
    for (int idx = tid; idx < 128*kch; idx += NQS_RU_THREADS)
    {
      const int r = idx/kch, kc = idx-r*kch;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (kbase+r < a.K)
      {
        const int8_t * src = a.spins+(size_t)(kbase+r)*N+kc*16;
        if (vec) { if (kc*16 < N) val = *reinterpret_cast<const uint4*>(src); }
        else
        {
          unsigned char b[16];
#pragma unroll
          for (int t = 0; t < 16; ++t) b[t] = (kc*16+t < N) ? (unsigned char)src[t] : (unsigned char)0;
          val.x = b[0]|(b[1]<<8)|(b[2]<<16)|((uint32_t)b[3]<<24);   val.y = b[4]|(b[5]<<8)|(b[6]<<16)|((uint32_t)b[7]<<24);
          val.z = b[8]|(b[9]<<8)|(b[10]<<16)|((uint32_t)b[11]<<24); val.w = b[12]|(b[13]<<8)|(b[14]<<16)|((uint32_t)b[15]<<24);
        }
      }
      *reinterpret_cast<uint4*>(As+(size_t)(r>>3)*sbo+kc*128+(r&7)*16) = val;
    }
  }
  if (EPI != ROWS_EPI_SJS)
    for (int j = tid; j < M; j += NQS_RU_THREADS)
    {
      bias_s[j] = a.bias[j];
      if (MODEL == MODEL_FFNN && EPI != ROWS_EPI_THETA) w1o_s[j] = a.w1o[j];
    }
  // everything above overlapped the tail of ozaki_split_kernel (programmatic dependent launch); its digit planes and scales are
  // read from here on
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int c = tid; c < nch*NC; c += NQS_RU_THREADS) scs[c] = (c < M2) ? scale[c] : 0.0;
  if (ZF) wait_T();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes of A before the tensor core reads them
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = *tptr;
  NQS_RU_STAMP();
  // instruction descriptor: D = s32, A = B = s8, both K-major, N = 224, M = 128
  constexpr uint32_t idesc = (2u<<4)|(1u<<7)|(1u<<10)|((uint32_t)(NB>>3)<<17)|((128u>>4)<<24);
  const uint64_t adesc = ru_desc(smem_u32(As), 128u, sbo);

  auto issue_tma = [&](const int c)
  {
    const int b = c%nbuf;
    mbar_expect_tx(bfull+b, chunk_bytes);
    tma_load_1d(Bs+(size_t)b*chunk_bytes, Bq+(size_t)c*chunk_bytes, chunk_bytes, bfull+b);
  };
  auto issue_mma = [&](const int c)
  {
    const int b = c%nbuf;
    ru_mbar_wait(bfull+b, (uint32_t)((c/nbuf)&1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint64_t bdesc = ru_desc(smem_u32(Bs+(size_t)b*chunk_bytes), 128u, sbo);
    for (int kk = 0; kk < npad/32; ++kk)     // 32 sites = two 16-byte K chunks = 256 bytes further along both operands
      ru_mma_i8(tbase+(uint32_t)(c&1)*NQS_RU_TBUF, adesc+(uint64_t)(kk*16), bdesc+(uint64_t)(kk*16), idesc, kk > 0 ? 1u : 0u);
    ru_commit(mdone+(c&1));
  };
  if (tid == 0)
    for (int c = 0; c < nbuf && c < nch; ++c) issue_tma(c);
  // while the first tiles are in flight: the factors of chunk 0 and the visible-bias term (RBM; every column group takes a
  // quarter of the sites)
  const unsigned char * arow = As+(size_t)(row>>3)*sbo+(row&7)*16;     // this chain's spins: site i at arow[(i/16) 128 + i%16]
  const bool live = (k < a.K);
  cd Ln[4];
  auto fetch_L = [&](const int c)       // FFNN: log cosh theta of the next chunk, thread-per-chain (no room to stage it as well)
  {
    const int j0 = (c*NC+g*8)>>1;
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (live && j0+t < M) Ln[t] = a.L[k*M+j0+t];
  };
  if (ZF && MODEL == MODEL_FFNN) fetch_L(0);
  cd sv = cmake(0.0, 0.0);
  if (MODEL == MODEL_RBM && EPI != ROWS_EPI_SJS && live)
  {
    const int per = (N+3)/4, i1 = (g+1)*per < N ? (g+1)*per : N;
    for (int i = g*per; i < i1; ++i)
    {
      const double s = ZF ? (double)(int8_t)arow[(i>>4)*128+(i&15)] : (double)a.sa_spins[k*N+i];
      const cd ai = a.avis[i];
      sv.x = fma(s, ai.x, sv.x); sv.y = fma(s, ai.y, sv.y);
    }
  }
  if (tid == 0) issue_mma(0);
  __syncwarp();
  NQS_RU_STAMP();
  cd rsum = cmake(0.0, 0.0);
  for (int c = 0; c < nch; ++c)
  {
    // here: MMAs of chunk c issued, later tiles on their way; the accumulator buffer (c+1)&1 was drained before the
    // __syncthreads that ended the previous iteration
    if (tid == 0 && c+1 < nch) issue_mma(c+1);
    __syncwarp();
    if (c >= 4 && c < 8) NQS_RU_STAMP();
    cd Lv[4];
    if (ZF)
    {
      if (MODEL == MODEL_FFNN)
      {
#pragma unroll
        for (int t = 0; t < 4; ++t) Lv[t] = Ln[t];
        if (c+1 < nch) fetch_L(c+1);
      }
      stage_T(c+nt-1);                         // nt-1 chunks ahead, into the buffer chunk c-1 was read from
    }
    else if (store_theta && c > 0) flush_theta(c-1);
    if (c >= 4 && c < 8) NQS_RU_STAMP();
    ru_mbar_wait(mdone+(c&1), (uint32_t)((c>>1)&1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (c >= 4 && c < 8) NQS_RU_STAMP();
    if (tid == 0 && c+nbuf < nch) issue_tma(c+nbuf);     // the tile buffer of chunk c is free: its MMAs completed
    __syncwarp();
    uint32_t v[NS][8];
    const uint32_t taddr = tbase+((uint32_t)(32*q)<<16)+(uint32_t)(c&1)*NQS_RU_TBUF+(uint32_t)(g*8);
#pragma unroll
    for (int s = 0; s < NS; ++s) ru_tmem_ld8(taddr+(uint32_t)(s*NC), v[s]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (c >= 4 && c < 8) NQS_RU_STAMP();
    const int col0 = c*NC+g*8;
    // sum_s 256^s acc_s with as little fp64-pipe work as possible (the conversions and FMAs of this recombination were the
    // longest phase of a chunk): planes are paired in int32 (|acc| <= 128 N <= 2^16: acc_s + 256 acc_{s+1} < 2^25), pairs in
    // int64 (lo = planes 0..3 < 2^41, hi = planes 4..6 < 2^33), both converted by the 2^52+2^51 bias trick (an integer add
    // and one DADD instead of I2F), then hi 2^32 + lo in one FMA: a single rounding.
    double x[8];
#pragma unroll
    for (int cc = 0; cc < 8; ++cc)
    {
      const int p0 = (int)v[0][cc]+256*(int)v[1][cc], p1 = (int)v[2][cc]+256*(int)v[3][cc], p2 = (int)v[4][cc]+256*(int)v[5][cc];
      const long long lo = (long long)p1*65536+(long long)p0, hi = (long long)(int)v[6][cc]*65536+(long long)p2;
      const double dlo = __longlong_as_double(0x4338000000000000ll+lo)-6755399441055744.0;
      const double dhi = __longlong_as_double(0x4338000000000000ll+hi)-6755399441055744.0;
      x[cc] = fma(dhi, 4294967296.0, dlo)*scs[col0+cc];
    }
    if (EPI == ROWS_EPI_SJS)
    { // x = (S J)[k][col]: dot with the chain's own spins (0 in the padding)
#pragma unroll
      for (int cc = 0; cc < 8; ++cc)
      {
        const int col = col0+cc;
        if (col < N) rsum.x = fma(x[cc], (double)(int8_t)arow[(col>>4)*128+(col&15)], rsum.x);
      }
    }
    else if (live)
    {
#pragma unroll
      for (int cc = 0; cc < 8; cc += 2)
      {
        const int j = (col0+cc)>>1;
        if (j >= M) continue;
        const cd bj = bias_s[j];
        cd wj = cmake(1.0, 0.0);
        if (MODEL == MODEL_FFNN && EPI != ROWS_EPI_THETA) wj = w1o_s[j];
        const cd val = cmake(x[cc]+bj.x, x[cc+1]+bj.y);
        if (ZF)
        {
          const cd Tkj = *reinterpret_cast<const cd*>(Ts+(size_t)tb_cur*NQS_RU_TTILE+row*NQS_RU_TPITCH+(g*4+(cc>>1))*16);
          rsum.x = fma(Tkj.x, val.x, rsum.x); rsum.x = fma(-Tkj.y, val.y, rsum.x);
          rsum.y = fma(Tkj.x, val.y, rsum.y); rsum.y = fma(Tkj.y, val.x, rsum.y);
          if (MODEL == MODEL_FFNN) rsum = cadd(rsum, cmul(Lv[cc>>1], wj));
        }
        else
        {
          if (store_theta) *reinterpret_cast<cd*>(Ts+(size_t)(c&1)*NQS_RU_TTILE+row*NQS_RU_TPITCH+(g*4+(cc>>1))*16) = val;
          if (EPI == ROWS_EPI_LNPSI)
          {
            const cd lc = c_logcosh(val);
            rsum = cadd(rsum, (MODEL == MODEL_RBM) ? lc : cmul(wj, lc));
          }
        }
      }
    }
    if (ZF) tb_cur = (tb_cur+1 == nt) ? 0 : tb_cur+1;
    if (c >= 4 && c < 8) NQS_RU_STAMP();
    if (ZF) wait_T();                          // T of chunk c+1 has landed (this thread's pieces; the barrier covers the rest)
    if (c >= 4 && c < 8) NQS_RU_STAMP();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (c >= 3 && c < 8) NQS_RU_STAMP();
  }
  NQS_RU_STAMP();
  if (store_theta) { flush_theta(nch-1); __syncthreads(); }     // red[] lives in the staging tiles
  red[g*128+row] = rsum;
  red[(4+g)*128+row] = sv;
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tbase), "r"((uint32_t)NQS_RU_TMEM_COLS) : "memory");
  if (tid < 128 && live)
  { // (tid < 128 <=> g == 0, row == tid) the final value of the chain, column groups in fixed order
    cd svt = cmake(0.0, 0.0), tot = cmake(0.0, 0.0);
    for (int gg = 0; gg < 4; ++gg) { svt = cadd(svt, red[(4+gg)*128+row]); tot = cadd(tot, red[gg*128+row]); }
    tot = cadd(tot, svt);
    if (ZF) a.zk[k] = tot;
    else if (EPI == ROWS_EPI_SJS) a.sjs[k] = tot.x;
    else
    {
      if (a.sa) a.sa[k] = svt;
      if (EPI == ROWS_EPI_LNPSI) a.lnpsi[k] = tot;
    }
  }
#ifdef NQS_RU_TRACE_BUILD
  if (trace)
  {
    NQS_RU_STAMP();
    printf("ru trace EPI %d nbuf %d nt %d | prologue %lld first-issue %lld | chunk 3 ends %lld |", EPI, nbuf, nt, ts[1]-ts[0], ts[2]-ts[1], ts[3]-ts[2]);
    for (int i = 4; i+6 < nts-1; i += 7)
      printf(" [issue %lld stage %lld mma-wait %lld ldtm %lld epi %lld T-wait %lld sync %lld]", ts[i]-ts[i-1], ts[i+1]-ts[i], ts[i+2]-ts[i+1], ts[i+3]-ts[i+2],
             ts[i+4]-ts[i+3], ts[i+5]-ts[i+4], ts[i+6]-ts[i+5]);
    printf(" | rest of loop %lld tail %lld\n", ts[nts-2]-ts[nts-3], ts[nts-1]-ts[nts-2]);
  }
#endif
#undef NQS_RU_STAMP
}

} // namespace nqs
