// fp64 tensor-core (DMMA) kernels for the GEMM-shaped pieces of the path.  Every row of O is an outer product of the spin
// row and the hidden-unit row plus two short blocks (ref RBM__GetGradientsOfParameters__, impl_neural_quantum_state.cuh:1426-1449;
// FFNN :1622-1690),
//   RBM   O_k = [ s_ki T_kj (i*M+j) | s_ki | T_kj ],   T = tanh(theta)
//   FFNN  O_k = [ s_ki T_kj (j*N+i) | T_kj | L_kj ],   T = tanh(theta) w1o,  L = logcosh(theta)
// so three products are real-by-complex GEMMs with the +-1 spin matrix S [K][N] as one operand:
//   theta  = S W + b                                              (ref RBM::initialize / forward, Zgemm at :78,:114)
//   z_k    = (O v)_k   = sum_j T_kj ((S V)_kj + vh_j) + ...       (ref SMatrixForCG::dot, Zgemm 1xKxP, functor_for_CG.cuh:110)
//   traw   = O^H z     : W block = S^T C,  C_kj = conj(T_kj) z_k  (ref Zgemv PxK, functor_for_CG.cuh:121)
// A complex [.][M] operand is a real [.][2M] operand (re, im interleaved), and the two doubles a lane holds of a DMMA
// accumulator tile are exactly (re, im) of one complex element.  Both kernels issue mma.sync.aligned.m8n8k4 f64 (SASS
// DMMA.8x8x4; measured 37.1 TFLOP/s on B200 = the fp64 peak, scripts/micro/dmma_bench.cu) with operands staged in shared
// memory at a row pitch of 4 (mod 16) doubles, which makes every fragment load a 2-wavefront (minimal) LDS.64.
//
// spin_rows_dmma_kernel   m = chains, n = 2M, k = sites:   theta (+ fused lnpsi = sum_j logcosh theta + a.s epilogue), or z
// spin_cols_dmma_kernel   m = sites,  n = 2M, k = chains:  split over chain chunks, partials folded in fixed order by
//                         cg_fused_kernel exactly like the cluster partials of the one-pass kernel
// With NQS_FLAG_STRUCTURED_SV the CG uses the last two instead of streaming O: 2 x 2.1 GFLOP instead of 8.7 GB per S*v at
// N=128, M=256, K=16384, and O is never written.  The explicit-O path (sv_fused.cuh) stays the default: it is the
// formulation the reference and the headline roofline are stated on.
#pragma once
#include "device_math.cuh"
#include "sampler_kernels.cuh"

namespace nqs
{
__device__ __forceinline__ void dmma884(double & c0, double & c1, const double a, const double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// T (and, FFNN, L) of the current chain state: one pass over theta, once per SR step
template <int MODEL>
__global__ void hidden_values_kernel(const int N, const int M, const long long K, const cd * params, const cd * __restrict__ theta,
  cd * __restrict__ T, cd * __restrict__ L)
{
  const ModelPtrs mp = model_ptrs(MODEL, params, N, M);
  const long long total = K*(long long)M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x)
  {
    const cd th = theta[idx];
    cd t = c_tanh(th);
    if (MODEL == MODEL_FFNN)
    {
      const int j = (int)(idx%M);
      t = cmul(t, mp.w1o[j]);
      L[idx] = c_logcosh(th);
    }
    T[idx] = t;
  }
}

__global__ void spins_to_double_kernel(const long long n, const int8_t * __restrict__ spins, double * __restrict__ out)
{
  for (long long i = (long long)blockIdx.x*blockDim.x+threadIdx.x; i < n; i += (long long)gridDim.x*blockDim.x) out[i] = (double)spins[i];
}

// FFNN keeps the W block of O / v transposed (j*N+i): bring it to the natural i*M+j layout once per product
__global__ void transpose_wblock_kernel(const int N, const int M, const cd * __restrict__ v, cd * __restrict__ out, const int * __restrict__ done)
{
  if (done != nullptr && *done) return;
  const long long total = (long long)N*M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x)
  {
    const int i = (int)(idx/M), j = (int)(idx-(long long)i*M);
    out[idx] = v[(size_t)j*N+i];
  }
}

// ---- rows: C[k][c] = sum_i s_ki B[i][c] ------------------------------------------------------------------------------------
enum { ROWS_EPI_THETA = 0, ROWS_EPI_LNPSI = 1, ROWS_EPI_Z = 2, ROWS_EPI_SJS = 3 };
#define NQS_DR_THREADS 256
#define NQS_DR_KC 32        // sites per staged slab of B
#define NQS_DR_NTW 4        // n-tiles (8 real columns each) per warp
#define NQS_DR_CW (8*NQS_DR_NTW*8)   // 256 real = 128 complex columns per pass

struct RowsArgs
{
  int N, M;
  long long K;
  const int8_t * spins;      // [K][N]
  const double * B;          // [N][2M]: W (theta) or the W block of v in natural layout (z)
  const cd * bias;           // [M]: b (theta) or the hidden-unit block of v that multiplies T (z)
  // theta / lnpsi
  cd * theta;                // [K][M] out (may be null)
  const int8_t * sa_spins;   // RBM visible-bias term taken from these spins (ref forward(spins, lnpsi, false) quirk)
  const cd * avis;           // [N] RBM a (theta: sa; z: the a block of v)
  const cd * w1o;            // [M] FFNN output weights (lnpsi) / the w1o block of v (z)
  cd * sa;                   // [K] out (may be null)
  cd * lnpsi;                // [K] out (LNPSI)
  // z
  const cd * T;              // [K][M]
  const cd * L;              // [K][M] FFNN
  cd * zk;                   // [K] out
  const int * done;
  // s.J.s (ROWS_EPI_SJS): B = J [N][N] real, N even; out sjs[k] = sum_ij s_ki J_ij s_kj (ref c5 Zgemm + k10, impl_hamiltonians.cuh:226-231)
  double * sjs;
};

inline size_t rows_dmma_smem(int N, int MT)
{
  const int n4 = (N+3)/4*4;
  return ((size_t)8*MT*(n4+4)+(size_t)2*NQS_DR_KC*(NQS_DR_CW+4))*sizeof(double)+(size_t)(NQS_DR_THREADS/32)*8*MT*sizeof(cd);
}

// 16-byte asynchronous copy global -> shared; src_bytes = 0 zero-fills (out-of-range rows / columns of the slab)
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc, const int src_bytes)
{
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N_) : "memory"); }

// One CTA per SM: 8 MT chains x all columns; the slabs of B ([32 sites][256 columns], 66 KB) are double-buffered with cp.async so
// the L2 stream of B overlaps the DMMAs.  B is re-streamed by every CTA, so fewer, taller CTAs (MT = 8: 64 chains) halve the
// L2 traffic of MT = 4; MT = 2 / 4 keep every SM busy when a rank holds few chains.
template <int MODEL, int EPI, int MT>
__global__ void __launch_bounds__(NQS_DR_THREADS, 1) spin_rows_dmma_kernel(const RowsArgs a)
{
  if (EPI == ROWS_EPI_Z && a.done != nullptr && *a.done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int RT = 8*MT, NTW = NQS_DR_NTW, CW = NQS_DR_CW, KC = NQS_DR_KC, PB = CW+4;
  const int N = a.N, M = a.M, n4 = (N+3)/4*4, PA = n4+4, M2 = (EPI == ROWS_EPI_SJS) ? N : 2*M;
  double * As = reinterpret_cast<double*>(smem_raw);            // [RT][PA] spins of the CTA's chains (0 for padding)
  double * Bs = As+(size_t)RT*PA;                               // [2][KC][PB] slabs of B
  cd * red = reinterpret_cast<cd*>(Bs+(size_t)2*KC*PB);         // [warps][RT]
  const int tid = threadIdx.x, lane = tid&31, w = tid>>5, g = lane>>2, t = lane&3;
  const long long kbase = (long long)blockIdx.x*RT;
  const int nslab = (n4+KC-1)/KC, npass = (M2+CW-1)/CW, nstage = nslab*npass;
  // slab q = (column pass q / nslab, site slab q % nslab) -> buffer q & 1
  auto issue = [&](const int q)
  {
    const int c0 = (q/nslab)*CW, i0 = (q%nslab)*KC;
    double * dst = Bs+(size_t)(q&1)*KC*PB;
    for (int idx = tid; idx < KC*(CW/2); idx += NQS_DR_THREADS)
    {
      const int r = idx/(CW/2), c = 2*(idx-r*(CW/2));
      const bool okb = (i0+r < N && c0+c < M2);
      cp_async16(dst+r*PB+c, okb ? a.B+(size_t)(i0+r)*M2+c0+c : a.B, okb ? 16 : 0);
    }
    cp_async_commit();
  };
  issue(0);
  for (int idx = tid; idx < RT*n4; idx += NQS_DR_THREADS)
  {
    const int r = idx/n4, i = idx-r*n4;
    As[r*PA+i] = (kbase+r < a.K && i < N) ? (double)a.spins[(kbase+r)*N+i] : 0.0;
  }
  cd rsum[MT];
#pragma unroll
  for (int m = 0; m < MT; ++m) rsum[m] = cmake(0.0, 0.0);

  for (int pass = 0; pass < npass; ++pass)
  {
    const int c0 = pass*CW;
    double acc[MT][NTW][2];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int n = 0; n < NTW; ++n) { acc[m][n][0] = 0.0; acc[m][n][1] = 0.0; }
    for (int sl = 0; sl < nslab; ++sl)
    {
      const int q = pass*nslab+sl, i0 = sl*KC;
      if (q+1 < nstage) { issue(q+1); cp_async_wait<1>(); } else cp_async_wait<0>();
      __syncthreads();
      const int ksteps = ((n4-i0 < KC) ? n4-i0 : KC)/4;
      const double * ap = As+g*PA+i0+t;
      const double * bp = Bs+(size_t)(q&1)*KC*PB+t*PB+w*NTW*8+g;
      for (int ks = 0; ks < ksteps; ++ks)
      {
        double af[MT], bf[NTW];
#pragma unroll
        for (int m = 0; m < MT; ++m) af[m] = ap[m*8*PA+ks*4];
#pragma unroll
        for (int n = 0; n < NTW; ++n) bf[n] = bp[ks*4*PB+n*8];
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
          for (int n = 0; n < NTW; ++n) dmma884(acc[m][n][0], acc[m][n][1], af[m], bf[n]);
      }
      __syncthreads();     // buffer q & 1 is refilled by the issue of the next iteration
    }
    // epilogue of this column pass: lane holds (re, im) of chain kbase + 8m + g, hidden unit (c0 + (w NTW + n) 8)/2 + t
    if (EPI == ROWS_EPI_SJS)
    { // lane holds (S J)[k][c], c = c0 + (w NTW + n) 8 + 2t + {0, 1}: dot with the chain's own spins
#pragma unroll
      for (int n = 0; n < NTW; ++n)
      {
        const int c = c0+(w*NTW+n)*8+2*t;
        if (c >= N) continue;
#pragma unroll
        for (int m = 0; m < MT; ++m)
        {
          const double * srow = As+(8*m+g)*PA+c;
          rsum[m].x = fma(acc[m][n][0], srow[0], rsum[m].x);
          rsum[m].x = fma(acc[m][n][1], srow[1], rsum[m].x);
        }
      }
    }
    else
#pragma unroll
    for (int n = 0; n < NTW; ++n)
    {
      const int j = (c0+(w*NTW+n)*8)/2+t;
      if (j >= M) continue;
      const cd bj = a.bias[j];
      cd wj = cmake(1.0, 0.0);
      if (MODEL == MODEL_FFNN && EPI != ROWS_EPI_THETA) wj = a.w1o[j];
#pragma unroll
      for (int m = 0; m < MT; ++m)
      {
        const long long k = kbase+8*m+g;
        if (k >= a.K) continue;
        const cd val = cmake(acc[m][n][0]+bj.x, acc[m][n][1]+bj.y);
        if (EPI == ROWS_EPI_Z)
        {
          cd term = cmul(a.T[k*M+j], val);
          if (MODEL == MODEL_FFNN) term = cadd(term, cmul(a.L[k*M+j], wj));
          rsum[m] = cadd(rsum[m], term);
        }
        else
        {
          if (a.theta) a.theta[k*M+j] = val;
          if (EPI == ROWS_EPI_LNPSI)
          {
            const cd lc = c_logcosh(val);
            rsum[m] = cadd(rsum[m], (MODEL == MODEL_RBM) ? lc : cmul(wj, lc));
          }
        }
      }
    }
  }
  // row sums: over the 4 lanes of a quad, then over the warps (fixed order)
  if (EPI != ROWS_EPI_THETA)
  {
#pragma unroll
    for (int m = 0; m < MT; ++m)
    {
      cd s = rsum[m];
      s.x += __shfl_xor_sync(0xffffffffu, s.x, 1); s.y += __shfl_xor_sync(0xffffffffu, s.y, 1);
      s.x += __shfl_xor_sync(0xffffffffu, s.x, 2); s.y += __shfl_xor_sync(0xffffffffu, s.y, 2);
      if (t == 0) red[w*RT+8*m+g] = s;
    }
  }
  __syncthreads();
  // visible-bias term (RBM) and the final value of chain kbase + r: 8 lanes per chain
  for (int r = tid>>3; r < RT; r += NQS_DR_THREADS/8)
  {
    const int q = tid&7;
    const long long k = kbase+r;
    cd sv = cmake(0.0, 0.0);
    if (MODEL == MODEL_RBM && EPI != ROWS_EPI_SJS && k < a.K)
    {
      for (int i = q; i < N; i += 8)
      {
        const double s = (EPI == ROWS_EPI_Z) ? As[r*PA+i] : (double)a.sa_spins[k*N+i];
        const cd ai = a.avis[i];
        sv.x = fma(s, ai.x, sv.x); sv.y = fma(s, ai.y, sv.y);
      }
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) { sv.x += __shfl_xor_sync(0xffffffffu, sv.x, o); sv.y += __shfl_xor_sync(0xffffffffu, sv.y, o); }
    if (q == 0 && k < a.K)
    {
      cd tot = sv;
      if (EPI != ROWS_EPI_THETA)
        for (int ww = 0; ww < NQS_DR_THREADS/32; ++ww) tot = cadd(tot, red[ww*RT+r]);
      if (EPI == ROWS_EPI_Z) a.zk[k] = tot;
      else if (EPI == ROWS_EPI_SJS) a.sjs[k] = tot.x;
      else
      {
        if (a.sa) a.sa[k] = sv;
        if (EPI == ROWS_EPI_LNPSI) a.lnpsi[k] = tot;
      }
    }
  }
}

// ---- cols: part[chunk][{re,im}][p] = sum_{k in chunk} conj(O_kp) z_k from the factors ---------------------------------------
#define NQS_DC_THREADS 256
#define NQS_DC_KC 32        // chains per stage

struct ColsArgs
{
  int N, M;
  long long K, P;
  const int8_t * spins;
  const cd * T;
  const cd * L;              // FFNN
  const cd * zk;
  double * part;
  long long rows_per_chunk;  // multiple of NQS_DC_KC
  const int * done;
  // SR setup (zmode = 1, gridDim.z = 2): slice z = 0 takes z_k = 1 (-> conj(sum_k O_kp)) and also sums |T|^2 (|L|^2) per hidden
  // unit into abs2[chunk][3M]; slice z = 1 takes zk (= htilda -> conj(sum_k O_kp conj(h_k))).  Slice z writes part + z*part_stride.
  int zmode;
  long long part_stride;
  double * abs2;
};

// warp grid WM (sites) x 8/WM (columns); a warp owns MTW x NTW accumulator tiles.  nsc = sites covered, cw = real columns per CTA
inline int cols_dmma_nsc(int mtw, int wm) { return 8*mtw*wm; }
inline int cols_dmma_cw(int ntw, int wm) { return (8/wm)*ntw*8; }
inline size_t cols_dmma_smem(int nsc, int cw)
{
  return (size_t)2*NQS_DC_KC*((nsc+4)+(cw+4))*sizeof(double)+(size_t)2*NQS_DC_KC*sizeof(cd);
}

template <int MODEL, int MTW, int NTW, int WM>
__global__ void __launch_bounds__(NQS_DC_THREADS, 1) spin_cols_dmma_kernel(const ColsArgs a)
{
  if (a.done != nullptr && *a.done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int KC = NQS_DC_KC;
  constexpr int WN = 8/WM, CW = WN*NTW*8, PC = CW+4, CH = CW/2, ns = 8*MTW*WM, PS = ns+4;
  static_assert(ns <= 256 && CW <= 256, "tile too large for the producer mapping");
  const int N = a.N, M = a.M;
  double * Ss = reinterpret_cast<double*>(smem_raw);             // [2][KC][PS] spins of the stage's chains
  double * Cs = Ss+(size_t)2*KC*PS;                              // [2][KC][PC] conj(T) z
  cd * zs = reinterpret_cast<cd*>(Cs+(size_t)2*KC*PC);           // [2][KC]
  const int tid = threadIdx.x, lane = tid&31, w = tid>>5, g = lane>>2, t = lane&3;
  const int wm = w%WM, wn = w/WM;
  const int j0 = blockIdx.x*CH;                                  // first hidden unit of this column group
  const long long k0 = (long long)blockIdx.y*a.rows_per_chunk;
  const long long k1 = (k0+a.rows_per_chunk < a.K) ? k0+a.rows_per_chunk : a.K;
  const int nstages = (int)((k1-k0+KC-1)/KC);
  const bool do_a = (MODEL == MODEL_RBM && blockIdx.x == 0);
  const bool ones = (a.zmode == 1 && blockIdx.z == 0);

  double acc[MTW][NTW][2];
#pragma unroll
  for (int m = 0; m < MTW; ++m)
#pragma unroll
    for (int n = 0; n < NTW; ++n) { acc[m][n][0] = 0.0; acc[m][n][1] = 0.0; }
  double bsum = 0.0, asx = 0.0, asy = 0.0, b2sum = 0.0;

  // producer mapping: thread -> chain kk = tid / 8 of the stage, hidden units jc = tid % 8 + 8 u
  constexpr int TU = CH/8;
  const int pkk = tid>>3, pj = tid&7;
  cd tv[TU];
  cd zv;
  constexpr int sper = (KC*ns)/NQS_DC_THREADS;                   // spins per thread and stage
  int8_t sv[sper];

  auto prefetch = [&](const int s)
  {
    const long long k = k0+(long long)s*KC+pkk;
    const bool ok = (k < k1);
    zv = ok ? (ones ? cmake(1.0, 0.0) : a.zk[k]) : cmake(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < TU; ++u)
    {
      const int jc = pj+8*u;
      tv[u] = (ok && j0+jc < M) ? a.T[k*M+j0+jc] : cmake(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < sper; ++u)
    {
      const int idx = u*NQS_DC_THREADS+tid, kk = idx/ns, i = idx-kk*ns;
      const long long kq = k0+(long long)s*KC+kk;
      sv[u] = (kq < k1 && i < N) ? a.spins[kq*N+i] : (int8_t)0;
    }
  };
  auto commit = [&](const int buf)
  {
    double * cs = Cs+(size_t)buf*KC*PC+pkk*PC;
#pragma unroll
    for (int u = 0; u < TU; ++u)
    {
      const int jc = pj+8*u;
      *reinterpret_cast<double2*>(cs+2*jc) = make_double2(tv[u].x*zv.x+tv[u].y*zv.y, tv[u].x*zv.y-tv[u].y*zv.x);   // conj(T) z
    }
    if (pj == 0) zs[buf*KC+pkk] = zv;
    double * ss = Ss+(size_t)buf*KC*PS;
#pragma unroll
    for (int u = 0; u < sper; ++u)
    {
      const int idx = u*NQS_DC_THREADS+tid, kk = idx/ns, i = idx-kk*ns;
      ss[kk*PS+i] = (double)sv[u];
    }
  };

  if (nstages > 0) { prefetch(0); commit(0); }
  __syncthreads();
  for (int s = 0; s < nstages; ++s)
  {
    const int buf = s&1;
    if (s+1 < nstages) prefetch(s+1);
    const double * ss = Ss+(size_t)buf*KC*PS;
    const double * cs = Cs+(size_t)buf*KC*PC;
    {
      const double * ap = ss+t*PS+wm*MTW*8+g;
      const double * bp = cs+t*PC+wn*NTW*8+g;
#pragma unroll 2
      for (int ks = 0; ks < KC/4; ++ks)
      {
        double af[MTW], bf[NTW];
#pragma unroll
        for (int m = 0; m < MTW; ++m) af[m] = ap[ks*4*PS+m*8];
#pragma unroll
        for (int n = 0; n < NTW; ++n) bf[n] = bp[ks*4*PC+n*8];
#pragma unroll
        for (int m = 0; m < MTW; ++m)
#pragma unroll
          for (int n = 0; n < NTW; ++n) dmma884(acc[m][n][0], acc[m][n][1], af[m], bf[n]);
      }
    }
    // short blocks: column sums of C (b block / FFNN b1 block), S^T z (RBM a block), conj(L) z (FFNN w1o block)
    if (tid < CW)
    {
#pragma unroll 8
      for (int kk = 0; kk < KC; ++kk) { const double v = cs[kk*PC+tid]; bsum += v; b2sum = fma(v, v, b2sum); }
    }
    if (do_a && tid < N)
    {
#pragma unroll 8
      for (int kk = 0; kk < KC; ++kk)
      {
        const double sp = ss[kk*PS+tid];
        const cd z = zs[buf*KC+kk];
        asx = fma(sp, z.x, asx); asy = fma(sp, z.y, asy);
      }
    }
    if (s+1 < nstages) commit(buf^1);
    __syncthreads();
  }

  const long long P = a.P, NM = (long long)N*M;
  double * base = a.part+(size_t)blockIdx.z*(size_t)a.part_stride+(size_t)blockIdx.y*2*(size_t)P;
  if (ones)
  {
    double * q = a.abs2+(size_t)blockIdx.y*3*M;
    if (tid < CW && j0+(tid>>1) < M) q[2*j0+tid] = b2sum;
  }
#pragma unroll
  for (int m = 0; m < MTW; ++m)
  {
    const int i = (wm*MTW+m)*8+g;
    if (i >= N) continue;
#pragma unroll
    for (int n = 0; n < NTW; ++n)
    {
      const int j = j0+(wn*NTW+n)*4+t;
      if (j >= M) continue;
      const long long p = (MODEL == MODEL_RBM) ? (long long)i*M+j : (long long)j*N+i;
      base[p] = acc[m][n][0]; base[P+p] = acc[m][n][1];
    }
  }
  if (tid < CW)
  {
    const int j = j0+(tid>>1);
    if (j < M)
    {
      const long long p = (MODEL == MODEL_RBM) ? NM+N+j : NM+j;
      base[(size_t)(tid&1)*P+p] = bsum;
    }
  }
  if (do_a && tid < N) { base[NM+tid] = asx; base[P+NM+tid] = asy; }
}

// FFNN w1o block of the same product: part[chunk][.][NM+M+j] = sum_k conj(L_kj) z_k (a [K][M] GEMV, too thin for the tensor
// cores), and sum_k |L_kj|^2 for the setup slice.  Same chunks / slices / output planes as spin_cols_dmma_kernel.
#define NQS_LB_THREADS 128
__global__ void __launch_bounds__(NQS_LB_THREADS) ffnn_lblock_kernel(const ColsArgs a)
{
  if (a.done != nullptr && *a.done) return;
  const int M = a.M, j = blockIdx.x*NQS_LB_THREADS+threadIdx.x;
  const bool ones = (a.zmode == 1 && blockIdx.z == 0);
  const long long k0 = (long long)blockIdx.y*a.rows_per_chunk;
  const long long k1 = (k0+a.rows_per_chunk < a.K) ? k0+a.rows_per_chunk : a.K;
  if (j >= M) return;
  double lx = 0.0, ly = 0.0, l2 = 0.0;
  long long k = k0;
  for (; k+8 <= k1; k += 8)
  {
    cd l[8], z[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { l[u] = a.L[(k+u)*M+j]; z[u] = ones ? cmake(1.0, 0.0) : a.zk[k+u]; }
#pragma unroll
    for (int u = 0; u < 8; ++u) { lx += l[u].x*z[u].x+l[u].y*z[u].y; ly += l[u].x*z[u].y-l[u].y*z[u].x; l2 += cnorm(l[u]); }
  }
  for (; k < k1; ++k)
  {
    const cd l = a.L[k*M+j], z = ones ? cmake(1.0, 0.0) : a.zk[k];
    lx += l.x*z.x+l.y*z.y; ly += l.x*z.y-l.y*z.x; l2 += cnorm(l);
  }
  const long long P = a.P, NM = (long long)a.N*M;
  double * base = a.part+(size_t)blockIdx.z*(size_t)a.part_stride+(size_t)blockIdx.y*2*(size_t)P;
  base[NM+M+j] = lx; base[P+NM+M+j] = ly;
  if (ones) a.abs2[(size_t)blockIdx.y*3*M+2*M+j] = l2;
}

// SR setup sums from the two slices of spin_cols_dmma_kernel (zmode = 1), chunks folded in fixed order:
//   sums = [sum O (re P | im P) | sum O conj(h) (re P | im P) | sum |O|^2 (P)]   (layout of setup_finalize_kernel)
// |O_kp|^2 of the W block does not depend on the site (s^2 = 1); the RBM a block is K_loc.
template <int MODEL>
__global__ void setup_fold_struct_kernel(const int N, const int M, const long long P, const long long K, const int nchunks,
  const double * __restrict__ part, const long long part_stride, const double * __restrict__ abs2, double * __restrict__ sums)
{
  const long long NM = (long long)N*M;
  for (long long p = (long long)blockIdx.x*blockDim.x+threadIdx.x; p < P; p += (long long)gridDim.x*blockDim.x)
  {
    double ax = 0.0, ay = 0.0, bx = 0.0, by = 0.0;
    // the loads of 8 chunks are issued before the first add: a serial chain pays one L2 round trip per chunk (28 us for the ~37
    // chunks of a 2048-chain shard, ncu launch list of round 2); the order of the adds stays fixed
    for (int c0 = 0; c0 < nchunks; c0 += 8)
    {
      double v[8][4];
#pragma unroll
      for (int u = 0; u < 8; ++u)
      {
        const bool ok = (c0+u < nchunks);
        const double * b0 = part+(size_t)(c0+u)*2*P+p;
        v[u][0] = ok ? __ldcg(b0) : 0.0; v[u][1] = ok ? __ldcg(b0+P) : 0.0;
        v[u][2] = ok ? __ldcg(b0+part_stride) : 0.0; v[u][3] = ok ? __ldcg(b0+part_stride+P) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) { ax += v[u][0]; ay += v[u][1]; bx += v[u][2]; by += v[u][3]; }
    }
    sums[p] = ax; sums[P+p] = -ay;
    sums[2*P+p] = bx; sums[3*P+p] = -by;
    int col;            // index into abs2's [2M | M] row: pair (2j, 2j+1) of T, or 2M + j of L
    bool pair = true, visible = false;
    if (MODEL == MODEL_RBM)
    {
      if (p < NM) col = (int)(p%M);
      else if (p < NM+N) { col = 0; visible = true; }
      else col = (int)(p-NM-N);
    }
    else
    {
      if (p < NM) col = (int)(p/N);
      else if (p < NM+M) col = (int)(p-NM);
      else { col = (int)(p-NM-M); pair = false; }
    }
    double d = 0.0;
    if (visible) d = (double)K;
    else
      for (int c0 = 0; c0 < nchunks; c0 += 8)
      {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
        {
          const double * q = abs2+(size_t)(c0+u)*3*M;
          v[u] = (c0+u < nchunks) ? (pair ? (__ldcg(q+2*col)+__ldcg(q+2*col+1)) : __ldcg(q+2*M+col)) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) d += v[u];
      }
    sums[4*P+p] = d;
  }
}
} // namespace nqs
