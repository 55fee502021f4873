// Structured S*v (opt-in, NQS_FLAG_STRUCTURED_SV): the same product  traw_p = sum_k conj(O_kp) (O_k . v)  WITHOUT the O matrix.
//
// Every row of O is an outer product plus two short blocks (see setup_structured_kernel in sr_kernels.cuh),
//   RBM   O_k = [ s_ki T_kj (i*M+j) | s_ki | T_kj ],   T = tanh(theta)
//   FFNN  O_k = [ s_ki T_kj (j*N+i) | T_kj | L_kj ],   T = tanh(theta) w1o,  L = logcosh(theta)
// so with v = [V | v1 | v2]
//   z_k = (O v)_k       = sum_j T_kj ( sum_i s_ki V_ij + v2_j ) + s_k . v1              (FFNN: ... + v1_j ..., + sum_j L_kj v2_j)
//   (O^H z)_ij          = sum_k s_ki conj(T_kj) z_k ,  a-block sum_k s_ki z_k ,  b-block sum_k conj(T_kj) z_k
// i.e. two signed accumulations of K*N*M complex terms each (2.1 G fp64 FMAs at N=128, M=256, K=16384) that read the
// [K][N] spins and the [K][M] hidden-unit values (70 MB) instead of streaming the 8.7 GB of O -- the HBM bound of the
// reference's Zgemm + Zgemv pair (gpu/include/functor_for_CG.cuh:110,121) disappears and O need not exist at all (SURVEY 7,
// last bullet).  The north star grades the explicit-O formulation, which stays the default; this mode is reported
// separately by bench.py.
#pragma once
#include "device_math.cuh"
#include "sampler_kernels.cuh"
#include "sr_kernels.cuh"

namespace nqs
{
// T (and, FFNN, L) of the current chain state: one pass over theta
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

// FFNN keeps the W block of v transposed (j*N+i): bring it to the natural i*M+j layout once per product
__global__ void transpose_wblock_kernel(const int N, const int M, const cd * __restrict__ v, cd * __restrict__ out)
{
  const long long total = (long long)N*M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x)
  {
    const int i = (int)(idx/M), j = (int)(idx-(long long)i*M);
    out[idx] = v[(size_t)j*N+i];
  }
}

// z_k = O_k . v.  Same tiling as theta_tiled_kernel: a CTA takes NQS_TH_CH chains, thread t owns hidden unit j = t (+256, ..).
//   Vw: W block of v in i*M+j layout;  vh: the per-hidden-unit block that multiplies T (RBM v2 = b block, FFNN v1 = b1 block);
//   va: RBM a block (multiplies s), nullptr for FFNN;  vl: FFNN w1o block (multiplies L), nullptr for RBM.
template <int MODEL>
__global__ void __launch_bounds__(NQS_TH_THREADS) sv_struct_z_kernel(const int N, const int M, const long long K,
  const int8_t * __restrict__ spins, const cd * __restrict__ T, const cd * __restrict__ L, const cd * __restrict__ Vw,
  const cd * __restrict__ vh, const cd * __restrict__ va, const cd * __restrict__ vl, cd * __restrict__ zk, const int * __restrict__ done)
{
  if (done != nullptr && *done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double * sp = reinterpret_cast<double*>(smem_raw);                    // [N][NQS_TH_CH]
  cd * red = reinterpret_cast<cd*>(sp+(size_t)N*NQS_TH_CH);              // [warps][NQS_TH_CH]
  const int t = threadIdx.x, lane = t&31, w = t>>5;
  const long long kbase = (long long)blockIdx.x*NQS_TH_CH;
  const int nk = (int)((K-kbase < NQS_TH_CH) ? K-kbase : NQS_TH_CH);
  for (int idx = t; idx < N*NQS_TH_CH; idx += NQS_TH_THREADS)
  {
    const int i = idx/NQS_TH_CH, c = idx-i*NQS_TH_CH;
    sp[idx] = (c < nk) ? (double)spins[(kbase+c)*N+i] : 0.0;
  }
  __syncthreads();
  cd zs[NQS_TH_CH];
#pragma unroll
  for (int c = 0; c < NQS_TH_CH; ++c) zs[c] = cmake(0.0, 0.0);
  for (int j = t; j < M; j += NQS_TH_THREADS)
  {
    cd acc[NQS_TH_CH];
    const cd h0 = vh[j];
#pragma unroll
    for (int c = 0; c < NQS_TH_CH; ++c) acc[c] = h0;
    for (int i = 0; i < N; ++i)
    {
      const cd wv = Vw[(size_t)i*M+j];
      const double2 * srow = reinterpret_cast<const double2*>(sp+(size_t)i*NQS_TH_CH);
#pragma unroll
      for (int c2 = 0; c2 < NQS_TH_CH/2; ++c2)
      {
        const double2 s2 = srow[c2];
        acc[2*c2].x = fma(s2.x, wv.x, acc[2*c2].x); acc[2*c2].y = fma(s2.x, wv.y, acc[2*c2].y);
        acc[2*c2+1].x = fma(s2.y, wv.x, acc[2*c2+1].x); acc[2*c2+1].y = fma(s2.y, wv.y, acc[2*c2+1].y);
      }
    }
    const cd lw = (MODEL == MODEL_FFNN) ? vl[j] : cmake(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < NQS_TH_CH; ++c)
    {
      if (c < nk)
      {
        cd term = cmul(T[(kbase+c)*M+j], acc[c]);
        if (MODEL == MODEL_FFNN) term = cadd(term, cmul(L[(kbase+c)*M+j], lw));
        zs[c] = cadd(zs[c], term);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NQS_TH_CH; ++c)
  {
    const cd s = warp_sum(zs[c]);
    if (lane == 0) red[w*NQS_TH_CH+c] = s;
  }
  __syncthreads();
  for (int c = w; c < nk; c += NQS_TH_THREADS/32)
  {
    cd sa = cmake(0.0, 0.0);
    if (MODEL == MODEL_RBM)
    {
      for (int i = lane; i < N; i += 32)
      {
        const double s = sp[(size_t)i*NQS_TH_CH+c];
        const cd ai = va[i];
        sa.x = fma(s, ai.x, sa.x);
        sa.y = fma(s, ai.y, sa.y);
      }
      sa = warp_sum(sa);
    }
    if (lane == 0)
    {
      cd tot = sa;
      for (int ww = 0; ww < NQS_TH_THREADS/32; ++ww) tot = cadd(tot, red[ww*NQS_TH_CH+c]);
      zk[kbase+c] = tot;
    }
  }
}

// part[rb][{re,im}][p] = sum_{k in row block rb} conj(O_kp) z_k from the factors; same tiling and output convention as
// setup_structured_kernel (grid = hidden-unit tiles of 16 x row blocks; a thread owns hidden unit j and IPT consecutive sites)
template <int MODEL, int IPT>
__global__ void __launch_bounds__(NQS_SS_THREADS) sv_struct_cols_kernel(const int N, const int M, const long long K,
  const int8_t * __restrict__ spins, const cd * __restrict__ T, const cd * __restrict__ L, const cd * __restrict__ zk,
  double * __restrict__ part, const long long rows_per_block, const int * __restrict__ done)
{
  if (done != nullptr && *done) return;
  __shared__ cd Cz[NQS_SS_CH][NQS_SS_JT], Lz[NQS_SS_CH][NQS_SS_JT];     // conj(T_kj) z_k, conj(L_kj) z_k
  __shared__ cd zsh[NQS_SS_CH];
  extern __shared__ __align__(16) unsigned char smem_raw[];               // spins chunk [NQS_SS_CH][16*IPT] as doubles
  constexpr int npad = 16*IPT;
  double * sp = reinterpret_cast<double*>(smem_raw);
  const int t = threadIdx.x, jl = t%NQS_SS_JT, ig = t/NQS_SS_JT;
  const int j = blockIdx.x*NQS_SS_JT+jl;
  const bool jok = (j < M);
  const long long k0 = (long long)blockIdx.y*rows_per_block;
  const long long k1 = (k0+rows_per_block < K) ? k0+rows_per_block : K;
  double ax[IPT], ay[IPT];
#pragma unroll
  for (int m = 0; m < IPT; ++m) { ax[m] = 0; ay[m] = 0; }
  double bx = 0, by = 0, lx = 0, ly = 0, sx = 0, sy = 0;
  const bool do_a = (MODEL == MODEL_RBM && blockIdx.x == 0);
  for (long long kc = k0; kc < k1; kc += NQS_SS_CH)
  {
    const int nk = (int)((k1-kc < NQS_SS_CH) ? k1-kc : NQS_SS_CH);
    __syncthreads();
    {
      const int kk = ig;
      cd c = cmake(0.0, 0.0), l = c;
      if (kk < nk && jok)
      {
        const cd z = zk[kc+kk], tv = T[(kc+kk)*M+j];
        c = cmake(tv.x*z.x+tv.y*z.y, tv.x*z.y-tv.y*z.x);             // conj(T) z
        if (MODEL == MODEL_FFNN)
        {
          const cd lv = L[(kc+kk)*M+j];
          l = cmake(lv.x*z.x+lv.y*z.y, lv.x*z.y-lv.y*z.x);
        }
      }
      Cz[kk][jl] = c;
      if (MODEL == MODEL_FFNN) Lz[kk][jl] = l;
      if (t < NQS_SS_CH) zsh[t] = (t < nk) ? zk[kc+t] : cmake(0.0, 0.0);
    }
    for (int idx = t; idx < NQS_SS_CH*npad; idx += NQS_SS_THREADS)
    {
      const int kk = idx/npad, i = idx-kk*npad;
      sp[idx] = (kk < nk && i < N) ? (double)spins[(kc+kk)*N+i] : 0.0;
    }
    __syncthreads();
    for (int kk = 0; kk < nk; ++kk)
    {
      const cd c = Cz[kk][jl];
      const double * srow = sp+kk*npad+ig*IPT;
#pragma unroll
      for (int m = 0; m < IPT; ++m)
      {
        const double s = srow[m];
        ax[m] = fma(s, c.x, ax[m]); ay[m] = fma(s, c.y, ay[m]);
      }
      if (ig == 0) { bx += c.x; by += c.y; }
      if (MODEL == MODEL_FFNN && ig == 1) { const cd l = Lz[kk][jl]; lx += l.x; ly += l.y; }
    }
    if (do_a)
      for (int i = t; i < N; i += NQS_SS_THREADS)
        for (int kk = 0; kk < nk; ++kk)
        {
          const double s = sp[kk*npad+i];
          const cd z = zsh[kk];
          sx = fma(s, z.x, sx); sy = fma(s, z.y, sy);
        }
  }
  const long long P = (MODEL == MODEL_RBM) ? (long long)N*M+N+M : (long long)N*M+2*M;
  const long long NM = (long long)N*M;
  double * base = part+(size_t)blockIdx.y*2*P;
  if (jok)
  {
#pragma unroll
    for (int m = 0; m < IPT; ++m)
    {
      const int i = ig*IPT+m;
      if (i < N)
      {
        const long long p = (MODEL == MODEL_RBM) ? (long long)i*M+j : (long long)j*N+i;
        base[p] = ax[m]; base[P+p] = ay[m];
      }
    }
    if (ig == 0)
    {
      const long long p = (MODEL == MODEL_RBM) ? NM+N+j : NM+j;
      base[p] = bx; base[P+p] = by;
    }
    if (MODEL == MODEL_FFNN && ig == 1)
    {
      const long long p = NM+M+j;
      base[p] = lx; base[P+p] = ly;
    }
  }
  if (do_a)
    for (int i = t; i < N; i += NQS_SS_THREADS)
    {
      const long long p = NM+i;
      base[p] = sx; base[P+p] = sy;
    }
}
} // namespace nqs
