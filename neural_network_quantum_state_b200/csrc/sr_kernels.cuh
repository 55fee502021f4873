// Stochastic-reconfiguration side of the path: O writer, single-pass SR setup sums, the two-pass fallback of the S*v product
// (the one-pass cluster kernel is in sv_fused.cuh, the CG iteration body in cg_fused.cuh) and the parameter update.
// O is [K_loc][P] row-major complex fp64 exactly like the reference's lnpsiGradients (k*P + p); all indices are 64-bit
// (ref int32 overflow at cfg5).  Every reduction is two-stage with a fixed order -> run-to-run deterministic.
#pragma once
#include "device_math.cuh"
#include "sampler_kernels.cuh"

namespace nqs
{
// streaming (evict-first) 16-byte accesses for the O matrix: it is read once per pass and never fits L2
__device__ __forceinline__ cd ld_stream(const cd * p)
{
  cd v;
  asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(cd * p, const cd v)
{
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// O writer.  ref: RBM__GetGradientsOfParameters__ (impl_neural_quantum_state.cuh:1426-1449) + the K*P D2D copy (:152);
// FFNN__GetGradientsOfParameters__ + FFNN__GetlnpsiGradients__ (:1622-1663, W block transposed to j*N+i).
// One CTA per chain: tanh(theta_kj) is evaluated ONCE per (k,j) into shared memory (the reference recomputes it N times),
// then the row O_k is streamed out with coalesced 16-byte stores.
// ---------------------------------------------------------------------------------------------------------------------
template <int MODEL>
__global__ void __launch_bounds__(256) oderiv_kernel(const int N, const int M, const long long K, const cd * params,
  const int8_t * __restrict__ spins, const cd * __restrict__ theta, cd * __restrict__ O)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cd * T = reinterpret_cast<cd*>(smem_raw);            // [M]  RBM: tanh(theta_j); FFNN: tanh(theta_j)*w1o_j
  cd * L = T+M;                                        // [M]  FFNN only: logcosh(theta_j)
  double * s = reinterpret_cast<double*>(T+(MODEL == MODEL_FFNN ? 2*M : M)); // [N]
  const long long k = blockIdx.x;
  const ModelPtrs mp = model_ptrs(MODEL, params, N, M);
  for (int j = threadIdx.x; j < M; j += blockDim.x)
  {
    const cd th = theta[k*M+j];
    const cd t = c_tanh(th);
    if (MODEL == MODEL_RBM) T[j] = t;
    else { T[j] = cmul(t, mp.w1o[j]); L[j] = c_logcosh(th); }
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    s[i] = (double)spins[k*N+i];
  __syncthreads();
  const long long P = (MODEL == MODEL_RBM) ? (long long)N*M+N+M : (long long)N*M+2*M;
  cd * row = O+k*P;
  const int NM = N*M;
  if (MODEL == MODEL_RBM)
  {
    for (int p = threadIdx.x; p < NM; p += blockDim.x)
    {
      const int i = p/M, j = p-i*M;
      st_stream(row+p, cscale(T[j], s[i]));
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) st_stream(row+NM+i, cmake(s[i], 0.0));
    for (int j = threadIdx.x; j < M; j += blockDim.x) st_stream(row+NM+N+j, T[j]);
  }
  else
  {
    for (int p = threadIdx.x; p < NM; p += blockDim.x)
    {
      const int j = p/N, i = p-j*N;
      st_stream(row+p, cscale(T[j], s[i]));
    }
    for (int j = threadIdx.x; j < M; j += blockDim.x) st_stream(row+NM+j, T[j]);
    for (int j = threadIdx.x; j < M; j += blockDim.x) st_stream(row+NM+M+j, L[j]);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Translation-symmetric RBM (ref RBMTrSymm, impl_neural_quantum_state.cuh:301-538).
// expand: wf[i][f*N+j] = w[f][(i+j)%N], af[i] = a[0], bf[f*N+j] = b[f]  (ref RBMTrSymm__ConstructWeightAndBias__, :1523-1553) into
// the plain-RBM parameter layout [wf (i*M+j') | af | bf], M = alpha*N, which every sampler kernel of this library reads.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void trsymm_expand_kernel(const int N, const int alpha, const cd * __restrict__ vars, cd * __restrict__ full)
{
  const int M = alpha*N;
  const cd * w = vars;
  const cd a0 = vars[(size_t)N*alpha];
  const cd * b = vars+(size_t)N*alpha+1;
  const long long NM = (long long)N*M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < NM+N+M; idx += (long long)gridDim.x*blockDim.x)
  {
    if (idx < NM)
    {
      const int i = (int)(idx/M), c = (int)(idx-(long long)i*M), f = c/N, j = c-f*N;
      full[idx] = w[(size_t)f*N+(i+j)%N];
    }
    else if (idx < NM+N) full[idx] = a0;
    else full[idx] = b[(idx-NM-N)/N];
  }
}

// O writer (ref RBMTrSymm__GetGradientsOfParameters__, :1487-1521): per chain, with T = tanh(theta) of the expanded hidden layer,
//   d_w[f*N+i] = sum_j T[f*N+j] s[(N+i-j)%N] ,  d_a = sum_i s_i ,  d_b[f] = sum_j T[f*N+j]      (row = [d_w | d_a | d_b]).
// One CTA per chain; tanh once per hidden unit into shared memory (the reference evaluates it N times per unit).
__global__ void __launch_bounds__(256) oderiv_trsymm_kernel(const int N, const int alpha, const long long K,
  const int8_t * __restrict__ spins, const cd * __restrict__ theta, cd * __restrict__ O)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = alpha*N;
  cd * T = reinterpret_cast<cd*>(smem_raw);             // [M]
  double * s = reinterpret_cast<double*>(T+M);         // [N]
  const long long k = blockIdx.x;
  for (int j = threadIdx.x; j < M; j += blockDim.x) T[j] = c_tanh(theta[k*M+j]);
  for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = (double)spins[k*N+i];
  __syncthreads();
  const long long P = (long long)N*alpha+1+alpha;
  cd * row = O+k*P;
  for (int q = threadIdx.x; q < M; q += blockDim.x)
  {
    const int f = q/N, i = q-f*N;
    cd acc = cmake(0.0, 0.0);
    for (int j = 0; j < N; ++j)
    {
      const cd t = T[f*N+j];
      const double sv = s[(N+i-j)%N];
      acc.x += t.x*sv; acc.y += t.y*sv;
    }
    st_stream(row+q, acc);
  }
  if (threadIdx.x == 0)
  {
    double sa = 0.0;
    for (int i = 0; i < N; ++i) sa += s[i];
    st_stream(row+M, cmake(sa, 0.0));
  }
  for (int f = threadIdx.x; f < alpha; f += blockDim.x)
  {
    cd acc = cmake(0.0, 0.0);
    for (int j = 0; j < N; ++j) acc = cadd(acc, T[f*N+j]);
    st_stream(row+M+1+f, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Z2- and parity-symmetric RBM (ref RBMZ2PrSymm, impl_neural_quantum_state.cuh:540-745; driver gpu/src/LICH-train_rbmz2prsymm.cu).
// Variables [w (i*alpha+f) | b (alpha)], P = N*alpha + alpha; four hidden units per filter.
// expand (ref RBMZ2PrSymm__ConstructWeightAndBias__, :1588-1618) into the plain-RBM layout [wf (i*M+c) | af = 0 | bf], M = 4 alpha:
//   wf[i][4f+0] = w[i][f], wf[i][4f+1] = -w[i][f], wf[i][4f+2] = w[N-1-i][f], wf[i][4f+3] = -w[N-1-i][f], bf[4f+j] = b[f]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void z2pr_expand_kernel(const int N, const int alpha, const cd * __restrict__ vars, cd * __restrict__ full)
{
  const int M = 4*alpha;
  const cd * w = vars;
  const cd * b = vars+(size_t)N*alpha;
  const long long NM = (long long)N*M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < NM+N+M; idx += (long long)gridDim.x*blockDim.x)
  {
    if (idx < NM)
    {
      const int i = (int)(idx/M), c = (int)(idx-(long long)i*M), f = c>>2, j = c&3;
      const cd v = w[(size_t)((j&2) ? N-1-i : i)*alpha+f];
      full[idx] = (j&1) ? cmake(-v.x, -v.y) : v;
    }
    else if (idx < NM+N) full[idx] = cmake(0.0, 0.0);
    else full[idx] = b[(idx-NM-N)>>2];
  }
}

// O writer (ref RBMZ2PrSymm__GetGradientsOfParameters__, :1556-1585): with T = tanh(theta) of the four units of filter f,
//   d_w[i*alpha+f] = (T[4f] - T[4f+1]) s_i + (T[4f+2] - T[4f+3]) s_{N-1-i} ,  d_b[f] = T[4f] + T[4f+1] + T[4f+2] + T[4f+3]
__global__ void __launch_bounds__(256) oderiv_z2pr_kernel(const int N, const int alpha, const long long K,
  const int8_t * __restrict__ spins, const cd * __restrict__ theta, cd * __restrict__ O)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = 4*alpha;
  cd * T = reinterpret_cast<cd*>(smem_raw);             // [M]
  double * s = reinterpret_cast<double*>(T+M);         // [N]
  const long long k = blockIdx.x;
  for (int j = threadIdx.x; j < M; j += blockDim.x) T[j] = c_tanh(theta[k*M+j]);
  for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = (double)spins[k*N+i];
  __syncthreads();
  const long long P = (long long)N*alpha+alpha;
  cd * row = O+k*P;
  const int NA = N*alpha;
  for (int q = threadIdx.x; q < NA; q += blockDim.x)
  {
    const int i = q/alpha, f = q-i*alpha;
    const cd d0 = csub(T[4*f], T[4*f+1]), d1 = csub(T[4*f+2], T[4*f+3]);
    st_stream(row+q, cadd(cscale(d0, s[i]), cscale(d1, s[N-1-i])));
  }
  for (int f = threadIdx.x; f < alpha; f += blockDim.x)
    st_stream(row+NA+f, cadd(cadd(T[4*f], T[4*f+1]), cadd(T[4*f+2], T[4*f+3])));
}

// ---------------------------------------------------------------------------------------------------------------------
// Translation-symmetric FNN (ref FFNNTrSymm, impl_neural_quantum_state.cuh:1019-1223; driver gpu/src/LICH-train_ffnntrsymm.cu).
// Variables [wi1 (f*N+i) | b1 (alpha) | w1o (alpha)], P = N*alpha + 2 alpha.
// expand (ref FFNNTrSymm__ConstructWeightAndBias__, :1693-1717) into the plain-FFNN layout [W1 (i*M+c) | b1f | w1of], M = alpha*N:
//   W1[i][f*N+j] = wi1[f][(i+j)%N], b1f[f*N+j] = b1[f], w1of[f*N+j] = w1o[f]
// ---------------------------------------------------------------------------------------------------------------------
__global__ void ffnntr_expand_kernel(const int N, const int alpha, const cd * __restrict__ vars, cd * __restrict__ full)
{
  const int M = alpha*N;
  const cd * w = vars;
  const cd * b1 = vars+(size_t)N*alpha;
  const cd * w1o = b1+alpha;
  const long long NM = (long long)N*M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < NM+2*M; idx += (long long)gridDim.x*blockDim.x)
  {
    if (idx < NM)
    {
      const int i = (int)(idx/M), c = (int)(idx-(long long)i*M), f = c/N, j = c-f*N;
      full[idx] = w[(size_t)f*N+(i+j)%N];
    }
    else if (idx < NM+M) full[idx] = b1[(idx-NM)/N];
    else full[idx] = w1o[(idx-NM-M)/N];
  }
}

// O writer (ref FFNNTrSymm__GetGradientsOfParameters__, :1720-1750): with T' = tanh(theta) w1of and L = log cosh(theta),
//   d_wi1[f*N+i] = sum_j T'[f*N+j] s[(N+i-j)%N] ,  d_b1[f] = sum_j T'[f*N+j] ,  d_w1o[f] = sum_j L[f*N+j]
__global__ void __launch_bounds__(256) oderiv_ffnntr_kernel(const int N, const int alpha, const long long K, const cd * __restrict__ w1of,
  const int8_t * __restrict__ spins, const cd * __restrict__ theta, cd * __restrict__ O)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int M = alpha*N;
  cd * T = reinterpret_cast<cd*>(smem_raw);             // [M] tanh(theta_c) w1of_c
  cd * L = T+M;                                         // [M] log cosh(theta_c)
  double * s = reinterpret_cast<double*>(L+M);         // [N]
  const long long k = blockIdx.x;
  for (int j = threadIdx.x; j < M; j += blockDim.x)
  {
    const cd th = theta[k*M+j];
    T[j] = cmul(c_tanh(th), w1of[j]);
    L[j] = c_logcosh(th);
  }
  for (int i = threadIdx.x; i < N; i += blockDim.x) s[i] = (double)spins[k*N+i];
  __syncthreads();
  const long long P = (long long)N*alpha+2*alpha;
  cd * row = O+k*P;
  for (int q = threadIdx.x; q < M; q += blockDim.x)
  {
    const int f = q/N, i = q-f*N;
    cd acc = cmake(0.0, 0.0);
    for (int j = 0; j < N; ++j)
    {
      const cd t = T[f*N+j];
      const double sv = s[(N+i-j)%N];
      acc.x += t.x*sv; acc.y += t.y*sv;
    }
    st_stream(row+q, acc);
  }
  for (int f = threadIdx.x; f < 2*alpha; f += blockDim.x)
  {
    const cd * src = (f < alpha) ? T+(size_t)f*N : L+(size_t)(f-alpha)*N;
    cd acc = cmake(0.0, 0.0);
    for (int j = 0; j < N; ++j) acc = cadd(acc, src[j]);
    st_stream(row+M+f, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Column-direction passes over O.  Grid = (column tiles, row blocks); one thread owns ONE column p of its tile and walks the
// rows of its row block, so a warp reads 32 consecutive complex numbers (512 B) per row: fully coalesced.  Row-block
// partials go to part[rb][...][P] and are summed in fixed order by colsum_reduce_kernel.
// ---------------------------------------------------------------------------------------------------------------------
#define NQS_COL_THREADS 128
#define NQS_COL_UNROLL 8

// SR setup sums in ONE pass (the reference takes four: c6 x2, c7, k19; optimizer.cuh:140-143, functor_for_CG.cuh:99-102):
//   part[rb][0..1][p] = sum_k O_kp ; part[rb][2..3][p] = sum_k O_kp conj(h_k) ; part[rb][4][p] = sum_k |O_kp|^2
__global__ void __launch_bounds__(NQS_COL_THREADS) setup_partial_kernel(const long long K, const long long P,
  const cd * __restrict__ O, const cd * __restrict__ htilda, double * __restrict__ part, const long long rows_per_block)
{
  const long long p = (long long)blockIdx.x*NQS_COL_THREADS+threadIdx.x;
  const long long k0 = (long long)blockIdx.y*rows_per_block;
  const long long k1 = (k0+rows_per_block < K) ? k0+rows_per_block : K;
  if (p >= P) return;
  double so_x = 0, so_y = 0, sh_x = 0, sh_y = 0, s2 = 0;
  long long k = k0;
  for (; k+NQS_COL_UNROLL <= k1; k += NQS_COL_UNROLL)
  {
    cd o[NQS_COL_UNROLL];
#pragma unroll
    for (int u = 0; u < NQS_COL_UNROLL; ++u) o[u] = ld_stream(O+(k+u)*P+p);
#pragma unroll
    for (int u = 0; u < NQS_COL_UNROLL; ++u)
    {
      const cd h = htilda[k+u];
      so_x += o[u].x; so_y += o[u].y;
      sh_x += o[u].x*h.x+o[u].y*h.y;     // O * conj(h)
      sh_y += o[u].y*h.x-o[u].x*h.y;
      s2 += o[u].x*o[u].x+o[u].y*o[u].y;
    }
  }
  for (; k < k1; ++k)
  {
    const cd o = ld_stream(O+k*P+p), h = htilda[k];
    so_x += o.x; so_y += o.y;
    sh_x += o.x*h.x+o.y*h.y;
    sh_y += o.y*h.x-o.x*h.y;
    s2 += o.x*o.x+o.y*o.y;
  }
  double * base = part+(size_t)blockIdx.y*5*P;
  base[p] = so_x; base[P+p] = so_y; base[2*P+p] = sh_x; base[3*P+p] = sh_y; base[4*P+p] = s2;
}

// second pass of S*v:  part[rb][0..1][p] = sum_{k in rb} conj(O_kp) z_k      (ref c9 Zgemv + conj tricks, functor_for_CG.cuh:113-124)
__global__ void __launch_bounds__(NQS_COL_THREADS) matvec_cols_partial_kernel(const long long K, const long long P,
  const cd * __restrict__ O, const cd * __restrict__ zk, double * __restrict__ part, const long long rows_per_block,
  const int * __restrict__ done)
{
  if (done != nullptr && *done) return;
  const long long p = (long long)blockIdx.x*NQS_COL_THREADS+threadIdx.x;
  const long long k0 = (long long)blockIdx.y*rows_per_block;
  const long long k1 = (k0+rows_per_block < K) ? k0+rows_per_block : K;
  if (p >= P) return;
  double ax = 0, ay = 0;
  long long k = k0;
  for (; k+NQS_COL_UNROLL <= k1; k += NQS_COL_UNROLL)
  {
    cd o[NQS_COL_UNROLL];
#pragma unroll
    for (int u = 0; u < NQS_COL_UNROLL; ++u) o[u] = ld_stream(O+(k+u)*P+p);
#pragma unroll
    for (int u = 0; u < NQS_COL_UNROLL; ++u)
    {
      const cd z = zk[k+u];
      ax += o[u].x*z.x+o[u].y*z.y;   // conj(O) * z
      ay += o[u].x*z.y-o[u].y*z.x;
    }
  }
  for (; k < k1; ++k)
  {
    const cd o = ld_stream(O+k*P+p), z = zk[k];
    ax += o.x*z.x+o.y*z.y;
    ay += o.x*z.y-o.y*z.x;
  }
  double * base = part+(size_t)blockIdx.y*2*P;
  base[p] = ax; base[P+p] = ay;
}

// ---------------------------------------------------------------------------------------------------------------------
// SR setup sums WITHOUT reading O back.  Every row of O is an outer product plus two short blocks,
//   RBM  O_k = [ s_ki T_kj (i*M+j) | s_ki | T_kj ],  T = tanh(theta)         (ref k13, impl_neural_quantum_state.cuh:1426-1449)
//   FFNN O_k = [ s_ki T'_kj (j*N+i) | T'_kj | L_kj ], T' = tanh(theta) w1o, L = logcosh(theta)        (ref k15, :1622-1663)
// so  sum_k O_kp,  sum_k O_kp conj(h_k)  and  sum_k |O_kp|^2  follow from the [K][N] spins and the [K][M] hidden-unit values:
// the W block is S^T T and S^T (T conj(h)) -- signed accumulation, K*N*M*4 fp64 FMAs -- and |O|^2 of the W block does not
// depend on i at all.  This replaces the 8.7 GB pass of setup_partial_kernel (1.49 ms at N=128, M=256, K=16384) by
// ~70 MB of reads; the reference makes FOUR passes over O for the same numbers (optimizer.cuh:140-143, functor_for_CG.cuh:99-102).
// Grid = (hidden-unit tiles of 16, row blocks); a thread owns hidden unit j = t%16 of the tile and IPT consecutive sites.
// Output layout = setup_partial_kernel's: part[rb][0..1] = sum O, [2..3] = sum O conj(h), [4] = sum |O|^2.
// ---------------------------------------------------------------------------------------------------------------------
#define NQS_SS_THREADS 256
#define NQS_SS_JT 16
#define NQS_SS_CH 16

template <int MODEL, int IPT>
__global__ void __launch_bounds__(NQS_SS_THREADS) setup_structured_kernel(const int N, const int M, const long long K,
  const cd * params, const int8_t * __restrict__ spins, const cd * __restrict__ theta, const cd * __restrict__ htilda,
  double * __restrict__ part, const long long rows_per_block)
{
  __shared__ cd Tsh[NQS_SS_CH][NQS_SS_JT], Thsh[NQS_SS_CH][NQS_SS_JT], Lsh[NQS_SS_CH][NQS_SS_JT];
  __shared__ cd hsh[NQS_SS_CH];
  extern __shared__ __align__(16) unsigned char smem_raw[];     // spins chunk [NQS_SS_CH][16*IPT] as doubles, zero beyond N
  constexpr int npad = 16*IPT;
  double * sp = reinterpret_cast<double*>(smem_raw);
  const int t = threadIdx.x, jl = t%NQS_SS_JT, ig = t/NQS_SS_JT;   // ig = 0..15: the thread owns sites ig*IPT .. ig*IPT+IPT-1
  const int j = blockIdx.x*NQS_SS_JT+jl;
  const bool jok = (j < M);
  const ModelPtrs mp = model_ptrs(MODEL, params, N, M);
  const cd w1o = (MODEL == MODEL_FFNN && jok) ? mp.w1o[j] : cmake(1.0, 0.0);
  const long long k0 = (long long)blockIdx.y*rows_per_block;
  const long long k1 = (k0+rows_per_block < K) ? k0+rows_per_block : K;
  double a1x[IPT], a1y[IPT], a2x[IPT], a2y[IPT];
#pragma unroll
  for (int m = 0; m < IPT; ++m) { a1x[m] = 0; a1y[m] = 0; a2x[m] = 0; a2y[m] = 0; }
  double bT[2] = {0, 0}, bTh[2] = {0, 0}, bT2 = 0;      // hidden-bias block (thread group ig == 0)
  double bL[2] = {0, 0}, bLh[2] = {0, 0}, bL2 = 0;      // FFNN: logcosh block (ig == 1)
  const bool do_a = (MODEL == MODEL_RBM && blockIdx.x == 0);    // visible-bias block: hidden-unit tile 0, threads t < N
  double as = 0, ahx = 0, ahy = 0;

  for (long long kc = k0; kc < k1; kc += NQS_SS_CH)
  {
    const int nk = (int)((k1-kc < NQS_SS_CH) ? k1-kc : NQS_SS_CH);
    __syncthreads();
    { // one hidden-unit value per thread: chain kc + ig, hidden unit j
      const int kk = ig;
      cd T = cmake(0.0, 0.0), Th = T, L = T;
      if (kk < nk && jok)
      {
        const cd th = theta[(kc+kk)*M+j], h = htilda[kc+kk];
        T = c_tanh(th);
        if (MODEL == MODEL_FFNN) { T = cmul(T, w1o); L = c_logcosh(th); }
        Th = cmake(T.x*h.x+T.y*h.y, T.y*h.x-T.x*h.y);      // T conj(h)
      }
      Tsh[kk][jl] = T; Thsh[kk][jl] = Th;
      if (MODEL == MODEL_FFNN) Lsh[kk][jl] = L;
      if (t < NQS_SS_CH) hsh[t] = (t < nk) ? htilda[kc+t] : cmake(0.0, 0.0);
    }
    for (int idx = t; idx < NQS_SS_CH*npad; idx += NQS_SS_THREADS)
    {
      const int kk = idx/npad, i = idx-kk*npad;
      sp[idx] = (kk < nk && i < N) ? (double)spins[(kc+kk)*N+i] : 0.0;
    }
    __syncthreads();
    for (int kk = 0; kk < nk; ++kk)
    {
      const cd T = Tsh[kk][jl], Th = Thsh[kk][jl];
      const double * srow = sp+kk*npad+ig*IPT;
#pragma unroll
      for (int m = 0; m < IPT; ++m)
      {
        const double s = srow[m];                          // 0 beyond N
        a1x[m] = fma(s, T.x, a1x[m]); a1y[m] = fma(s, T.y, a1y[m]);
        a2x[m] = fma(s, Th.x, a2x[m]); a2y[m] = fma(s, Th.y, a2y[m]);
      }
      if (ig == 0)
      {
        bT[0] += T.x; bT[1] += T.y; bTh[0] += Th.x; bTh[1] += Th.y; bT2 += T.x*T.x+T.y*T.y;
      }
      if (MODEL == MODEL_FFNN && ig == 1)
      {
        const cd L = Lsh[kk][jl], h = hsh[kk];
        bL[0] += L.x; bL[1] += L.y; bLh[0] += L.x*h.x+L.y*h.y; bLh[1] += L.y*h.x-L.x*h.y; bL2 += L.x*L.x+L.y*L.y;
      }
    }
    if (do_a)
      for (int i = t; i < N; i += NQS_SS_THREADS)         // at most one site per thread for N <= 256
        for (int kk = 0; kk < nk; ++kk)
        {
          const double s = sp[kk*npad+i];
          const cd h = hsh[kk];
          as += s; ahx = fma(s, h.x, ahx); ahy = fma(-s, h.y, ahy);   // s conj(h)
        }
  }
  // sum_k |T_kj|^2 is the |O|^2 sum of every W-block column with this hidden unit: share it through shared memory
  __syncthreads();
  double * t2sh = reinterpret_cast<double*>(&Tsh[0][0]);
  if (ig == 0) t2sh[jl] = bT2;
  __syncthreads();
  const double t2 = t2sh[jl];
  const long long P = (MODEL == MODEL_RBM) ? (long long)N*M+N+M : (long long)N*M+2*M;
  const long long NM = (long long)N*M;
  double * base = part+(size_t)blockIdx.y*5*P;
  if (jok)
  {
#pragma unroll
    for (int m = 0; m < IPT; ++m)
    {
      const int i = ig*IPT+m;
      if (i < N)
      {
        const long long p = (MODEL == MODEL_RBM) ? (long long)i*M+j : (long long)j*N+i;
        base[p] = a1x[m]; base[P+p] = a1y[m]; base[2*P+p] = a2x[m]; base[3*P+p] = a2y[m]; base[4*P+p] = t2;
      }
    }
    if (ig == 0)
    {
      const long long p = (MODEL == MODEL_RBM) ? NM+N+j : NM+j;
      base[p] = bT[0]; base[P+p] = bT[1]; base[2*P+p] = bTh[0]; base[3*P+p] = bTh[1]; base[4*P+p] = bT2;
    }
    if (MODEL == MODEL_FFNN && ig == 1)
    {
      const long long p = NM+M+j;
      base[p] = bL[0]; base[P+p] = bL[1]; base[2*P+p] = bLh[0]; base[3*P+p] = bLh[1]; base[4*P+p] = bL2;
    }
  }
  if (do_a)
  {
    const double nrows = (double)(k1 > k0 ? k1-k0 : 0);
    for (int i = t; i < N; i += NQS_SS_THREADS)
    {
      const long long p = NM+i;
      base[p] = as; base[P+p] = 0.0; base[2*P+p] = ahx; base[3*P+p] = ahy; base[4*P+p] = nrows;   // s^2 = 1
    }
  }
}

// out[c][p] = sum_rb part[rb][c][p], fixed order.  ncomp = 5 (setup) or 2 (matvec)
__global__ void colsum_reduce_kernel(const long long P, const int ncomp, const int nrb, const double * __restrict__ part,
  double * __restrict__ out, const int * __restrict__ done)
{
  if (done != nullptr && *done) return;
  const long long total = (long long)ncomp*P;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x)
  {
    double s = 0;
    for (int rb = 0; rb < nrb; ++rb)
      s += part[(size_t)rb*total+idx];
    out[idx] = s;
  }
}

// sum_k h_k and sum_k |h_k|^2 -> hs[0..2]   (ref t1 thrust::reduce, optimizer.cuh:133; l2_norm :156); single CTA, deterministic
__global__ void __launch_bounds__(1024) htilda_sums_kernel(const long long K, const cd * __restrict__ htilda, double * __restrict__ hs)
{
  __shared__ double sh[3][32];
  double a = 0, b = 0, c = 0;
  for (long long k = threadIdx.x; k < K; k += blockDim.x)
  {
    const cd h = htilda[k];
    a += h.x; b += h.y; c += h.x*h.x+h.y*h.y;
  }
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int w = threadIdx.x>>5, lane = threadIdx.x&31;
  if (lane == 0) { sh[0][w] = a; sh[1][w] = b; sh[2][w] = c; }
  __syncthreads();
  if (w == 0)
  {
    const int nw = blockDim.x>>5;
    a = (lane < nw) ? sh[0][lane] : 0; b = (lane < nw) ? sh[1][lane] : 0; c = (lane < nw) ? sh[2][lane] : 0;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    if (lane == 0) { hs[0] = a; hs[1] = b; hs[2] = c; }
  }
}

// sums = [sum O (re P | im P) | sum O conj(h) (re P | im P) | sum |O|^2 (P) | sum h (2) | sum |h|^2 (1)]  (all-reduced over ranks before)
//   aO = sumO/K ; F = conj( sumOh/K - conj(<h>) aO )  (ref SR__FStep2__, impl_optimizer.cuh:82-96) ; diag = sumO2/K - |aO|^2 (ref k19)
__global__ void setup_finalize_kernel(const long long P, const double inv_ktot, const double * __restrict__ sums,
  cd * __restrict__ aO, cd * __restrict__ F, double * __restrict__ diag, double * __restrict__ hsall)
{
  const double * hs = sums+5*P;
  const cd conj_havg = cmake(hs[0]*inv_ktot, -hs[1]*inv_ktot);
  if (blockIdx.x == 0 && threadIdx.x < 3) hsall[threadIdx.x] = hs[threadIdx.x];   // where the CG kernel and the host read <h>, <|h|^2>
  for (long long p = (long long)blockIdx.x*blockDim.x+threadIdx.x; p < P; p += (long long)gridDim.x*blockDim.x)
  {
    const cd ao = cmake(sums[p]*inv_ktot, sums[P+p]*inv_ktot);
    if (aO) aO[p] = ao;
    if (F)
    {
      const cd fr = cmake(sums[2*P+p]*inv_ktot, sums[3*P+p]*inv_ktot);
      F[p] = cconj(csub(fr, cmul(conj_havg, ao)));
    }
    if (diag) diag[p] = sums[4*P+p]*inv_ktot-cnorm(ao);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// first pass of S*v:  z_k = sum_p O_kp v_p   (ref c8 Zgemm 1xKxP, functor_for_CG.cuh:110).  One CTA per group of
// NQS_ROWS_PER_CTA rows so each v_p fetched (L2-resident, P*16 B) is reused for several rows.
// ---------------------------------------------------------------------------------------------------------------------
#define NQS_ROW_THREADS 256
#define NQS_ROWS_PER_CTA 4
#define NQS_ROW_UNROLL 2

__global__ void __launch_bounds__(NQS_ROW_THREADS) matvec_rows_kernel(const long long K, const long long P,
  const cd * __restrict__ O, const cd * __restrict__ v, cd * __restrict__ zk, const int * __restrict__ done)
{
  if (done != nullptr && *done) return;
  __shared__ double sh[NQS_ROWS_PER_CTA][2][NQS_ROW_THREADS/32];
  const long long k0 = (long long)blockIdx.x*NQS_ROWS_PER_CTA;
  double ax[NQS_ROWS_PER_CTA], ay[NQS_ROWS_PER_CTA];
#pragma unroll
  for (int r = 0; r < NQS_ROWS_PER_CTA; ++r) { ax[r] = 0; ay[r] = 0; }
  const int nrows = (int)((K-k0 < NQS_ROWS_PER_CTA) ? K-k0 : NQS_ROWS_PER_CTA);
  if (nrows == NQS_ROWS_PER_CTA)
  {
    long long p = threadIdx.x;
    for (; p+(NQS_ROW_UNROLL-1)*NQS_ROW_THREADS < P; p += NQS_ROW_UNROLL*NQS_ROW_THREADS)
    {
      cd o[NQS_ROW_UNROLL][NQS_ROWS_PER_CTA], vv[NQS_ROW_UNROLL];
#pragma unroll
      for (int u = 0; u < NQS_ROW_UNROLL; ++u)
      {
        vv[u] = v[p+u*NQS_ROW_THREADS];
#pragma unroll
        for (int r = 0; r < NQS_ROWS_PER_CTA; ++r) o[u][r] = ld_stream(O+(k0+r)*P+p+u*NQS_ROW_THREADS);
      }
#pragma unroll
      for (int u = 0; u < NQS_ROW_UNROLL; ++u)
#pragma unroll
        for (int r = 0; r < NQS_ROWS_PER_CTA; ++r)
        {
          ax[r] += o[u][r].x*vv[u].x-o[u][r].y*vv[u].y;
          ay[r] += o[u][r].x*vv[u].y+o[u][r].y*vv[u].x;
        }
    }
    for (; p < P; p += NQS_ROW_THREADS)
    {
      const cd vv = v[p];
#pragma unroll
      for (int r = 0; r < NQS_ROWS_PER_CTA; ++r)
      {
        const cd o = ld_stream(O+(k0+r)*P+p);
        ax[r] += o.x*vv.x-o.y*vv.y;
        ay[r] += o.x*vv.y+o.y*vv.x;
      }
    }
  }
  else
  {
    for (long long p = threadIdx.x; p < P; p += NQS_ROW_THREADS)
    {
      const cd vv = v[p];
      for (int r = 0; r < nrows; ++r)
      {
        const cd o = ld_stream(O+(k0+r)*P+p);
        ax[r] += o.x*vv.x-o.y*vv.y;
        ay[r] += o.x*vv.y+o.y*vv.x;
      }
    }
  }
  const int w = threadIdx.x>>5, lane = threadIdx.x&31;
#pragma unroll
  for (int r = 0; r < NQS_ROWS_PER_CTA; ++r)
  {
    const double sx = warp_sum(ax[r]), sy = warp_sum(ay[r]);
    if (lane == 0) { sh[r][0][w] = sx; sh[r][1][w] = sy; }
  }
  __syncthreads();
  if (threadIdx.x < NQS_ROWS_PER_CTA && threadIdx.x < nrows)
  {
    double sx = 0, sy = 0;
    for (int ww = 0; ww < NQS_ROW_THREADS/32; ++ww) { sx += sh[threadIdx.x][0][ww]; sy += sh[threadIdx.x][1][ww]; }
    zk[k0+threadIdx.x] = cmake(sx, sy);
  }
}

// Device-resident scalars of the preconditioned CG (csrc/cg_fused.cuh); the host only ever reads a copy.
struct CgScalars
{
  double rho, rho_old, alpha, beta, res2, thr, rhs2, tp;
  double aov_x, aov_y;     // <O> . p for the direction currently held in p
  double tol2;
  int done, iters, zero_rhs, fixed;
  int peer_timeout;        // multi-GPU: a peer's flag did not arrive within 20 s
  int barrier_timeout;     // the software grid barrier gave up (CTAs not co-resident on a non-cooperative launch)
  int nonfinite;           // <h> was not finite: the solve and the update were skipped (ref optimizer.cuh:134-138)
  int pad_;
};

// ref: update_parameters (impl_neural_quantum_state.cuh:1300-1312) / FFNN__UpdateParameters__ (:1665-1690, un-transposes the W block)
// `sc` (may be null): scalars of the solve that produced dx.  The update is enqueued behind the solve without a host round trip,
// so the kernel itself declines when <h> was not finite (the reference stops before the update then, optimizer.cuh:134-138) or
// when `need_done` is set and the solve has not converged within the iterations enqueued so far (the host then finishes the
// solve and enqueues the update again).
__global__ void update_params_kernel(const int N, const int M, const int model, const long long P, const cd * __restrict__ dx,
  const double lr, cd * __restrict__ params, const CgScalars * __restrict__ sc, const int need_done)
{
  if (sc != nullptr && (sc->nonfinite || (need_done && !sc->done))) return;
  const long long NM = (long long)N*M;
  for (long long q = (long long)blockIdx.x*blockDim.x+threadIdx.x; q < P; q += (long long)gridDim.x*blockDim.x)
  {
    long long dst = q;
    if (model == MODEL_FFNN && q < NM)
    { // dx index q = j*N+i  ->  W1[i*M+j]
      const long long j = q/N, i = q-j*N;
      dst = i*M+j;
    }
    const cd d = dx[q];
    cd v = params[dst];
    v.x -= lr*d.x; v.y -= lr*d.y;
    params[dst] = v;
  }
}
} // namespace nqs
