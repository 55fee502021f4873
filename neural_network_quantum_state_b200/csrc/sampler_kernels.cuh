// Generic ("direct log cosh") kernels of the sampler side of the path: tiled theta init, Metropolis sweep with the
// chain state resident in shared memory, local energy, single-flip forward.  They accept any N, M and both ansaetze and
// follow the reference arithmetic literally; the specialised register-resident kernels live in fast_kernels.cuh.
//
// Mapping: ONE WARP PER CHAIN.  theta_k (M complex) and s_k (N int8) of the warp's chain sit in shared memory for the
// whole launch, W rows are read with coalesced 16-byte loads (W is [N][M] row-major, so the "column of W^T" needed by
// a flip of site i is the contiguous row i), sum_j is a strided per-lane sum followed by a shuffle butterfly.
#pragma once
#include "device_math.cuh"

namespace nqs
{
enum { MODEL_RBM = 0, MODEL_FFNN = 1 };

struct ModelPtrs
{
  const cd * W;    // [N][M]
  const cd * a;    // RBM: visible bias [N]; FFNN: unused
  const cd * b;    // hidden bias [M]
  const cd * w1o;  // FFNN: output weights [M]; RBM: unused
};

__host__ __device__ inline ModelPtrs model_ptrs(int model, const cd * params, int N, int M)
{
  ModelPtrs p;
  p.W = params;
  if (model == MODEL_RBM) { p.a = params+(size_t)N*M; p.b = p.a+N; p.w1o = nullptr; }
  else { p.a = nullptr; p.b = params+(size_t)N*M; p.w1o = p.b+M; }
  return p;
}

// sum_j f_j(theta_j - two_s*W_ij): the body of ref k3 (impl_neural_quantum_state.cuh:1264-1278) + c1 (:102 / :828),
// warp-reduced.  f = logcosh (RBM) or w1o_j*logcosh (FFNN).
template <int MODEL>
__device__ __forceinline__ cd flip_sum(const cd * th, const cd * __restrict__ Wrow, const cd * __restrict__ w1o,
  const int M, const double two_s, const int lane)
{
  cd acc = cmake(0.0, 0.0);
  for (int j = lane; j < M; j += 32)
  {
    const cd w = Wrow[j], t = th[j];
    const cd lc = c_logcosh(cmake(t.x-w.x*two_s, t.y-w.y*two_s));
    if (MODEL == MODEL_RBM) acc = cadd(acc, lc);
    else acc = cadd(acc, cmul(w1o[j], lc));
  }
  return warp_sum(acc);
}

struct SweepArgs
{
  int N, M, model;
  long long K;
  const cd * params;
  int8_t * spins;      // [K][N]
  cd * theta;          // [K][M]
  cd * lnpsi0;         // [K]
  cd * sa;             // [K] (RBM)
  unsigned char * fresh; // [K] set to 1 when the chain accepted at least once (lnpsi0 then equals lnpsi(theta))
  const int * order;   // [N]
  int pos0;            // index into order of the first site to visit
  long long nsteps;    // proposals per chain in this launch
  const double * uniforms; // pre-drawn [nsteps][K] (already offset to the first step) or nullptr
  unsigned long long seed, step0;
  long long chain_offset;
  unsigned char * acc_log; // [nsteps][K] or nullptr
};

inline size_t sweep_smem_bytes(int N, int M, int warps)
{
  const size_t npad = (size_t)((N+15)/16)*16;
  return (size_t)warps*M*sizeof(cd)+(size_t)warps*npad+(size_t)N*sizeof(int);
}

// ref: BaseParallelSampler::do_mcmc_steps (impl_mcmc_sampler.cuh:28-39) with sampling_/accept_next_state_ of LITFIChain
// (impl_hamiltonians.cuh:207-218) -- the reference's 7 launches per proposal (k3,k4,c1,k5,k6,k7,k8,k9) fused into one launch
// per call, chain state resident on chip.
template <int MODEL>
__global__ void __launch_bounds__(256) sweep_generic_kernel(const SweepArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  const int N = a.N, M = a.M;
  const int npad = ((N+15)/16)*16;
  cd * th = reinterpret_cast<cd*>(smem_raw)+(size_t)w*M;
  int8_t * sp = reinterpret_cast<int8_t*>(reinterpret_cast<cd*>(smem_raw)+(size_t)warps*M)+(size_t)w*npad;
  int * ord = reinterpret_cast<int*>(reinterpret_cast<int8_t*>(reinterpret_cast<cd*>(smem_raw)+(size_t)warps*M)+(size_t)warps*npad);
  for (int i = threadIdx.x; i < N; i += blockDim.x)
    ord[i] = a.order[i];
  __syncthreads();
  const long long k = (long long)blockIdx.x*warps+w;
  if (k >= a.K) return;
  const ModelPtrs mp = model_ptrs(MODEL, a.params, N, M);
  for (int j = lane; j < M; j += 32) th[j] = a.theta[k*M+j];
  for (int i = lane; i < N; i += 32) sp[i] = a.spins[k*N+i];
  cd ln0 = a.lnpsi0[k];
  cd sa = (MODEL == MODEL_RBM) ? a.sa[k] : cmake(0.0, 0.0);
  __syncwarp();
  int pos = a.pos0;
  double ubuf = 0.0;
  bool any_acc = false;
  for (long long t = 0; t < a.nsteps; ++t)
  {
    if ((t&31) == 0)
    { // each lane fetches the uniform of one of the next 32 proposals
      const long long tt = t+lane;
      if (tt < a.nsteps)
        ubuf = a.uniforms ? a.uniforms[tt*a.K+k] : philox_uniform(a.seed, (unsigned long long)(a.chain_offset+k), a.step0+(unsigned long long)tt);
    }
    const double u = __shfl_sync(0xffffffffu, ubuf, (int)(t&31));
    const int site = ord[pos];
    pos = (pos+1 == N) ? 0 : pos+1;
    const double sig = (double)sp[site], two_s = 2.0*sig;
    const cd * Wrow = mp.W+(size_t)site*M;
    cd ln1 = flip_sum<MODEL>(th, Wrow, mp.w1o, M, two_s, lane);
    cd da = cmake(0.0, 0.0);
    if (MODEL == MODEL_RBM)
    { // ref k4 RBM__sadot__ (:1391-1404): lnpsi' = sa - 2 s a_i, then += sum_j
      const cd ai = mp.a[site];
      da = cmake(two_s*ai.x, two_s*ai.y);
      ln1 = cadd(csub(sa, da), ln1);
    }
    // ref k6 Sampler__ParallelMetropolisUpdate__ (impl_mcmc_sampler.cuh:75-102)
    const double d = ln1.x-ln0.x;
    const double ratio = exp(2.0*((d < 0) ? 1.0 : 0.0)*d);
    const bool acc = (u < ratio);
    const double delta = acc ? 1.0 : 0.0;
    ln0 = cmake(ln0.x+delta*(ln1.x-ln0.x), ln0.y+delta*(ln1.y-ln0.y));
    if (a.acc_log && lane == 0) a.acc_log[t*a.K+k] = acc ? 1 : 0;
    if (acc)
    { // ref k7 conditional_y_update (:1314-1329), k8 RBM__saUpdate__ (:1451-1465), k9 conditional_spin_update (:1353-1368)
      for (int j = lane; j < M; j += 32)
      {
        const cd wv = Wrow[j];
        cd tv = th[j];
        tv.x -= wv.x*two_s; tv.y -= wv.y*two_s;
        th[j] = tv;
      }
      sa = csub(sa, da);
      any_acc = true;
      if (lane == 0) sp[site] = (int8_t)(-sp[site]);
    }
    __syncwarp();
  }
  for (int j = lane; j < M; j += 32) a.theta[k*M+j] = th[j];
  for (int i = lane; i < N; i += 32) a.spins[k*N+i] = sp[i];
  if (lane == 0)
  {
    a.lnpsi0[k] = ln0;
    if (MODEL == MODEL_RBM) a.sa[k] = sa;
    if (any_acc) a.fresh[k] = 1;
  }
}

struct ElocArgs
{
  int N, M, model;
  long long K;
  const cd * params;
  const int8_t * spins;
  const cd * theta;
  const cd * lnpsi0;
  const cd * sa;
  const double * Jmat; // [N][N]
  const double * sjs;  // [K] sum_ij s_i J_ij s_j from the tensor-core GEMM (sv_struct.cuh), or null: computed here
  double hfield;
  cd * htilda;         // [K]
  cd * lnpsi1;         // optional [K]: lnpsi' of flipping `single_site` (forward(int) for tests); nullptr in E_loc mode
  int single_site;
};

// ref: LITFIChain::get_htilda_ (impl_hamiltonians.cuh:220-241): c5 Zgemm SJ = J s, k10 diag, N x (forward(i) + k11), k12 -- one launch.
// With lnpsi1 != nullptr the kernel instead evaluates a single forward(int) (ref :93-104).
template <int MODEL>
__global__ void __launch_bounds__(256) eloc_generic_kernel(const ElocArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x>>5, w = threadIdx.x>>5, lane = threadIdx.x&31;
  const int N = a.N, M = a.M;
  const int npad = ((N+15)/16)*16;
  cd * th = reinterpret_cast<cd*>(smem_raw)+(size_t)w*M;
  int8_t * sp = reinterpret_cast<int8_t*>(reinterpret_cast<cd*>(smem_raw)+(size_t)warps*M)+(size_t)w*npad;
  const long long k = (long long)blockIdx.x*warps+w;
  if (k >= a.K) return;
  const ModelPtrs mp = model_ptrs(MODEL, a.params, N, M);
  for (int j = lane; j < M; j += 32) th[j] = a.theta[k*M+j];
  for (int i = lane; i < N; i += 32) sp[i] = a.spins[k*N+i];
  const cd ln0 = a.lnpsi0[k];
  const cd sa = (MODEL == MODEL_RBM) ? a.sa[k] : cmake(0.0, 0.0);
  __syncwarp();
  if (a.lnpsi1 != nullptr)
  {
    const int site = a.single_site;
    const double two_s = 2.0*(double)sp[site];
    cd ln1 = flip_sum<MODEL>(th, mp.W+(size_t)site*M, mp.w1o, M, two_s, lane);
    if (MODEL == MODEL_RBM)
    {
      const cd ai = mp.a[site];
      ln1 = cadd(csub(sa, cmake(two_s*ai.x, two_s*ai.y)), ln1);
    }
    if (lane == 0) a.lnpsi1[k] = ln1;
    return;
  }
  // 1/2 sum_ij s_i J_ij s_j: lanes over i
  double diag = 0.0;
  for (int i = lane; i < N && a.sjs == nullptr; i += 32)
  {
    const double * Jrow = a.Jmat+(size_t)i*N;
    double sj = 0.0;
    for (int j = 0; j < N; ++j)
      sj = fma(Jrow[j], (double)sp[j], sj);
    diag = fma(sj, (double)sp[i], diag);
  }
  diag = 0.5*((a.sjs != nullptr) ? a.sjs[k] : warp_sum(diag));
  cd hsum = cmake(diag, 0.0);
  for (int site = 0; site < N; ++site)
  {
    const double two_s = 2.0*(double)sp[site];
    cd ln1 = flip_sum<MODEL>(th, mp.W+(size_t)site*M, mp.w1o, M, two_s, lane);
    if (MODEL == MODEL_RBM)
    {
      const cd ai = mp.a[site];
      ln1 = cadd(csub(sa, cmake(two_s*ai.x, two_s*ai.y)), ln1);
    }
    // ref k11 TFI__GetOffDiagElem__ (:857-869): htilda += h exp(lnpsi1 - lnpsi0)
    const cd e = c_exp(csub(ln1, ln0));
    hsum.x = fma(a.hfield, e.x, hsum.x);
    hsum.y = fma(a.hfield, e.y, hsum.y);
  }
  if (lane == 0) a.htilda[k] = cscale(hsum, 1.0/(double)N); // ref k12 (:240)
}

struct ThetaArgs
{
  int N, M, model;
  long long K;
  const cd * params;
  const int8_t * spins;     // spins the amplitudes are evaluated on [K][N]
  const int8_t * sa_spins;  // spins used for the visible-bias term (member spins; see the forward(spins) quirk, ref :119-120)
  cd * theta;               // out [K][M] (may be nullptr)
  cd * sa;                  // out [K] (RBM; may be nullptr)
  cd * lnpsi;               // out [K] (may be nullptr)
};

// ref: RBM::initialize / forward(spins) / tail of update_variables (impl_neural_quantum_state.cuh:67-91,107-129,161-169):
// theta = W^T s + b (c2+c3), sa = a.s (c4), lnpsi = sum_j logcosh(theta_j) + sa (k2 + c1).  FFNN :799-847.
// The reference runs this as Zgemm on spins stored as complex numbers.  Spins are +-1, so theta is a SIGNED ACCUMULATION of W
// rows: a CTA takes NQS_TH_CH chains, thread t owns hidden unit j = t (+256, ...) for all of them, so every W_ij fetched (one
// coalesced 16-byte load per thread) is used for NQS_TH_CH chains and the spins come from shared memory as broadcast doubles;
// log cosh and the sum over j are fused behind the accumulation (no theta round trip through HBM for lnpsi).
#define NQS_TH_THREADS 256
#define NQS_TH_CH 16

template <int MODEL, bool LNPSI>
__global__ void __launch_bounds__(NQS_TH_THREADS) theta_tiled_kernel(const ThetaArgs a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = a.N, M = a.M;
  double * sp = reinterpret_cast<double*>(smem_raw);                    // [N][NQS_TH_CH] spins as doubles (0 for missing chains)
  cd * red = reinterpret_cast<cd*>(sp+(size_t)N*NQS_TH_CH);              // [NQS_TH_THREADS/32][NQS_TH_CH]
  const int t = threadIdx.x, lane = t&31, w = t>>5;
  const long long kbase = (long long)blockIdx.x*NQS_TH_CH;
  const int nk = (int)((a.K-kbase < NQS_TH_CH) ? a.K-kbase : NQS_TH_CH);
  const ModelPtrs mp = model_ptrs(MODEL, a.params, N, M);
  for (int idx = t; idx < N*NQS_TH_CH; idx += NQS_TH_THREADS)
  {
    const int i = idx/NQS_TH_CH, c = idx-i*NQS_TH_CH;
    sp[idx] = (c < nk) ? (double)a.spins[(kbase+c)*N+i] : 0.0;
  }
  __syncthreads();
  cd lsum[LNPSI ? NQS_TH_CH : 1];
#pragma unroll
  for (int c = 0; c < (LNPSI ? NQS_TH_CH : 1); ++c) lsum[c] = cmake(0.0, 0.0);
  for (int j = t; j < M; j += NQS_TH_THREADS)
  {
    cd acc[NQS_TH_CH];
    const cd bj = mp.b[j];
#pragma unroll
    for (int c = 0; c < NQS_TH_CH; ++c) acc[c] = bj;
    for (int i = 0; i < N; ++i)
    {
      const cd wv = mp.W[(size_t)i*M+j];
      const double2 * srow = reinterpret_cast<const double2*>(sp+(size_t)i*NQS_TH_CH);
#pragma unroll
      for (int c2 = 0; c2 < NQS_TH_CH/2; ++c2)
      {
        const double2 s2 = srow[c2];
        acc[2*c2].x = fma(s2.x, wv.x, acc[2*c2].x); acc[2*c2].y = fma(s2.x, wv.y, acc[2*c2].y);
        acc[2*c2+1].x = fma(s2.y, wv.x, acc[2*c2+1].x); acc[2*c2+1].y = fma(s2.y, wv.y, acc[2*c2+1].y);
      }
    }
    cd wo = cmake(1.0, 0.0);
    if (MODEL == MODEL_FFNN && LNPSI) wo = mp.w1o[j];
#pragma unroll
    for (int c = 0; c < NQS_TH_CH; ++c)
    {
      if (c < nk)
      {
        if (a.theta) a.theta[(kbase+c)*M+j] = acc[c];
        if (LNPSI)
        {
          const cd lc = c_logcosh(acc[c]);
          lsum[LNPSI ? c : 0] = cadd(lsum[LNPSI ? c : 0], (MODEL == MODEL_RBM) ? lc : cmul(wo, lc));
        }
      }
    }
  }
  if (LNPSI)
  {
#pragma unroll
    for (int c = 0; c < (LNPSI ? NQS_TH_CH : 1); ++c)
    {
      const cd s = warp_sum(lsum[c]);
      if (lane == 0) red[w*NQS_TH_CH+c] = s;
    }
  }
  __syncthreads();
  // visible-bias term (RBM) from sa_spins, and the final lnpsi: warp w finishes chains w, w+8
  for (int c = w; c < nk; c += NQS_TH_THREADS/32)
  {
    cd sa = cmake(0.0, 0.0);
    if (MODEL == MODEL_RBM)
    {
      for (int i = lane; i < N; i += 32)
      {
        const double s = (double)a.sa_spins[(kbase+c)*N+i];
        const cd ai = mp.a[i];
        sa.x = fma(s, ai.x, sa.x);
        sa.y = fma(s, ai.y, sa.y);
      }
      sa = warp_sum(sa);
      if (a.sa && lane == 0) a.sa[kbase+c] = sa;
    }
    if (LNPSI && lane == 0)
    {
      cd tot = sa;
      for (int ww = 0; ww < NQS_TH_THREADS/32; ++ww) tot = cadd(tot, red[ww*NQS_TH_CH+c]);
      a.lnpsi[kbase+c] = tot;
    }
  }
}

// ref: the all-true accept_next_state_ of warm_up (impl_mcmc_sampler.cuh:21-22) -> Ansatz::spin_flip(all, index_) (:172-182):
// theta -= 2 s W_i, sa -= 2 s a_i, s_i = -s_i on every chain; lnpsi0 untouched.
__global__ void flip_site_all_kernel(const int N, const int M, const long long K, const int model, const cd * params,
  int8_t * spins, cd * theta, cd * sa, const int site)
{
  const ModelPtrs mp = model_ptrs(model, params, N, M);
  const long long total = K*(long long)M;
  for (long long idx = (long long)blockIdx.x*blockDim.x+threadIdx.x; idx < total; idx += (long long)gridDim.x*blockDim.x)
  {
    const long long k = idx/M;
    const int j = (int)(idx-k*M);
    const double two_s = 2.0*(double)spins[k*N+site];
    const cd wv = mp.W[(size_t)site*M+j];
    cd t = theta[idx];
    t.x -= wv.x*two_s; t.y -= wv.y*two_s;
    theta[idx] = t;
  }
}
__global__ void flip_site_all_finish_kernel(const int N, const int M, const long long K, const int model, const cd * params,
  int8_t * spins, cd * sa, const int site)
{
  const ModelPtrs mp = model_ptrs(model, params, N, M);
  for (long long k = (long long)blockIdx.x*blockDim.x+threadIdx.x; k < K; k += (long long)gridDim.x*blockDim.x)
  {
    const double two_s = 2.0*(double)spins[k*N+site];
    if (model == MODEL_RBM)
    {
      const cd ai = mp.a[site];
      sa[k] = cmake(sa[k].x-two_s*ai.x, sa[k].y-two_s*ai.y);
    }
    spins[k*N+site] = (int8_t)(-spins[k*N+site]);
  }
}
} // namespace nqs
