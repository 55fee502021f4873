// libnqs_b200.so -- implementation of include/nqs_b200.h.  See DESIGN.md for the data layout and the kernel inventory.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <random>
#include <sstream>

#include "engine.h"
#include "sampler_kernels.cuh"
#include "sr_kernels.cuh"
#include "sv_fused.cuh"
#include "sv_struct.cuh"
#include "cg_fused.cuh"
#include "cg_persist.cuh"

using namespace nqs;

// ------------------------------------------------------------------------------------------------------------------
// NCCL through dlopen: the library must load (and the single-GPU path must run) without NCCL present, and inside a
// PyTorch process it must bind to the libnccl.so.2 that torch already loaded instead of pulling a second copy.
// ------------------------------------------------------------------------------------------------------------------
namespace
{
typedef struct { char internal[128]; } ncclUniqueIdT;
typedef int (*ncclGetUniqueId_t)(ncclUniqueIdT *);
typedef int (*ncclCommInitRank_t)(void **, int, ncclUniqueIdT, int);
typedef int (*ncclAllReduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*ncclCommDestroy_t)(void *);
typedef const char * (*ncclGetErrorString_t)(int);
struct NcclApi
{
  void * lib = nullptr;
  ncclGetUniqueId_t getUniqueId = nullptr;
  ncclCommInitRank_t commInitRank = nullptr;
  ncclAllReduce_t allReduce = nullptr;
  ncclCommDestroy_t commDestroy = nullptr;
  ncclGetErrorString_t getErrorString = nullptr;
  std::string why;
  bool load()
  {
    if (lib) return true;
    const char * names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char * n : names)
    {
      lib = dlopen(n, RTLD_NOW|RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { why = std::string("dlopen(libnccl.so.2) failed: ")+dlerror(); return false; }
    getUniqueId = (ncclGetUniqueId_t)dlsym(lib, "ncclGetUniqueId");
    commInitRank = (ncclCommInitRank_t)dlsym(lib, "ncclCommInitRank");
    allReduce = (ncclAllReduce_t)dlsym(lib, "ncclAllReduce");
    commDestroy = (ncclCommDestroy_t)dlsym(lib, "ncclCommDestroy");
    getErrorString = (ncclGetErrorString_t)dlsym(lib, "ncclGetErrorString");
    if (!getUniqueId || !commInitRank || !allReduce || !commDestroy)
    { why = "libnccl lacks required symbols"; lib = nullptr; return false; }
    return true;
  }
};
NcclApi g_nccl;
const int kNcclFloat64 = 8, kNcclSum = 0; // ncclDataType_t / ncclRedOp_t values of NCCL 2.x

thread_local std::string g_create_error;

template <typename F>
nqs_status guarded(nqs_handle * h, F && f)
{
  try { f(); return NQS_OK; }
  catch (const Error & e) { if (h) h->err = e.what(); else g_create_error = e.what(); return e.code; }
  catch (const std::bad_alloc &) { if (h) h->err = "host allocation failed"; return NQS_ERR_NOMEM; }
  catch (const std::exception & e) { if (h) h->err = e.what(); else g_create_error = e.what(); return NQS_ERR_INVALID; }
}

inline int grid_for(long long n, int threads, int cap)
{
  long long g = (n+threads-1)/threads;
  if (g < 1) g = 1;
  return (int)std::min<long long>(g, cap);
}

enum { TAG_SWEEP = 0, TAG_ELOC, TAG_ODERIV, TAG_SETUP, TAG_CG, TAG_UPDATE, TAG_ROWS, TAG_COLS };

int ev_next(nqs_handle * h)
{
  if (h->ev_used == h->evpool.size())
  {
    cudaEvent_t e;
    NQS_CUDA(cudaEventCreate(&e));
    h->evpool.push_back(e);
  }
  return (int)h->ev_used++;
}

struct Span
{ // CUDA-event bracket on the handle's stream, recorded without any synchronisation; resolve_spans() reads them later
  nqs_handle * h; int tag, b; bool on; cudaStream_t st;
  Span(nqs_handle * h_, int tag_, cudaStream_t st_ = nullptr): h(h_), tag(tag_), b(-1), on(h_->timing_on), st(st_ ? st_ : h_->stream)
  { if (on) { b = ev_next(h); cudaEventRecord(h->evpool[b], st); } }
  ~Span()
  {
    if (!on) return;
    const int e = ev_next(h);
    cudaEventRecord(h->evpool[e], st);
    h->spans.push_back({tag, b, e});
  }
};

void reset_phase_times(nqs_handle * h)
{
  nqs_timing & t = h->timing;
  t.sweep_ms = t.eloc_ms = t.oderiv_ms = t.setup_ms = t.cg_ms = t.update_ms = t.rows_ms = t.cols_ms = 0;
  t.rows_count = t.cols_count = 0;
}

void resolve_spans(nqs_handle * h)
{
  if (h->spans.empty()) { h->ev_used = 0; return; }
  NQS_CUDA(cudaStreamSynchronize(h->stream));
  nqs_timing & t = h->timing;
  std::vector<float> ms_of(h->spans.size());
  float max_rows = 0, max_cols = 0;
  for (size_t i = 0; i < h->spans.size(); ++i)
  {
    const nqs_handle::Span & sp = h->spans[i];
    float ms = 0;
    cudaEventElapsedTime(&ms, h->evpool[sp.b], h->evpool[sp.e]);
    ms_of[i] = ms;
    if (sp.tag == TAG_ROWS) max_rows = std::max(max_rows, ms);
    if (sp.tag == TAG_COLS) max_cols = std::max(max_cols, ms);
  }
  for (size_t i = 0; i < h->spans.size(); ++i)
  {
    const nqs_handle::Span & sp = h->spans[i];
    const float ms = ms_of[i];
    switch (sp.tag)
    {
      case TAG_SWEEP: t.sweep_ms += ms; break;
      case TAG_ELOC: t.eloc_ms += ms; break;
      case TAG_ODERIV: t.oderiv_ms += ms; break;
      case TAG_SETUP: t.setup_ms += ms; break;
      case TAG_CG: t.cg_ms += ms; break;
      case TAG_UPDATE: t.update_ms += ms; break;
      // launches issued after the CG converged return immediately (device-side `done` flag): they are not O passes
      case TAG_ROWS: if (ms >= 0.25f*max_rows) { t.rows_ms += ms; t.rows_count += 1; } break;
      case TAG_COLS: if (ms >= 0.25f*max_cols) { t.cols_ms += ms; t.cols_count += 1; } break;
    }
  }
  h->spans.clear();
  h->ev_used = 0;
}

void check_launch(nqs_handle * h, const char * what)
{
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    throw Error(NQS_ERR_CUDA, std::string("launch of ")+what+" failed: "+cudaGetErrorString(e));
  h->timing.kernel_launches += 1;
}

// the vector the optimiser moves: the parameters themselves, or the tied variables of the translation-symmetric RBM
cd * var_ptr(nqs_handle * h) { return h->trsymm ? h->vars.p : h->params.p; }
// vars -> expanded network (ref symmetrize_variables_, impl_neural_quantum_state.cuh:534-538); no-op for the plain ansaetze
void expand_vars(nqs_handle * h)
{
  if (!h->trsymm) return;
  const int grid = grid_for(h->Pfull, 256, 148*4);
  if (h->tied == TIED_RBM_Z2PR) z2pr_expand_kernel<<<grid, 256, 0, h->stream>>>(h->N, h->alpha_f, h->vars.p, h->params.p);
  else if (h->tied == TIED_FFNN_TR) ffnntr_expand_kernel<<<grid, 256, 0, h->stream>>>(h->N, h->alpha_f, h->vars.p, h->params.p);
  else trsymm_expand_kernel<<<grid, 256, 0, h->stream>>>(h->N, h->alpha_f, h->vars.p, h->params.p);
  check_launch(h, "trsymm_expand_kernel");
}

int warps_for_smem(const nqs_handle * h, size_t per_warp, size_t fixed)
{ // as many warps per CTA (<= 8) as fit the opt-in shared memory
  int w = 8;
  while (w > 1 && per_warp*w+fixed > h->smem_optin) w >>= 1;
  NQS_REQUIRE(per_warp*w+fixed <= h->smem_optin, NQS_ERR_UNSUPPORTED, "n_hiddens too large for the shared-memory resident chain state");
  return w;
}

template <typename Kern>
void set_smem(Kern kern, size_t bytes)
{
  if (bytes > 48*1024)
    NQS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

// ---- launches ------------------------------------------------------------------------------------------------------
// chains per CTA of the rows GEMM (8 MT): one CTA per SM, so the tile height decides how full the last wave is
int rows_dmma_mt(const nqs_handle * h)
{ // fewest waves x tile height over the tile heights that fit (ties: the taller tile, less L2 traffic)
  static const int cand[] = {8, 7, 6, 4, 2};
  int best = 2;
  long long best_cost = -1;
  for (const int mt : cand)
  {
    if (mt > 2 && rows_dmma_smem(h->N, mt) > h->smem_optin) continue;
    const long long ctas = (h->K+8*mt-1)/(8*mt), waves = (ctas+h->sm_count-1)/h->sm_count;
    const long long cost = waves*mt;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = mt; }
  }
  return best;
}

template <int MODEL, int EPI, int MT>
void launch_rows_dmma_t(nqs_handle * h, const RowsArgs & a)
{
  const size_t smem = rows_dmma_smem(h->N, MT);
  set_smem(spin_rows_dmma_kernel<MODEL, EPI, MT>, smem);
  spin_rows_dmma_kernel<MODEL, EPI, MT><<<(unsigned)((h->K+8*MT-1)/(8*MT)), NQS_DR_THREADS, smem, h->stream>>>(a);
}

template <int MODEL, int EPI>
void launch_rows_dmma(nqs_handle * h, const RowsArgs & a)
{
  const int mt = rows_dmma_mt(h);
  if (mt == 8) launch_rows_dmma_t<MODEL, EPI, 8>(h, a);
  else if (mt == 7) launch_rows_dmma_t<MODEL, EPI, 7>(h, a);
  else if (mt == 6) launch_rows_dmma_t<MODEL, EPI, 6>(h, a);
  else if (mt == 4) launch_rows_dmma_t<MODEL, EPI, 4>(h, a);
  else launch_rows_dmma_t<MODEL, EPI, 2>(h, a);
  check_launch(h, "spin_rows_dmma_kernel");
}

// tcgen05 int8 rows kernel (rows_umma.cuh): one CTA per 128 chains, so it only pays when a rank holds enough chains to cover
// most of the SMs; N <= 512 keeps the int64 recombination exact.  NQS_ROWS_UMMA=0/1 overrides.
bool rows_umma_ok(nqs_handle * h)
{
  if (h->rows_umma < 0)
  {
    const char * e = std::getenv("NQS_ROWS_UMMA");
    bool on = (h->K >= 64ll*128);
    if (e) on = (std::atoi(e) != 0);
    const int m2 = std::max(2*h->M, h->N);     // widest B: W / the W block of v (2M real columns), J (N columns)
    on = on && !(h->cfg.flags & NQS_FLAG_NO_DMMA) && h->N <= 512 && rows_umma_smem(h->N, h->M, m2, 2, 2) <= h->smem_optin;
    if (on)
    {
      const size_t bytes = (size_t)ru_nchunks(m2)*ru_chunk_bytes(h->N);
      h->bq.alloc(bytes);
      NQS_CUDA(cudaMemsetAsync(h->bq.p, 0, bytes, h->stream));
      h->bscale.alloc((size_t)m2);
    }
    h->rows_umma = on ? 1 : 0;
  }
  return h->rows_umma == 1;
}

template <int MODEL, int EPI>
void launch_rows_umma(nqs_handle * h, const RowsArgs & a)
{
  const int m2 = (EPI == ROWS_EPI_SJS) ? h->N : 2*h->M;
  ozaki_split_kernel<<<(unsigned)ru_nchunks(m2), 32*(ru_npad(h->N)/16), 0, h->stream>>>(h->N, m2, a.B, h->bq.p, h->bscale.p, EPI == ROWS_EPI_Z ? a.done : nullptr);
  check_launch(h, "ozaki_split_kernel");
  int nbuf, nt;
  rows_umma_plan(h->N, h->M, m2, EPI == ROWS_EPI_Z, h->smem_optin, nbuf, nt);
  const size_t smem = rows_umma_smem(h->N, h->M, m2, nbuf, nt);
  set_smem(spin_rows_umma_kernel<MODEL, EPI>, smem);
  // programmatic dependent launch: the CTAs start (TMEM allocation, spin tile) while the split kernel drains
  cudaLaunchConfig_t lc;
  std::memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)((h->K+127)/128)); lc.blockDim = dim3(NQS_RU_THREADS); lc.dynamicSmemBytes = smem; lc.stream = h->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = 1;
  const int8_t * bq = h->bq.p;
  const double * bs = h->bscale.p;
  static const int trace_cta = std::getenv("NQS_RU_TRACE") ? std::atoi(std::getenv("NQS_RU_TRACE")) : -1;
  if (const char * e = std::getenv("NQS_RU_NBUF")) { nbuf = std::max(2, std::min(std::atoi(e), NQS_RU_MAXBUF)); }
  if (const char * e = std::getenv("NQS_RU_NT")) { if (EPI == ROWS_EPI_Z) nt = std::max(2, std::min(std::atoi(e), 4)); }
  const size_t smem2 = rows_umma_smem(h->N, h->M, m2, nbuf, nt);
  if (smem2 != smem) { set_smem(spin_rows_umma_kernel<MODEL, EPI>, smem2); lc.dynamicSmemBytes = smem2; }
  NQS_CUDA(cudaLaunchKernelEx(&lc, spin_rows_umma_kernel<MODEL, EPI>, a, bq, bs, nbuf, nt, trace_cta));
  check_launch(h, "spin_rows_umma_kernel");
}

bool rows_dmma_ok(const nqs_handle * h)
{ // the spin tile of 16 chains and two slabs of B must fit the shared memory of one CTA (N up to ~700)
  return !(h->cfg.flags & NQS_FLAG_NO_DMMA) && rows_dmma_smem(h->N, 2) <= h->smem_optin;
}

// sum_ij s_i J_ij s_j of every chain as the GEMM (S J) . S on the fp64 tensor cores (ref c5 Zgemm + k10, impl_hamiltonians.cuh:226-231).
// Returns the device vector, or null (odd N, very long chains, NQS_FLAG_NO_DMMA): the local-energy kernels then form it themselves.
const double * launch_sjs(nqs_handle * h)
{
  if (!rows_dmma_ok(h) || h->N%2 != 0) return nullptr;
  if (h->sjs.p == nullptr) h->sjs.alloc((size_t)h->K);
  RowsArgs a;
  std::memset(&a, 0, sizeof(a));
  a.N = h->N; a.M = h->M; a.K = h->K; a.spins = h->spins.p; a.B = h->Jmat.p; a.sjs = h->sjs.p;
  if (rows_umma_ok(h)) launch_rows_umma<MODEL_RBM, ROWS_EPI_SJS>(h, a); else launch_rows_dmma<MODEL_RBM, ROWS_EPI_SJS>(h, a);
  return h->sjs.p;
}

// theta = S W + b (+ sa, + lnpsi): fp64 tensor-core GEMM with the log cosh row sum fused into its epilogue (sv_struct.cuh);
// theta_tiled_kernel (scalar FMAs) for very long chains or on request
void launch_theta(nqs_handle * h, const int8_t * spins_dev, const int8_t * sa_spins_dev, cd * theta, cd * sa, cd * lnpsi)
{
  if (rows_dmma_ok(h))
  {
    const ModelPtrs mp = model_ptrs(h->model, h->params.p, h->N, h->M);
    RowsArgs a;
    std::memset(&a, 0, sizeof(a));
    a.N = h->N; a.M = h->M; a.K = h->K; a.spins = spins_dev; a.B = reinterpret_cast<const double*>(mp.W); a.bias = mp.b;
    a.theta = theta; a.sa_spins = sa_spins_dev; a.avis = mp.a; a.w1o = mp.w1o; a.sa = sa; a.lnpsi = lnpsi;
    if (rows_umma_ok(h))
    {
      if (h->model == MODEL_RBM)
      { if (lnpsi) launch_rows_umma<MODEL_RBM, ROWS_EPI_LNPSI>(h, a); else launch_rows_umma<MODEL_RBM, ROWS_EPI_THETA>(h, a); }
      else
      { if (lnpsi) launch_rows_umma<MODEL_FFNN, ROWS_EPI_LNPSI>(h, a); else launch_rows_umma<MODEL_FFNN, ROWS_EPI_THETA>(h, a); }
      h->variant_theta = "umma_i8_ozaki7_rows";
      return;
    }
    if (h->model == MODEL_RBM)
    { if (lnpsi) launch_rows_dmma<MODEL_RBM, ROWS_EPI_LNPSI>(h, a); else launch_rows_dmma<MODEL_RBM, ROWS_EPI_THETA>(h, a); }
    else
    { if (lnpsi) launch_rows_dmma<MODEL_FFNN, ROWS_EPI_LNPSI>(h, a); else launch_rows_dmma<MODEL_FFNN, ROWS_EPI_THETA>(h, a); }
    h->variant_theta = "dmma_rows";
    return;
  }
  ThetaArgs a;
  a.N = h->N; a.M = h->M; a.model = h->model; a.K = h->K; a.params = h->params.p;
  a.spins = spins_dev; a.sa_spins = sa_spins_dev; a.theta = theta; a.sa = sa; a.lnpsi = lnpsi;
  const size_t smem = (size_t)h->N*NQS_TH_CH*sizeof(double)+(size_t)(NQS_TH_THREADS/32)*NQS_TH_CH*sizeof(cd);
  const int grid = (int)((h->K+NQS_TH_CH-1)/NQS_TH_CH);
#define NQS_TH_LAUNCH(MODEL_, LN_) do { set_smem(theta_tiled_kernel<MODEL_, LN_>, smem); \
    theta_tiled_kernel<MODEL_, LN_><<<grid, NQS_TH_THREADS, smem, h->stream>>>(a); } while (0)
  if (h->model == MODEL_RBM) { if (lnpsi) NQS_TH_LAUNCH(MODEL_RBM, true); else NQS_TH_LAUNCH(MODEL_RBM, false); }
  else { if (lnpsi) NQS_TH_LAUNCH(MODEL_FFNN, true); else NQS_TH_LAUNCH(MODEL_FFNN, false); }
#undef NQS_TH_LAUNCH
  check_launch(h, "theta_tiled_kernel");
  h->variant_theta = "tiled_fma";
}


// ---- specialised RBM path (fast_kernels.cuh) -------------------------------------------------------------------------
// pinned staging area (4096 B): [0,64) scalar read-backs | [512,520) theta bound | [1024,..) CG scalars read back | [2048,..) CG scalars upload
enum { PIN_HS = 0, PIN_BOUND = 512, PIN_CG_SNAP = 1024, PIN_CG_INIT = 2048 };

// Tables of the product-form kernels for the CURRENT parameters, plus the bound that decides whether those kernels may be used.
// The bound travels to pinned memory asynchronously and is picked up at the caller's next synchronisation (finish_tables), so
// a parameter update inside nqs_sr_step costs no host round trip of its own.
void build_tables_async(nqs_handle * h)
{
  if (h->jpl == 0 || h->tables_valid || h->bound_inflight) return;
  build_fast_tables_kernel<<<grid_for((long long)h->N*h->mpad, 256, 148*8), 256, 0, h->stream>>>(h->N, h->M, h->mpad, h->params.p,
    h->ftab_a.p, h->ftab_b.p, h->ctab_a.p, h->ctab_b.p, h->ctabT_a.p, h->ctabT_b.p, h->npad32, h->w2.p, h->afac.p, h->aexp.p, h->bound.p,
    h->model == MODEL_RBM ? 1 : 0, h->ftab32.p);
  check_launch(h, "build_fast_tables_kernel");
  if (h->model == MODEL_FFNN)
  { // the FNN kernels hold tanh(theta) and log f, nothing that could overflow: no bound to wait for
    h->theta_bound = 0.0;
    h->tables_valid = true;
    return;
  }
  theta_bound_kernel<<<(h->M+31)/32, NQS_TB_THREADS, 0, h->stream>>>(h->N, h->M, h->params.p, h->bound.p);
  check_launch(h, "theta_bound_kernel");
  NQS_CUDA(cudaMemcpyAsync((char*)h->pinned+PIN_BOUND, h->bound.p, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  h->bound_inflight = true;
}
void finish_tables(nqs_handle * h)
{ // call after a synchronisation of the stream
  if (!h->bound_inflight) return;
  std::memcpy(&h->theta_bound, (char*)h->pinned+PIN_BOUND, sizeof(double));
  h->bound_inflight = false;
  h->tables_valid = true;
}
void invalidate_tables(nqs_handle * h)
{ // the parameters are about to change
  if (h->bound_inflight) { NQS_CUDA(cudaStreamSynchronize(h->stream)); h->bound_inflight = false; }
  h->tables_valid = false;
}
void ensure_tables(nqs_handle * h)
{
  if (h->tables_valid || h->jpl == 0) return;
  build_tables_async(h);
  NQS_CUDA(cudaStreamSynchronize(h->stream));
  finish_tables(h);
}

bool fast_path_ok(nqs_handle * h)
{ // product form is overflow-free while JPL * 2 max|Re theta| stays far below log(DBL_MAX)
  if (h->jpl == 0) return false;
  ensure_tables(h);
  return std::isfinite(h->theta_bound) && h->theta_bound < 300.0/h->jpl;
}

// Chains per warp of the register-resident sweep.  More chains per warp reuse every table row fetched from shared memory and
// amortise the per-proposal bookkeeping (the throughput choice: 4 for <= 4 hidden-unit slots per lane, 2 for 8, 1 above), but a
// warp then walks its chains' proposals one after the other.  When the whole shard fits the SMs at once with FEWER chains per
// warp (a strong-scaled shard: 2048 chains per GPU at 8 GPUs), the extra warps run in parallel instead and the sweep gets
// shorter (0.32 -> ~0.2 ms there): take the smallest C whose grid is a single wave, else the throughput choice.
int sweep_chains_per_warp(const nqs_handle * h)
{
  const int cmax = (h->jpl <= 4) ? 4 : (h->jpl == 8 ? 2 : 1);
  { const char * e = std::getenv("NQS_SWEEP_C"); if (e) { const int c = std::atoi(e); if (c == 1 || (c == 2 && cmax >= 2) || (c == 4 && cmax == 4)) return c; } }
  if (h->jpl > 8) return 1;
  for (int c = 1; c < cmax; c <<= 1)
  {
    const int warps = (c == 1) ? 8 : 4;                      // SweepShape: one chain per warp runs 8 warps per CTA, else 4; 2 CTAs per SM
    const long long ctas = (h->K+(long long)warps*c-1)/((long long)warps*c);
    if (ctas <= 2LL*h->sm_count) return c;
  }
  return cmax;
}

// shared memory of the sweep variant launch_sweep picks (long chains with a narrow hidden layer can exceed the opt-in limit, and
// then the generic kernel takes over)
bool fast_sweep_fits(const nqs_handle * h)
{
  const int C = sweep_chains_per_warp(h);
  const int warps = (h->jpl > 16) ? SweepShape<16, 1, 2>::warps : ((C == 1 && h->jpl <= 8) ? 8 : 4);
  return fast_sweep_smem_bytes(h->N, C, warps, h->mpad) <= h->smem_optin;
}

template <int JPL, int C, int WPC = 1>
void launch_sweep_fast_t(nqs_handle * h, const FastSweepArgs & a)
{
  const int warps = SweepShape<JPL, C, WPC>::warps;
  const size_t smem = fast_sweep_smem_bytes(h->N, C, warps, h->mpad);
  set_smem(rbm_sweep_fast_kernel<JPL, C, WPC>, smem);
  const long long per_cta = (long long)(warps/WPC)*C;
  rbm_sweep_fast_kernel<JPL, C, WPC><<<(unsigned)((h->K+per_cta-1)/per_cta), warps*32, smem, h->stream>>>(a);
}

template <int C>
void launch_eloc_sites_t(nqs_handle * h, const FastElocArgs & a)
{
  const int nwarps = std::max(1, std::min(8, (h->N+31)/32));
  const size_t smem = fast_eloc_smem_bytes(h->N, h->M, C);
  set_smem(rbm_eloc_sites_kernel<C>, smem);
  rbm_eloc_sites_kernel<C><<<(unsigned)((h->K+C-1)/C), nwarps*32, smem, h->stream>>>(a);
}

void launch_eloc_fast(nqs_handle * h)
{
  FastElocArgs a;
  a.N = h->N; a.M = h->M; a.Npad = h->npad32; a.K = h->K; a.ctabT_a = h->ctabT_a.p; a.ctabT_b = h->ctabT_b.p; a.aexp = h->aexp.p;
  a.spins = h->spins.p; a.theta = h->theta.p; a.lnpsi0 = h->lnpsi0.p; a.sa = h->sa.p; a.fresh = h->fresh.p; a.Jmat = h->Jmat.p;
  a.hfield = h->cfg.h; a.htilda = h->htilda.p;
  a.sjs = launch_sjs(h);
  launch_eloc_sites_t<4>(h, a);     // chains per CTA: 2 -> 0.77 ms, 4 -> 0.59 ms, 8 -> 0.62 ms at cfg3
  check_launch(h, "rbm_eloc_sites_kernel");
}

// ---- trng::yarn2 feed (yarn2.cuh) --------------------------------------------------------------------------------------
constexpr size_t YARN2_FEED_BYTES = (size_t)256<<20;   // uniforms generated per sweep launch (whole sweeps)
long long yarn2_steps_per_launch(const nqs_handle * h)
{
  const long long sweeps = (long long)(YARN2_FEED_BYTES/sizeof(double))/(h->K*h->N);
  return std::max(1ll, sweeps)*h->N;
}
const double * yarn2_fill(nqs_handle * h, long long nsteps)
{
  if (h->yarn_tab.p == nullptr)
  {
    h->yarn_tab.alloc((size_t)YARN2_TAB0+YARN2_TAB1);
    yarn2_table_kernel<<<(YARN2_TAB0+YARN2_TAB1+255)/256, 256, 0, h->stream>>>(h->yarn_tab.p);
    check_launch(h, "yarn2_table_kernel");
  }
  if (h->yarn_state.n < (size_t)h->K) { h->yarn_state.alloc((size_t)h->K); h->yarn_state_at = -1; }
  if (h->yarn_u.n < (size_t)nsteps*h->K)
  { // the previous launch may still read the old buffer
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    h->yarn_u.alloc((size_t)nsteps*h->K);
  }
  Yarn2FillArgs y;
  y.K = h->K; y.chain_offset = h->koff; y.seed = h->cfg.seed; y.seed_distance = h->seed_distance; y.draws_done = h->step_counter;
  y.nsteps = nsteps; y.rebuild = (h->yarn_state_at != (long long)h->step_counter) ? 1 : 0;
  y.tab = h->yarn_tab.p; y.state = h->yarn_state.p; y.u = h->yarn_u.p;
  yarn2_fill_kernel<<<(unsigned)((h->K+127)/128), 128, 0, h->stream>>>(y);
  check_launch(h, "yarn2_fill_kernel");
  h->yarn_state_at = (long long)(h->step_counter+(unsigned long long)nsteps);
  return h->yarn_u.p;
}

void launch_sweep(nqs_handle * h, long long nsteps)
{
  if (nsteps <= 0) return;
  h->theta_matches_O = false; h->hidden_valid = false; h->o_pending = false;
  SweepArgs a;
  a.N = h->N; a.M = h->M; a.model = h->model; a.K = h->K; a.params = h->params.p;
  a.spins = h->spins.p; a.theta = h->theta.p; a.lnpsi0 = h->lnpsi0.p; a.sa = h->sa.p; a.fresh = h->fresh.p; a.order = h->order.p;
  a.pos0 = h->pos; a.nsteps = nsteps;
  a.uniforms = nullptr;
  if (h->u_steps > 0)
  {
    NQS_REQUIRE(h->u_used+nsteps <= h->u_steps, NQS_ERR_STATE, "pre-drawn uniform feed exhausted: call nqs_set_uniforms with enough steps");
    a.uniforms = (h->u_zc ? h->u_zc : h->uniforms.p)+(size_t)h->u_used*h->K;
  }
  else if (h->rng_kind == NQS_RNG_YARN2)
  { // the reference's trng::yarn2 stream (yarn2.cuh): the uniforms of this launch are generated into a feed buffer first
    const long long cap = yarn2_steps_per_launch(h);
    if (nsteps > cap && !(h->cfg.flags & NQS_FLAG_ACCEPT_LOG))
    { // bounded buffer: whole sweeps at a time
      for (long long done = 0; done < nsteps; done += cap) launch_sweep(h, std::min(cap, nsteps-done));
      return;
    }
    a.uniforms = yarn2_fill(h, nsteps);
  }
  a.seed = h->cfg.seed; a.step0 = h->step_counter; a.chain_offset = h->koff;
  a.acc_log = nullptr;
  if (h->cfg.flags & NQS_FLAG_ACCEPT_LOG)
  {
    if ((long long)h->acc_log.n < nsteps*h->K) h->acc_log.alloc((size_t)nsteps*h->K);
    a.acc_log = h->acc_log.p;
    h->acc_log_steps = nsteps;
  }
  if (h->model == MODEL_FFNN && fast_path_ok(h) && h->jpl <= 16 && nsteps%h->N == 0)
  { // hidden-unit state resident on chip, one complex log per (proposal, hidden unit): ffnn_fast_kernels.cuh
    // as many chains (warps) per CTA as the shared memory holds: the kernel is latency-bound (one dependent reduction per proposal)
    int warps = NQS_FF_MAX_WARPS;
    while (warps > 1 && ffnn_sweep_smem_bytes(h->N, warps, h->mpad) > h->smem_optin) --warps;
    { const char * e = std::getenv("NQS_FF_WARPS"); if (e && std::atoi(e) >= 1 && std::atoi(e) <= warps) warps = std::atoi(e); }
    if (ffnn_sweep_smem_bytes(h->N, warps, h->mpad) <= h->smem_optin)
    {
      FfnnSweepArgs f;
      f.N = h->N; f.M = h->M; f.Mpad = h->mpad; f.K = h->K; f.params = h->params.p; f.ctab_a = h->ctab_a.p; f.ctab_b = h->ctab_b.p; f.w2 = h->w2.p;
      f.spins = h->spins.p; f.theta = h->theta.p; f.lnpsi0 = h->lnpsi0.p; f.fresh = h->fresh.p; f.order = h->order.p;
      f.pos0 = h->pos; f.nsweeps = (int)(nsteps/h->N); f.uniforms = a.uniforms; f.seed = a.seed; f.step0 = a.step0;
      f.chain_offset = a.chain_offset; f.acc_log = a.acc_log;
      const size_t smem = ffnn_sweep_smem_bytes(h->N, warps, h->mpad);
      const unsigned grid = (unsigned)((h->K+warps-1)/warps);
#define NQS_FF_CASE(J) case J: set_smem(ffnn_sweep_fast_kernel<J>, smem); ffnn_sweep_fast_kernel<J><<<grid, warps*32, smem, h->stream>>>(f); break
      switch (h->jpl) { NQS_FF_CASE(1); NQS_FF_CASE(2); NQS_FF_CASE(4); NQS_FF_CASE(8); default: NQS_FF_CASE(16); }
#undef NQS_FF_CASE
      check_launch(h, "ffnn_sweep_fast_kernel");
      h->variant_sweep = "ffnn_resident_j"+std::to_string(h->jpl)+"_w"+std::to_string(warps);
      h->pos = (int)((h->pos+nsteps)%h->N);
      if (h->u_steps > 0) h->u_used += nsteps;
      h->step_counter += (unsigned long long)nsteps;
      return;
    }
  }
  // fp32-filtered exact sweep (sweep_f32.cuh): one chain per warp, fp64 state in shared memory.  OPT-IN (NQS_SWEEP_F32=1): exact
  // (tests/test_gpu_parity.py::test_f32_filtered_sweep_is_exact) but SLOWER than the all-fp64 register kernel -- 2.85 vs 1.94 ms
  // per sweep at N=128, M=256, K=16384, 0.43 vs 0.29 ms at K=2048: the accept path through shared memory and the fp64 decision
  // arithmetic cost what the fp32 product saves (~375 instructions per proposal and chain either way), so the dependency chain
  // per proposal did not get shorter.  Kept as the record of the experiment.  The fp32 product of up to 16 factors per lane
  // must stay far inside the fp32 range: 16 * 2 max|Re theta| * log2(e) < ~100.
  if (h->model == MODEL_RBM && fast_path_ok(h) && h->jpl <= 16 && nsteps%h->N == 0 && h->ftab32.p != nullptr &&
      h->theta_bound < 30.0/h->jpl && (std::getenv("NQS_SWEEP_F32") && std::atoi(std::getenv("NQS_SWEEP_F32")) != 0))
  {
    int warps = 8;
    while (warps > 1 && f32_sweep_smem_bytes(h->N, warps, h->mpad) > h->smem_optin) warps >>= 1;
    if (f32_sweep_smem_bytes(h->N, warps, h->mpad) <= h->smem_optin)
    {
      F32SweepArgs fa;
      FastSweepArgs & f = fa.b;
      f.N = h->N; f.M = h->M; f.Mpad = h->mpad; f.K = h->K; f.params = h->params.p; f.ftab_a = h->ftab_a.p; f.ftab_b = h->ftab_b.p; f.w2 = h->w2.p; f.afac = h->afac.p;
      f.spins = h->spins.p; f.theta = h->theta.p; f.lnpsi0 = h->lnpsi0.p; f.sa = h->sa.p; f.fresh = h->fresh.p; f.order = h->order.p;
      f.pos0 = h->pos; f.nsweeps = (int)(nsteps/h->N); f.uniforms = a.uniforms; f.seed = a.seed; f.step0 = a.step0;
      f.chain_offset = a.chain_offset; f.acc_log = a.acc_log;
      fa.ftab32 = h->ftab32.p; fa.stats = h->f32_stats.p;
      fa.delta_scale = 1.0f;
      { const char * e = std::getenv("NQS_SWEEP_F32_DELTA"); if (e) fa.delta_scale = (float)std::atof(e); }
      const size_t smem = f32_sweep_smem_bytes(h->N, warps, h->mpad);
      const unsigned grid = (unsigned)((h->K+warps-1)/warps);
#define NQS_S32_CASE(J) case J: set_smem(rbm_sweep_f32_kernel<J>, smem); rbm_sweep_f32_kernel<J><<<grid, warps*32, smem, h->stream>>>(fa); break
      switch (h->jpl) { NQS_S32_CASE(1); NQS_S32_CASE(2); NQS_S32_CASE(4); NQS_S32_CASE(8); default: NQS_S32_CASE(16); }
#undef NQS_S32_CASE
      check_launch(h, "rbm_sweep_f32_kernel");
      h->variant_sweep = "rbm_f32filter_j"+std::to_string(h->jpl)+"_w"+std::to_string(warps);
      h->pos = (int)((h->pos+nsteps)%h->N);
      if (h->u_steps > 0) h->u_used += nsteps;
      h->step_counter += (unsigned long long)nsteps;
      return;
    }
  }
  if (h->model == MODEL_RBM && fast_path_ok(h) && h->jpl <= 32 && nsteps%h->N == 0 && fast_sweep_fits(h))
  {
    FastSweepArgs f;
    f.N = h->N; f.M = h->M; f.Mpad = h->mpad; f.K = h->K; f.params = h->params.p; f.ftab_a = h->ftab_a.p; f.ftab_b = h->ftab_b.p; f.w2 = h->w2.p; f.afac = h->afac.p;
    f.spins = h->spins.p; f.theta = h->theta.p; f.lnpsi0 = h->lnpsi0.p; f.sa = h->sa.p; f.fresh = h->fresh.p; f.order = h->order.p;
    f.pos0 = h->pos; f.nsweeps = (int)(nsteps/h->N); f.uniforms = a.uniforms; f.seed = a.seed; f.step0 = a.step0;
    f.chain_offset = a.chain_offset; f.acc_log = a.acc_log;
    const int C = sweep_chains_per_warp(h);
#define NQS_SWEEP_CASE(J) case J: if (C == 4) launch_sweep_fast_t<J, 4>(h, f); else if (C == 2) launch_sweep_fast_t<J, 2>(h, f); else launch_sweep_fast_t<J, 1>(h, f); break
    switch (h->jpl)
    {
      NQS_SWEEP_CASE(1); NQS_SWEEP_CASE(2); NQS_SWEEP_CASE(4);
      case 8: if (C == 2) launch_sweep_fast_t<8, 2>(h, f); else launch_sweep_fast_t<8, 1>(h, f); break;
      case 16: launch_sweep_fast_t<16, 1>(h, f); break;
      default: launch_sweep_fast_t<16, 1, 2>(h, f); break;     // jpl = 32 (M <= 1024): two warps per chain
    }
#undef NQS_SWEEP_CASE
    check_launch(h, "rbm_sweep_fast_kernel");
    h->variant_sweep = "rbm_regs_j"+std::to_string(h->jpl)+"_c"+std::to_string(C);
    h->pos = (int)((h->pos+nsteps)%h->N);
    if (h->u_steps > 0) h->u_used += nsteps;
    h->step_counter += (unsigned long long)nsteps;
    return;
  }
  h->variant_sweep = "generic";
  const size_t npad = (size_t)((h->N+15)/16)*16;
  const size_t per_warp = (size_t)h->M*sizeof(cd)+npad, fixed = (size_t)h->N*sizeof(int);
  const int warps = warps_for_smem(h, per_warp, fixed);
  const size_t smem = per_warp*warps+fixed;
  const int grid = (int)((h->K+warps-1)/warps);
  if (h->model == MODEL_RBM)
  {
    set_smem(sweep_generic_kernel<MODEL_RBM>, smem);
    sweep_generic_kernel<MODEL_RBM><<<grid, warps*32, smem, h->stream>>>(a);
  }
  else
  {
    set_smem(sweep_generic_kernel<MODEL_FFNN>, smem);
    sweep_generic_kernel<MODEL_FFNN><<<grid, warps*32, smem, h->stream>>>(a);
  }
  check_launch(h, "sweep_generic_kernel");
  h->pos = (int)((h->pos+nsteps)%h->N);
  if (h->u_steps > 0) h->u_used += nsteps;
  h->step_counter += (unsigned long long)nsteps;
}

void launch_eloc(nqs_handle * h, cd * lnpsi1, int single_site)
{
  if (lnpsi1 == nullptr && h->model == MODEL_RBM && fast_path_ok(h))
  {
    launch_eloc_fast(h);
    h->variant_eloc = "rbm_sites_c4";
    return;
  }
  if (lnpsi1 == nullptr && h->model == MODEL_FFNN && fast_path_ok(h) && ffnn_eloc_smem_bytes(h->N, h->M, 4) <= h->smem_optin)
  {
    FfnnElocArgs f;
    f.N = h->N; f.M = h->M; f.Npad = h->npad32; f.K = h->K; f.params = h->params.p; f.ctabT_a = h->ctabT_a.p; f.ctabT_b = h->ctabT_b.p;
    f.spins = h->spins.p; f.theta = h->theta.p; f.lnpsi0 = h->lnpsi0.p; f.Jmat = h->Jmat.p; f.hfield = h->cfg.h; f.htilda = h->htilda.p;
    f.sjs = launch_sjs(h);
    const int nwarps = std::max(1, std::min(8, (h->N+31)/32));
    const size_t smem = ffnn_eloc_smem_bytes(h->N, h->M, 4);
    set_smem(ffnn_eloc_fast_kernel<4>, smem);
    ffnn_eloc_fast_kernel<4><<<(unsigned)((h->K+3)/4), nwarps*32, smem, h->stream>>>(f);
    check_launch(h, "ffnn_eloc_fast_kernel");
    h->variant_eloc = "ffnn_sites_c4";
    return;
  }
  if (lnpsi1 == nullptr) h->variant_eloc = "generic";
  ElocArgs a;
  a.N = h->N; a.M = h->M; a.model = h->model; a.K = h->K; a.params = h->params.p; a.spins = h->spins.p;
  a.theta = h->theta.p; a.lnpsi0 = h->lnpsi0.p; a.sa = h->sa.p; a.Jmat = h->Jmat.p; a.hfield = h->cfg.h;
  a.htilda = h->htilda.p; a.lnpsi1 = lnpsi1; a.single_site = single_site;
  a.sjs = (lnpsi1 == nullptr) ? launch_sjs(h) : nullptr;
  const size_t npad = (size_t)((h->N+15)/16)*16;
  const size_t per_warp = (size_t)h->M*sizeof(cd)+npad;
  const int warps = warps_for_smem(h, per_warp, 0);
  const size_t smem = per_warp*warps;
  const int grid = (int)((h->K+warps-1)/warps);
  if (h->model == MODEL_RBM)
  {
    set_smem(eloc_generic_kernel<MODEL_RBM>, smem);
    eloc_generic_kernel<MODEL_RBM><<<grid, warps*32, smem, h->stream>>>(a);
  }
  else
  {
    set_smem(eloc_generic_kernel<MODEL_FFNN>, smem);
    eloc_generic_kernel<MODEL_FFNN><<<grid, warps*32, smem, h->stream>>>(a);
  }
  check_launch(h, "eloc_generic_kernel");
}

void launch_oderiv(nqs_handle * h, cudaStream_t stream = nullptr)
{
  if (stream == nullptr) stream = h->stream;
  const size_t smem = (size_t)(h->model == MODEL_FFNN ? 2 : 1)*h->M*sizeof(cd)+(size_t)h->N*sizeof(double);
  NQS_REQUIRE(smem <= h->smem_optin, NQS_ERR_UNSUPPORTED, "n_hiddens too large for oderiv_kernel shared memory");
  if (h->tied == TIED_RBM_Z2PR)
  {
    set_smem(oderiv_z2pr_kernel, smem);
    oderiv_z2pr_kernel<<<(unsigned)h->K, 256, smem, stream>>>(h->N, h->alpha_f, h->K, h->spins.p, h->theta.p, h->O.p);
  }
  else if (h->tied == TIED_FFNN_TR)
  { // w1of = the third block of the expanded FFNN parameters
    set_smem(oderiv_ffnntr_kernel, smem);
    oderiv_ffnntr_kernel<<<(unsigned)h->K, 256, smem, stream>>>(h->N, h->alpha_f, h->K, h->params.p+(size_t)h->N*h->M+h->M,
      h->spins.p, h->theta.p, h->O.p);
  }
  else if (h->trsymm)
  {
    set_smem(oderiv_trsymm_kernel, smem);
    oderiv_trsymm_kernel<<<(unsigned)h->K, 256, smem, stream>>>(h->N, h->alpha_f, h->K, h->spins.p, h->theta.p, h->O.p);
  }
  else if (h->model == MODEL_RBM)
  {
    set_smem(oderiv_kernel<MODEL_RBM>, smem);
    oderiv_kernel<MODEL_RBM><<<(unsigned)h->K, 256, smem, stream>>>(h->N, h->M, h->K, h->params.p, h->spins.p, h->theta.p, h->O.p);
  }
  else
  {
    set_smem(oderiv_kernel<MODEL_FFNN>, smem);
    oderiv_kernel<MODEL_FFNN><<<(unsigned)h->K, 256, smem, stream>>>(h->N, h->M, h->K, h->params.p, h->spins.p, h->theta.p, h->O.p);
  }
  check_launch(h, "oderiv_kernel");
  h->o_pending = false;
  h->theta_matches_O = true;   // O was built from the current (spins, theta, params): the setup sums may use the factors
}

// ---- one-pass S*v (sv_fused.cuh): cluster launch + plan -------------------------------------------------------------------
template <int CPT, int DEFER, int GEN = 0>
cudaError_t sv_launch_t(const SvArgs & a, int cs, int nclusters, int nt, size_t smem, cudaStream_t stream, int * query_max_clusters)
{
  auto kern = sv_fused_kernel<CPT, DEFER, GEN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (cs > 8)
  {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(cs*nclusters), 1, 1);
  cfg.blockDim = dim3((unsigned)nt+32, 1, 1);   // nt consumer threads + the producer warp
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  if (query_max_clusters) return cudaOccupancyMaxActiveClusters(query_max_clusters, kern, &cfg);
  return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t sv_launch(int cpt, int defer, const SvArgs & a, int cs, int nclusters, int nt, size_t smem, cudaStream_t stream, int * q)
{
  switch (cpt)
  {
    case 1: return defer ? sv_launch_t<1, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<1, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 2: return defer ? sv_launch_t<2, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<2, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 3: return defer ? sv_launch_t<3, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<3, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 4: return defer ? sv_launch_t<4, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<4, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 5: return defer ? sv_launch_t<5, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<5, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 6: return defer ? sv_launch_t<6, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<6, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 7: return defer ? sv_launch_t<7, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<7, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 8: return defer ? sv_launch_t<8, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<8, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 9: return defer ? sv_launch_t<9, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<9, 0>(a, cs, nclusters, nt, smem, stream, q);
    case 10: return defer ? sv_launch_t<10, 1>(a, cs, nclusters, nt, smem, stream, q) : sv_launch_t<10, 0>(a, cs, nclusters, nt, smem, stream, q);
    default: return cudaErrorInvalidValue;
  }
}

// GEN variant (the rows of O are formed from their factors inside the kernel and written out): software-pipelined form only
cudaError_t sv_launch_gen(int cpt, const SvArgs & a, int cs, int nclusters, int nt, size_t smem, cudaStream_t stream)
{
  switch (cpt)
  {
    case 1: return sv_launch_t<1, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 2: return sv_launch_t<2, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 3: return sv_launch_t<3, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 4: return sv_launch_t<4, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 5: return sv_launch_t<5, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 6: return sv_launch_t<6, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 7: return sv_launch_t<7, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 8: return sv_launch_t<8, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 9: return sv_launch_t<9, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    case 10: return sv_launch_t<10, 1, 1>(a, cs, nclusters, nt, smem, stream, nullptr);
    default: return cudaErrorInvalidValue;
  }
}

// ---- persistent CG (cg_persist.cuh): same cluster geometry as sv_fused_kernel, cooperative so that every CTA is resident
template <int CPT, int DEFER>
cudaError_t cgp_launch_t(const CgpArgs & a, int cs, int nclusters, int nt, size_t smem, cudaStream_t stream, bool coop, int * query_max_clusters)
{
  auto kern = cg_persist_kernel<CPT, DEFER>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (cs > 8)
  {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(cs*nclusters), 1, 1);
  cfg.blockDim = dim3((unsigned)nt+32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative; attr[1].val.cooperative = 1;
  cfg.attrs = attr; cfg.numAttrs = coop ? 2 : 1;
  if (query_max_clusters) { cfg.numAttrs = 1; return cudaOccupancyMaxActiveClusters(query_max_clusters, kern, &cfg); }
  return cudaLaunchKernelEx(&cfg, kern, a);
}

cudaError_t cgp_launch(int cpt, int defer, const CgpArgs & a, int cs, int nclusters, int nt, size_t smem, cudaStream_t stream, bool coop, int * q)
{
  switch (cpt)
  {
#define NQS_CGP_CASE(C) case C: return defer ? cgp_launch_t<C, 1>(a, cs, nclusters, nt, smem, stream, coop, q) : cgp_launch_t<C, 0>(a, cs, nclusters, nt, smem, stream, coop, q)
    NQS_CGP_CASE(1); NQS_CGP_CASE(2); NQS_CGP_CASE(3); NQS_CGP_CASE(4); NQS_CGP_CASE(5);
    NQS_CGP_CASE(6); NQS_CGP_CASE(7); NQS_CGP_CASE(8); NQS_CGP_CASE(9); NQS_CGP_CASE(10);
#undef NQS_CGP_CASE
    default: return cudaErrorInvalidValue;
  }
}

// The persistent solve can be used when the one-pass plan exists, its grid fits the per-CTA reduction slots, every consumer
// thread owns at most EPT vector elements and all clusters are resident at once.
void plan_cgp(nqs_handle * h)
{
  h->cgp_ok = false;
  if (!h->sv_ok || h->gen_ok || h->struct_sv) return;
  // OPT-IN (NQS_CG_PERSIST=1).  Measured in round 2 (profiles/r2_cg_persistent.md): the persistent kernel's pass over O carries
  // ~8 % more instructions per row than sv_fused_kernel's (ring bookkeeping across products, register moves) and that loop is
  // issue-sensitive, so it streams 3-7 % slower; its vector phase (3 grid barriers + folds, ~20 us; ~37 us with the 8-GPU
  // exchange and rank skew) is no shorter than one cg_fused_kernel launch (~27 / ~42 us).  Net: 2 % slower on one GPU, equal on
  // eight, once the launch-per-iteration path stopped polling (cg_solve_async_begin).  Kept for the record and for A/B runs.
  { const char * e = std::getenv("NQS_CG_PERSIST"); if (!(e && std::atoi(e) != 0)) return; }
  const long long ctas = (long long)h->sv_nclusters*h->sv_cs;
  if (ctas > NQS_CGP_MAX_CTAS) return;
  const long long threads = ctas*h->sv_nt;
  const int ept = (h->sv_cpt <= 3) ? 1 : 2;
  if ((h->P+threads-1)/threads > ept) return;
  CgpArgs a;
  std::memset(&a, 0, sizeof(a));
  int maxc = 0;
  const cudaError_t e = cgp_launch(h->sv_cpt, h->sv_defer, a, h->sv_cs, 1, h->sv_nt, h->sv_smem, h->stream, false, &maxc);
  if (e != cudaSuccess || maxc < h->sv_nclusters) { cudaGetLastError(); return; }
  h->cgp_ok = true;
  // NQS_CG_COOP=0: plain cluster launch (profilers cannot replay a cooperative cluster launch); the barrier time-out then guards residency
  { const char * c = std::getenv("NQS_CG_COOP"); if (c && std::atoi(c) == 0) h->cgp_coop = 0; }
}

// Pick cluster size / columns per thread / pipeline depth for this (K, P); leaves sv_ok false when the column slice of a
// 16-CTA cluster still does not fit the register file (very large P: the two-pass kernels take over).
void plan_sv(nqs_handle * h)
{
  h->sv_ok = false;
  if (h->cfg.flags & NQS_FLAG_TWO_PASS_SV) return;
  // software-pipelined variant by default (1.50 ms vs 1.95 ms per S*v at N=128, M=256, K=16384; needs >= 3 row slots)
  int want_defer = 1;
  { const char * d = std::getenv("NQS_SV_DEFER"); if (d) want_defer = std::atoi(d); }
  // Candidate cluster sizes.  What matters is how many SMs the resident clusters cover: clusters must sit inside one GPC, so
  // on a B200 15 clusters of 8 cover 120 SMs, 15 clusters of 9 cover 135 (1.39 vs 1.47 ms per S*v at N=128, M=256, K=16384),
  // 11 clusters of 10 cover 110 and 7 clusters of 16 cover 112 (measured, profiles/r1d_sv_fused_experiments.md).  Sizes above 8
  // are "non-portable" and need cudaFuncAttributeNonPortableClusterSizeAllowed.  NQS_SV_CS pins the size (tests).
  // Small clusters (1 .. 7) serve NARROW rows: the reduction + DSMEM exchange costs every row ~0.75 us whatever the slice width, so
  // a slice must carry >= ~0.75 us of HBM time per row or the kernel is latency-bound (N=64, M=128: 16 CTAs x 8 KB slices ran at
  // 0.33 of the HBM peak; 2 CTAs x 67 KB slices are bandwidth-bound again).  The plan with the smallest ESTIMATED time wins.
  std::vector<int> cs_list = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16};
  { const char * c = std::getenv("NQS_SV_CS"); if (c && std::atoi(c) >= 1 && std::atoi(c) <= NQS_SV_MAX_CLUSTER) cs_list = {std::atoi(c)}; }
  double best_time = -1.0;
  for (const int cs : cs_list)
  {
    const long long pc = (h->P+cs-1)/cs;
    // columns per thread / consumer threads: every row costs each WARP a fixed ~150 instructions (reduction, exchange, waits),
    // so few fat warps beat many thin ones: prefer 8 columns per thread down to 128 threads, then 9-10 columns (register
    // limit), then whatever fits (tiny slices).  One more warp is the TMA producer.
    int cpt = 0, nt = 0;
    static const int fat_order[] = {8, 7, 6, 5, 4, 3, 2, 1, 9, 10}, thin_order[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10};
    const char * fat_env = std::getenv("NQS_SV_FAT");
    const bool fat = fat_env ? std::atoi(fat_env) != 0 : true;
    for (int pass = (fat ? 0 : 1); pass < 2 && !cpt; ++pass)
      for (int q = 0; q < NQS_SV_MAX_CPT && !cpt; ++q)
      {
        const int c = (pass == 0) ? fat_order[q] : thin_order[q];
        const long long need = ((pc+c-1)/c+31)/32*32;
        const int max_t = (c <= 3) ? 992 : 480;
        if (need <= max_t && (pass == 1 || need >= 128)) { cpt = c; nt = (int)std::max<long long>(64, need); }
      }
    if (!cpt) continue;
    // a slot holds one row slice, padded to CPT * consumer threads elements so the kernel reads it without bounds checks
    const size_t slot_bytes = (size_t)((std::max<long long>(pc, (long long)cpt*nt)*(long long)sizeof(cd)+127)/128*128);
    if (h->smem_optin < NQS_SV_TAIL_BYTES+2*slot_bytes) continue;
    const int nslot = (int)std::min<size_t>(NQS_SV_MAX_SLOTS, (h->smem_optin-NQS_SV_TAIL_BYTES)/slot_bytes);
    const size_t smem = (size_t)nslot*slot_bytes+NQS_SV_TAIL_BYTES;
    const int defer = (nslot >= 3) ? want_defer : 0;
    SvArgs a;
    std::memset(&a, 0, sizeof(a));
    int maxc = 0;
    cudaError_t e = sv_launch(cpt, defer, a, cs, 1, nt, smem, h->stream, &maxc);
    if (e != cudaSuccess || maxc < 1) { cudaGetLastError(); continue; }
    long long ncl = std::min<long long>(maxc, h->K);
    const long long rpc = (h->K+ncl-1)/ncl;
    ncl = (h->K+rpc-1)/rpc;
    // estimated time of one S*v: the HBM stream over the SMs the clusters cover (several narrow CTAs per SM cover it once),
    // against the per-row latency floor of the exchange (software-pipelined: ~0.75 us per row; in order: ~1.2 us)
    const double cover = std::min(1.0, (double)(ncl*cs)/(double)h->sm_count);
    const double t_hbm = (double)h->K*(double)h->P*16.0/(6.4e12*cover);
    const double t_lat = (double)rpc*(defer ? 0.75e-6 : 1.2e-6)*(cs == 1 ? 0.6 : 1.0);
    const double t_est = std::max(t_hbm, t_lat)*(1.0+0.002*cs);       // ties: the smaller cluster (fewer exchange partners)
    if (best_time >= 0.0 && t_est >= best_time) continue;
    best_time = t_est;
    h->sv_defer = defer;
    h->sv_depth = defer ? std::max(1, std::min(NQS_SV_MAX_DEPTH, (nslot-1)/2)) : 0;   // keep >= depth+1 slots for rows in flight
    { const char * d = std::getenv("NQS_SV_DEPTH"); if (d && defer) h->sv_depth = std::max(1, std::min(std::min(NQS_SV_MAX_DEPTH, nslot-2), std::atoi(d))); }
    h->sv_cs = cs; h->sv_cpt = cpt; h->sv_nt = nt; h->sv_nslot = nslot; h->sv_nclusters = (int)ncl;
    h->sv_smem = smem; h->sv_slot_bytes = slot_bytes; h->sv_pc = pc; h->sv_rpc = rpc;
    h->sv_ok = true;
    h->variant_sv = "fused_cs"+std::to_string(cs)+"_cpt"+std::to_string(cpt)+"_nt"+std::to_string(nt)+"_slots"+std::to_string(nslot)+
      "_clusters"+std::to_string(ncl)+(h->sv_defer ? "_defer"+std::to_string(h->sv_depth) : "");
  }
}

int cg_ctas(const nqs_handle * h);

void allreduce_sum(nqs_handle * h, double * buf, size_t count)
{
  if (h->comm == nullptr) return;
  const int rc = g_nccl.allReduce(buf, buf, count, kNcclFloat64, kNcclSum, h->comm, h->stream);
  if (rc != 0)
    throw Error(NQS_ERR_NCCL, std::string("ncclAllReduce failed: ")+(g_nccl.getErrorString ? g_nccl.getErrorString(rc) : "?"));
}

template <int MODEL>
void launch_setup_structured_m(nqs_handle * h, dim3 grid, size_t smem, int ipt)
{
#define NQS_SS_CASE(I) setup_structured_kernel<MODEL, I><<<grid, NQS_SS_THREADS, smem, h->stream>>>(h->N, h->M, h->K, h->params.p, \
    h->spins.p, h->theta.p, h->htilda.p, h->part.p, h->rows_per_block)
  if (ipt <= 1) NQS_SS_CASE(1);
  else if (ipt <= 2) NQS_SS_CASE(2);
  else if (ipt <= 4) NQS_SS_CASE(4);
  else if (ipt <= 8) NQS_SS_CASE(8);
  else NQS_SS_CASE(16);
#undef NQS_SS_CASE
}

void launch_setup_structured(nqs_handle * h, dim3 grid, size_t smem, int ipt)
{
  if (h->model == MODEL_RBM) launch_setup_structured_m<MODEL_RBM>(h, grid, smem, ipt);
  else launch_setup_structured_m<MODEL_FFNN>(h, grid, smem, ipt);
}

// ---- structured S*v (sv_struct.cuh) ---------------------------------------------------------------------------------------
void launch_hidden_values(nqs_handle * h)
{
  const int grid = grid_for((long long)h->K*h->M, 256, 148*8);
  if (h->model == MODEL_RBM)
    hidden_values_kernel<MODEL_RBM><<<grid, 256, 0, h->stream>>>(h->N, h->M, h->K, h->params.p, h->theta.p, h->Tm.p, nullptr);
  else
    hidden_values_kernel<MODEL_FFNN><<<grid, 256, 0, h->stream>>>(h->N, h->M, h->K, h->params.p, h->theta.p, h->Tm.p, h->Lm.p);
  check_launch(h, "hidden_values_kernel");
  if (h->cols_umma)
  { // per-hidden-unit bound of |T| for the int8 split of conj(T) z (cols_umma.cuh); T is fixed until the next sweep
    NQS_CUDA(cudaMemsetAsync(h->tmaxb.p, 0, sizeof(unsigned long long)*(size_t)h->M, h->stream));
    const long long rpb = std::max<long long>(64, (h->K+63)/64);
    dim3 cg((unsigned)((h->M+31)/32), (unsigned)((h->K+rpb-1)/rpb));
    colmax_abs_kernel<<<cg, 256, 0, h->stream>>>(h->K, h->M, h->Tm.p, h->tmaxb.p, rpb);
    check_launch(h, "colmax_abs_kernel");
  }
  h->hidden_valid = true;
  h->theta_matches_O = true;   // the factors (spins, theta, params) are what S is built from: structured setup sums allowed
}

// warp grids / tile shapes of spin_cols_dmma_kernel by chain length: {MTW, NTW, WM}
static const int kColsVariants[5][3] = {{1, 8, 2}, {1, 16, 4}, {1, 16, 8}, {2, 16, 8}, {4, 8, 8}};
int cols_variant_for(int N) { return N <= 16 ? 0 : N <= 32 ? 1 : N <= 64 ? 2 : N <= 128 ? 3 : 4; }

// geometry of spin_cols_dmma_kernel (structured S*v and the SR setup GEMM): column groups x chain chunks ~ one CTA per SM
void plan_cols(nqs_handle * h)
{
  h->cols_ok = false;
  if (h->N > 256 || (h->cfg.flags & NQS_FLAG_NO_DMMA) || h->trsymm) return;    // (tied weights: the rows of O are not outer products)
  { // tcgen05 int8 kernel (cols_umma.cuh): 128 sites per MMA, one CTA per SM; pays once a rank holds enough chains to give
    // every CTA several 64-chain blocks.  NQS_COLS_UMMA=0/1 overrides.
    // (FNN: measured slower than the DMMA kernel at cfg4, 0.277 vs 0.214 ms per product -- the FFNN instantiation spills -- so
    // it is taken for the RBM only unless asked for.)
    const char * e = std::getenv("NQS_COLS_UMMA");
    bool on = (h->K >= 8192 && h->model == MODEL_RBM);
    if (e) on = (std::atoi(e) != 0);
    on = on && h->N <= 128 && cols_umma_smem() <= h->smem_optin;
    if (on)
    {
      const int colgroups = (2*h->M+NQS_CU_NCC-1)/NQS_CU_NCC;
      long long nchunks = std::max(1, h->sm_count/colgroups);
      nchunks = std::min<long long>(nchunks, (h->K+NQS_CU_KB-1)/NQS_CU_KB);
      long long rpc = (h->K+nchunks-1)/nchunks;
      rpc = (rpc+NQS_CU_KB-1)/NQS_CU_KB*NQS_CU_KB;
      h->sc_variant = -1; h->sc_colgroups = colgroups; h->sc_rows_per_chunk = rpc; h->sc_nchunks = (int)((h->K+rpc-1)/rpc);
      h->cols_umma = 1;
      h->tmaxb.alloc((size_t)h->M);
      h->cols_ok = true;
      return;
    }
  }
  const int var = cols_variant_for(h->N);
  const int cw = cols_dmma_cw(kColsVariants[var][1], kColsVariants[var][2]);
  const int colgroups = (2*h->M+cw-1)/cw;
  long long nchunks = std::max(1, h->sm_count/colgroups);
  nchunks = std::min<long long>(nchunks, (h->K+NQS_DC_KC-1)/NQS_DC_KC);
  long long rpc = (h->K+nchunks-1)/nchunks;
  rpc = (rpc+NQS_DC_KC-1)/NQS_DC_KC*NQS_DC_KC;
  h->sc_variant = var; h->sc_colgroups = colgroups; h->sc_rows_per_chunk = rpc; h->sc_nchunks = (int)((h->K+rpc-1)/rpc);
  h->cols_ok = true;
}

void plan_struct(nqs_handle * h)
{
  NQS_REQUIRE(h->N <= 256, NQS_ERR_UNSUPPORTED, "NQS_FLAG_STRUCTURED_SV supports n_inputs <= 256");
  NQS_REQUIRE(!h->trsymm, NQS_ERR_UNSUPPORTED, "NQS_FLAG_STRUCTURED_SV is not available for the translation-symmetric RBM");
  NQS_REQUIRE(!(h->cfg.flags & (NQS_FLAG_SETUP_FROM_O | NQS_FLAG_TWO_PASS_SV | NQS_FLAG_NO_DMMA)), NQS_ERR_INVALID,
    "NQS_FLAG_STRUCTURED_SV excludes NQS_FLAG_SETUP_FROM_O / NQS_FLAG_TWO_PASS_SV / NQS_FLAG_NO_DMMA");
  const int var = h->sc_variant;
  h->struct_sv = true;
  if (h->cols_umma)
  {
    h->variant_sv = "structured_umma_i8_ozaki7_colgroups"+std::to_string(h->sc_colgroups)+"_chunks"+std::to_string(h->sc_nchunks);
    return;
  }
  h->variant_sv = "structured_dmma_mtw"+std::to_string(kColsVariants[var][0])+"_ntw"+std::to_string(kColsVariants[var][1])+
    "_wm"+std::to_string(kColsVariants[var][2])+"_colgroups"+std::to_string(h->sc_colgroups)+"_chunks"+std::to_string(h->sc_nchunks);
}

template <int MODEL, int MTW, int NTW, int WM>
void launch_cols_dmma_t(nqs_handle * h, const ColsArgs & a)
{
  const size_t smem = cols_dmma_smem(cols_dmma_nsc(MTW, WM), cols_dmma_cw(NTW, WM));
  set_smem(spin_cols_dmma_kernel<MODEL, MTW, NTW, WM>, smem);
  dim3 grid((unsigned)h->sc_colgroups, (unsigned)h->sc_nchunks, a.zmode == 1 ? 2u : 1u);
  spin_cols_dmma_kernel<MODEL, MTW, NTW, WM><<<grid, NQS_DC_THREADS, smem, h->stream>>>(a);
}
template <int MODEL>
void launch_cols_dmma(nqs_handle * h, const ColsArgs & a)
{
  if (h->cols_umma)
  {
    ColsUmmaArgs ua;
    ua.c = a; ua.tmax = h->tmaxb.p;
    const size_t smem = cols_umma_smem();
    set_smem(spin_cols_umma_kernel<MODEL>, smem);
    dim3 grid((unsigned)h->sc_colgroups, (unsigned)h->sc_nchunks, a.zmode == 1 ? 2u : 1u);
    spin_cols_umma_kernel<MODEL><<<grid, NQS_CU_THREADS, smem, h->stream>>>(ua);
    check_launch(h, "spin_cols_umma_kernel");
  }
  else
  switch (h->sc_variant)
  {
    case 0: launch_cols_dmma_t<MODEL, 1, 8, 2>(h, a); break;
    case 1: launch_cols_dmma_t<MODEL, 1, 16, 4>(h, a); break;
    case 2: launch_cols_dmma_t<MODEL, 1, 16, 8>(h, a); break;
    case 3: launch_cols_dmma_t<MODEL, 2, 16, 8>(h, a); break;
    default: launch_cols_dmma_t<MODEL, 4, 8, 8>(h, a); break;
  }
  check_launch(h, "spin_cols_dmma_kernel");
  if (MODEL == MODEL_FFNN)
  {
    dim3 grid((unsigned)((h->M+NQS_LB_THREADS-1)/NQS_LB_THREADS), (unsigned)h->sc_nchunks, a.zmode == 1 ? 2u : 1u);
    ffnn_lblock_kernel<<<grid, NQS_LB_THREADS, 0, h->stream>>>(a);
    check_launch(h, "ffnn_lblock_kernel");
  }
}

// local sums -> all ranks' sums -> <O>, F, diag S.  Multi-GPU with mapped peers: one kernel that exchanges over NVLink and
// finalises (setup_exchange_finalize_kernel); otherwise ncclAllReduce (if sharded) + setup_finalize_kernel.
void finish_setup(nqs_handle * h, bool want_F)
{
  const long long P = h->P;
  if (h->comm != nullptr && h->p2p_ok && h->xbuf_setup_off != 0)
  {
    SetupXArgs a;
    std::memset(&a, 0, sizeof(a));
    a.P = P; a.inv_ktot = 1.0/(double)h->Ktot; a.sums = h->sums.p; a.hsall = h->hsall.p; a.aO = h->aO.p; a.F = want_F ? h->F.p : nullptr;
    a.diag = h->diag.p; a.n_ranks = h->n_ranks; a.rank = h->rank; a.epoch = ++h->setup_epoch; a.timeout_flag = &h->scal.p->peer_timeout;
    for (int r = 0; r < h->n_ranks; ++r)
    {
      a.peer_x[r] = reinterpret_cast<double*>((char*)h->peer_base[r]+h->xbuf_setup_off);
      a.peer_flag[r] = reinterpret_cast<unsigned int*>((char*)h->peer_base[r]+h->xbuf_setup_flag_off);
    }
    setup_exchange_finalize_kernel<<<cg_ctas(h), 256, 0, h->stream>>>(a);
    check_launch(h, "setup_exchange_finalize_kernel");
    return;
  }
  allreduce_sum(h, h->sums.p, (size_t)(5*P+3));
  setup_finalize_kernel<<<grid_for(P, 256, 148*4), 256, 0, h->stream>>>(P, 1.0/(double)h->Ktot, h->sums.p, h->aO.p,
    want_F ? h->F.p : nullptr, h->diag.p, h->hsall.p);
  check_launch(h, "setup_finalize_kernel");
}

// <O>, F, diag from ONE pass over O (+ one all-reduce of 5P+3 doubles across ranks)
void sr_setup(nqs_handle * h, bool want_F)
{
  const long long P = h->P, K = h->K;
  htilda_sums_kernel<<<1, 1024, 0, h->stream>>>(K, h->htilda.p, h->sums.p+5*P);
  check_launch(h, "htilda_sums_kernel");
  const int ipt = (h->N+15)/16;
  if (h->cols_ok && h->theta_matches_O && !(h->cfg.flags & NQS_FLAG_SETUP_FROM_O))
  { // sums from the factors of O as ONE launch of the tensor-core column GEMM: S^T conj(T) and S^T (conj(T) h)
    if (!h->hidden_valid) launch_hidden_values(h);
    ColsArgs c;
    c.N = h->N; c.M = h->M; c.K = K; c.P = P; c.spins = h->spins.p; c.T = h->Tm.p; c.L = h->Lm.p; c.zk = h->htilda.p;
    c.part = h->part.p; c.rows_per_chunk = h->sc_rows_per_chunk; c.done = nullptr;
    c.zmode = 1; c.part_stride = (long long)h->sc_nchunks*2*P; c.abs2 = h->abs2.p;
    if (h->model == MODEL_RBM) launch_cols_dmma<MODEL_RBM>(h, c); else launch_cols_dmma<MODEL_FFNN>(h, c);
    const int g = grid_for(P, 256, 148*4);
    if (h->model == MODEL_RBM)
      setup_fold_struct_kernel<MODEL_RBM><<<g, 256, 0, h->stream>>>(h->N, h->M, P, K, h->sc_nchunks, h->part.p, c.part_stride, h->abs2.p, h->sums.p);
    else
      setup_fold_struct_kernel<MODEL_FFNN><<<g, 256, 0, h->stream>>>(h->N, h->M, P, K, h->sc_nchunks, h->part.p, c.part_stride, h->abs2.p, h->sums.p);
    check_launch(h, "setup_fold_struct_kernel");
    finish_setup(h, want_F);
    return;
  }
  if (ipt <= 16 && h->theta_matches_O && !h->trsymm && !(h->cfg.flags & NQS_FLAG_SETUP_FROM_O))
  { // sums from the factors of O (spins, tanh theta): no pass over O
    dim3 grid((unsigned)((h->M+NQS_SS_JT-1)/NQS_SS_JT), (unsigned)h->nrb);
    int ipt_t = 1;
    while (ipt_t < ipt) ipt_t <<= 1;                      // template instantiations: 1, 2, 4, 8, 16 sites per thread
    const size_t smem = (size_t)NQS_SS_CH*16*ipt_t*sizeof(double);
    launch_setup_structured(h, grid, smem, ipt);
    check_launch(h, "setup_structured_kernel");
  }
  else
  {
    dim3 grid((unsigned)((P+NQS_COL_THREADS-1)/NQS_COL_THREADS), (unsigned)h->nrb);
    setup_partial_kernel<<<grid, NQS_COL_THREADS, 0, h->stream>>>(K, P, h->O.p, h->htilda.p, h->part.p, h->rows_per_block);
    check_launch(h, "setup_partial_kernel");
  }
  colsum_reduce_kernel<<<grid_for(5*P, 256, 148*8), 256, 0, h->stream>>>(P, 5, h->nrb, h->part.p, h->sums.p, nullptr);
  check_launch(h, "colsum_reduce_kernel");
  finish_setup(h, want_F);
}

// z = O v and the chunk partials of O^H z from the factors: two tensor-core GEMMs, no pass over O
int matvec_structured(nqs_handle * h, const cd * v, const int * done)
{
  NQS_REQUIRE(h->hidden_valid, NQS_ERR_STATE, "structured S*v before the hidden-unit factors were computed");
  const long long NM = (long long)h->N*h->M;
  RowsArgs r;
  std::memset(&r, 0, sizeof(r));
  r.N = h->N; r.M = h->M; r.K = h->K; r.spins = h->spins.p; r.T = h->Tm.p; r.L = h->Lm.p; r.zk = h->zk.p; r.done = done;
  {
    Span sp(h, TAG_ROWS);
    if (h->model == MODEL_RBM)
    { // v = [V (i*M+j) | a block | b block]
      r.B = reinterpret_cast<const double*>(v); r.avis = v+NM; r.bias = v+NM+h->N;
      if (rows_umma_ok(h)) launch_rows_umma<MODEL_RBM, ROWS_EPI_Z>(h, r); else launch_rows_dmma<MODEL_RBM, ROWS_EPI_Z>(h, r);
    }
    else
    { // v = [V (j*N+i) | b1 block | w1o block]
      transpose_wblock_kernel<<<grid_for(NM, 256, 148*4), 256, 0, h->stream>>>(h->N, h->M, v, h->vnat.p, done);
      check_launch(h, "transpose_wblock_kernel");
      r.B = reinterpret_cast<const double*>(h->vnat.p); r.bias = v+NM; r.w1o = v+NM+h->M;
      if (rows_umma_ok(h)) launch_rows_umma<MODEL_FFNN, ROWS_EPI_Z>(h, r); else launch_rows_dmma<MODEL_FFNN, ROWS_EPI_Z>(h, r);
    }
  }
  ColsArgs c;
  c.N = h->N; c.M = h->M; c.K = h->K; c.P = h->P; c.spins = h->spins.p; c.T = h->Tm.p; c.L = h->Lm.p; c.zk = h->zk.p;
  c.part = h->part.p; c.rows_per_chunk = h->sc_rows_per_chunk; c.done = done; c.zmode = 0; c.part_stride = 0; c.abs2 = nullptr;
  {
    Span sp(h, TAG_COLS);
    if (h->model == MODEL_RBM) launch_cols_dmma<MODEL_RBM>(h, c); else launch_cols_dmma<MODEL_FFNN>(h, c);
  }
  return h->sc_nchunks;
}

// The pass(es) over O of one S*v: cluster / row-block partials of sum_k conj(O_kp) (O_k . v) land in h->part.  Multi-GPU: they
// are folded into traw and all-reduced (2P doubles) here; single GPU: cg_fused_kernel folds them itself.  Returns the number
// of partials cg_fused_kernel has to fold (0 = read traw).
int matvec_passes(nqs_handle * h, const cd * v, const int * done)
{
  const long long P = h->P, K = h->K;
  int nparts;
  if (h->struct_sv) nparts = matvec_structured(h, v, done);
  else if (h->sv_ok)
  {
    SvArgs a;
    a.K = K; a.P = P; a.O = h->O.p; a.v = v; a.part = h->part.p; a.done = done; a.pc = h->sv_pc; a.rows_per_cluster = h->sv_rpc;
    a.nslot = h->sv_nslot; a.slot_bytes = (unsigned int)h->sv_slot_bytes; a.depth = h->sv_depth;
    a.T = h->Tm.p; a.spins8 = h->spins.p; a.Ow = h->O.p; a.N = h->N; a.M = h->M; a.gen_q = 0;
    Span sp(h, TAG_ROWS);
    if (h->o_pending)
    { // first product after the sampling phase: this launch also WRITES O (no separate O writer ran)
      NQS_REQUIRE(done == nullptr, NQS_ERR_STATE, "the O-generating S*v cannot be skipped");
      a.nslot = NQS_SV_MAX_SLOTS; a.slot_bytes = (unsigned int)sv_gen_slot_bytes(h->N, h->M);
      a.depth = std::min(NQS_SV_MAX_DEPTH, (a.nslot-1)/2);
      const size_t smem = (size_t)a.nslot*a.slot_bytes+NQS_SV_TAIL_BYTES;
      // its own geometry: a consumer thread count that is a multiple of M makes every element one sign flip of one T value
      a.pc = h->gen_pc; a.rows_per_cluster = h->gen_rpc; a.gen_q = h->gen_q;
      NQS_CUDA(sv_launch_gen(h->gen_cpt, a, h->gen_cs, h->gen_nclusters, h->gen_nt, smem, h->stream));
      h->o_pending = false;
      nparts = h->gen_nclusters;
    }
    else
    {
      NQS_CUDA(sv_launch(h->sv_cpt, h->sv_defer, a, h->sv_cs, h->sv_nclusters, h->sv_nt, h->sv_smem, h->stream, nullptr));
      nparts = h->sv_nclusters;
    }
    check_launch(h, "sv_fused_kernel");
  }
  else
  {
    const unsigned gr = (unsigned)((K+NQS_ROWS_PER_CTA-1)/NQS_ROWS_PER_CTA);
    {
      Span sp(h, TAG_ROWS);
      matvec_rows_kernel<<<gr, NQS_ROW_THREADS, 0, h->stream>>>(K, P, h->O.p, v, h->zk.p, done);
      check_launch(h, "matvec_rows_kernel");
    }
    dim3 grid((unsigned)((P+NQS_COL_THREADS-1)/NQS_COL_THREADS), (unsigned)h->nrb);
    {
      Span sp(h, TAG_COLS);
      matvec_cols_partial_kernel<<<grid, NQS_COL_THREADS, 0, h->stream>>>(K, P, h->O.p, h->zk.p, h->part.p, h->rows_per_block, done);
      check_launch(h, "matvec_cols_partial_kernel");
    }
    nparts = h->nrb;
  }
  if (h->comm == nullptr || h->p2p_ok) return nparts;     // single GPU, or the all-reduce happens inside cg_fused_kernel
  colsum_reduce_kernel<<<grid_for(2*P, 256, 148*8), 256, 0, h->stream>>>(P, 2, nparts, h->part.p, h->traw.p, done);
  check_launch(h, "colsum_reduce_kernel");
  allreduce_sum(h, h->traw.p, (size_t)(2*P));
  return 0;
}

#define NQS_CG_TRACE_MAX 8192
// NQS_CG_TRACE=1: dump the time stamps of the last launches to stderr (microseconds relative to the launch's entry stamp)
void dump_cg_trace(nqs_handle * h)
{
  if (h->cg_trace.p != nullptr && h->cgp_ok)
  { // persistent solve: stamps of CTA 0 per product of the LAST solve (us relative to the product's start)
    std::vector<unsigned long long> t((size_t)64*NQS_CGP_TRACE_WORDS);
    if (cudaMemcpy(t.data(), h->cg_trace.p, t.size()*sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
    for (int q = 0; q < 64; ++q)
    {
      const unsigned long long * r = t.data()+(size_t)q*NQS_CGP_TRACE_WORDS;
      if (r[0] == 0 || r[1] == 0) break;
      auto us = [&](int i) { return r[i] ? ((double)r[i]-(double)r[0])*1e-3 : -1.0; };
      std::fprintf(stderr, "cgptrace rank %d product %d: since_prev_end %.1f rows_done %.1f barrierA %.1f flags_raised %.1f peers_seen %.1f sum1 %.1f sum2 %.1f end %.1f\n",
        h->rank, q, q > 0 ? ((double)r[0]-(double)t[(size_t)(q-1)*NQS_CGP_TRACE_WORDS+7])*1e-3 : 0.0, us(1), us(2), us(3), us(4), us(5), us(6), us(7));
    }
    return;
  }
  if (h->cg_trace.p == nullptr || h->cg_trace_n == 0) return;
  std::vector<unsigned long long> t((size_t)h->cg_trace_n*NQS_CG_TRACE_WORDS);
  if (cudaMemcpy(t.data(), h->cg_trace.p, t.size()*sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess) return;
  const long long first = std::max<long long>(0, h->cg_trace_n-60);
  for (long long q = first; q < h->cg_trace_n; ++q)
  {
    const unsigned long long * r = t.data()+(size_t)q*NQS_CG_TRACE_WORDS;
    std::fprintf(stderr, "cgtrace rank %d launch %lld gap_prev_end %.1f stored %.1f released %.1f arrivals", h->rank, q,
      q > 0 ? ((double)r[0]-(double)t[(size_t)(q-1)*NQS_CG_TRACE_WORDS+20])*1e-3 : 0.0, ((double)r[1]-(double)r[0])*1e-3, ((double)r[2]-(double)r[0])*1e-3);
    for (int k = 0; k < h->n_ranks; ++k) std::fprintf(stderr, " %.1f", r[3+k] ? ((double)r[3+k]-(double)r[0])*1e-3 : -1.0);
    std::fprintf(stderr, " waited %.1f end %.1f\n", r[19] ? ((double)r[19]-(double)r[0])*1e-3 : -1.0, ((double)r[20]-(double)r[0])*1e-3);
  }
}

int cg_ctas(const nqs_handle * h)
{
  return std::max(1, std::min<int>(std::min(NQS_CG_MAX_CTAS, h->sm_count), (int)((h->P+NQS_CG_THREADS-1)/NQS_CG_THREADS)));
}

void launch_cg_fused(nqs_handle * h, int mode, int nparts, double lambda, cd * v)
{
  CgArgs a;
  a.P = h->P; a.mode = mode; a.nparts = nparts; a.part = h->part.p; a.traw = h->traw.p;
  a.inv_ktot = 1.0/(double)h->Ktot; a.lambda = lambda; a.aO = h->aO.p; a.diag = h->diag.p; a.F = h->F.p;
  a.v = v; a.pvec = h->pvec.p; a.x = h->dx.p; a.r = h->r.p; a.t = h->t.p; a.sc = h->scal.p; a.slots = h->slots.p; a.barrier = h->cgbar.p;
  a.n_ranks = 1; a.rank = 0; a.epoch = 0;
  a.hsums = (mode == CG_MODE_INIT && h->cg_check_finite) ? h->hsall.p : nullptr;
  for (int r = 0; r < NQS_CG_MAX_RANKS; ++r) { a.peer_x[r] = nullptr; a.peer_flag[r] = nullptr; }
  if (h->p2p_ok && nparts > 0)
  {
    a.n_ranks = h->n_ranks; a.rank = h->rank; a.epoch = ++h->p2p_epoch;
    for (int r = 0; r < h->n_ranks; ++r)
    {
      a.peer_x[r] = reinterpret_cast<double*>(h->peer_base[r]);
      a.peer_flag[r] = reinterpret_cast<unsigned int*>((char*)h->peer_base[r]+h->xbuf_data_bytes);
    }
  }
  a.trace = nullptr;
  if (h->cg_trace.p != nullptr && h->cg_trace_n < NQS_CG_TRACE_MAX)
    a.trace = h->cg_trace.p+(size_t)(h->cg_trace_n++)*NQS_CG_TRACE_WORDS;
  if (!h->cg_attr_set)
  { // same shared-memory carve-out as the S*v kernels on either side: no SM reconfiguration between the launches of an iteration
    // (function attributes are per device: kept per handle, not per process)
    cudaFuncSetAttribute(cg_fused_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    h->cg_attr_set = true;
  }
  // The kernel spins on a software grid barrier, so every CTA must be resident: a COOPERATIVE launch makes the driver guarantee
  // that (or refuse the launch) whatever else runs on the device.  If cooperative launches are not available the plain launch
  // is kept; the barrier then gives up after NQS_CG_BARRIER_TIMEOUT_NS and the host reports NQS_ERR_CUDA instead of hanging.
  if (h->cg_coop != 0)
  {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)cg_ctas(h), 1, 1); cfg.blockDim = dim3(NQS_CG_THREADS, 1, 1); cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, cg_fused_kernel, a);
    if (e == cudaSuccess) { h->cg_coop = 1; check_launch(h, "cg_fused_kernel"); return; }
    cudaGetLastError();
    if (h->cg_coop == 1)
      throw Error(NQS_ERR_CUDA, std::string("cooperative launch of cg_fused_kernel failed: ")+cudaGetErrorString(e));
    h->cg_coop = 0;
  }
  cg_fused_kernel<<<cg_ctas(h), NQS_CG_THREADS, 0, h->stream>>>(a);
  check_launch(h, "cg_fused_kernel");
}

// ref: ConjugateGradient::solve(SMatrixFunctor_, F, dx), conjugate_gradient.cuh:29-74, warm start in dx.
// The launch-per-iteration solve WITHOUT polling: S x0 + INIT, then as many iterations as the previous solve needed plus two,
// and the scalars on their way to pinned memory -- nothing is synchronised.  Iterations past convergence exit at entry (`done`);
// the INIT launch also makes the finite-energy check.  Returns the number of iterations enqueued.  After the caller's next
// synchronisation cg_finish_async() reads the scalars and, in the rare case that the solve needs more iterations than were
// enqueued, runs them with the polling loop of cg_solve.
int cg_solve_async_begin(nqs_handle * h, double lambda, double tol, int max_iter, int fixed_iters)
{
  CgScalars init;
  std::memset(&init, 0, sizeof(init));
  init.tol2 = tol*tol;
  init.fixed = fixed_iters > 0 ? 1 : 0;
  std::memcpy((char*)h->pinned+PIN_CG_INIT, &init, sizeof(init));
  NQS_CUDA(cudaMemcpyAsync(h->scal.p, (char*)h->pinned+PIN_CG_INIT, sizeof(CgScalars), cudaMemcpyHostToDevice, h->stream));
  int * done = &h->scal.p->done;
  int nparts = matvec_passes(h, h->dx.p, nullptr);
  h->cg_check_finite = true;
  launch_cg_fused(h, CG_MODE_INIT, nparts, lambda, h->dx.p);
  h->cg_check_finite = false;
  const int n_max = fixed_iters > 0 ? fixed_iters : max_iter;
  const int n = fixed_iters > 0 ? n_max : std::max(1, std::min(n_max, h->cg_prev_iters+2));
  for (int q = 0; q < n; ++q)
  {
    nparts = matvec_passes(h, h->pvec.p, done);
    launch_cg_fused(h, CG_MODE_ITER, nparts, lambda, h->pvec.p);
  }
  NQS_CUDA(cudaMemcpyAsync((char*)h->pinned+PIN_CG_SNAP, h->scal.p, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
  return n;
}

void cg_check_errors(nqs_handle * h, const CgScalars & s)
{
  if (s.peer_timeout)
    throw Error(NQS_ERR_NCCL, "in-kernel NVLink exchange timed out after 20 s: a peer rank did not reach the same CG iteration");
  if (s.barrier_timeout)
  {
    cudaMemsetAsync(h->cgbar.p, 0, sizeof(unsigned int), h->stream);
    throw Error(NQS_ERR_CUDA, "grid barrier of cg_fused_kernel timed out after 20 s: its CTAs were not co-resident");
  }
}

// after a synchronisation.  Returns true when more iterations had to be run here (the caller's update, enqueued behind the first
// batch, declined on the device and must be enqueued again).
bool cg_finish_async(nqs_handle * h, double lambda, int max_iter, int fixed_iters, int enq, nqs_sr_stats * st, bool * nonfinite)
{
  CgScalars * snap = reinterpret_cast<CgScalars*>((char*)h->pinned+PIN_CG_SNAP);
  CgScalars s = *snap;
  cg_check_errors(h, s);
  *nonfinite = s.nonfinite != 0;
  bool more = false;
  const int n_max = fixed_iters > 0 ? fixed_iters : max_iter;
  int * done = &h->scal.p->done;
  while (!s.nonfinite && !s.done && enq < n_max)
  {
    more = true;
    // the update enqueued behind the first batch declined on the device (`done` was not set): parameters, theta and the
    // hidden-unit factors are still those of this step, whatever do_evolve() recorded on the host
    if (h->struct_sv) h->hidden_valid = true;
    const int n_here = std::min(2, n_max-enq);
    for (int q = 0; q < n_here; ++q)
    {
      const int nparts = matvec_passes(h, h->pvec.p, done);
      launch_cg_fused(h, CG_MODE_ITER, nparts, lambda, h->pvec.p);
    }
    enq += n_here;
    NQS_CUDA(cudaMemcpyAsync(snap, h->scal.p, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    s = *snap;
    cg_check_errors(h, s);
  }
  if (fixed_iters <= 0 && !s.nonfinite) h->cg_prev_iters = std::max(1, s.iters);
  if (st) { st->cg_iters = s.iters; st->cg_res2 = s.res2; st->cg_rhs2 = s.rhs2; }
  return more;
}

bool cgp_usable(const nqs_handle * h)
{ // multi-GPU: the packet exchange inside the kernel needs mapped peers and every rank on the persistent path
  return h->cgp_ok && (h->comm == nullptr || (h->p2p_ok && h->cgp_peers_agree));
}

// ref: ConjugateGradient::solve -- the whole solve as ONE launch (cg_persist.cuh).  Nothing is synchronised here: the scalars
// travel to pinned memory behind the kernel and collect_cg() reads them after the caller's next synchronisation.
void cg_solve_persistent(nqs_handle * h, double lambda, double tol, int max_iter, int fixed_iters)
{
  CgScalars init;
  std::memset(&init, 0, sizeof(init));
  init.tol2 = tol*tol;
  init.fixed = fixed_iters > 0 ? 1 : 0;
  std::memcpy((char*)h->pinned+PIN_CG_INIT, &init, sizeof(init));
  NQS_CUDA(cudaMemcpyAsync(h->scal.p, (char*)h->pinned+PIN_CG_INIT, sizeof(CgScalars), cudaMemcpyHostToDevice, h->stream));
  CgpArgs a;
  std::memset(&a, 0, sizeof(a));
  a.K = h->K; a.P = h->P; a.O = h->O.p; a.part = h->part.p; a.pc = h->sv_pc; a.rows_per_cluster = h->sv_rpc;
  a.nslot = h->sv_nslot; a.slot_bytes = (unsigned int)h->sv_slot_bytes; a.depth = h->sv_depth;
  a.inv_ktot = 1.0/(double)h->Ktot; a.lambda = lambda; a.tol2 = tol*tol; a.fixed_iters = fixed_iters; a.max_iter = max_iter;
  a.aO = h->aO.p; a.diag = h->diag.p; a.F = h->F.p; a.x = h->dx.p; a.r = h->r.p; a.pb[0] = h->t.p; a.pb[1] = h->pvec.p; a.zv = h->z.p; a.sc = h->scal.p;
  a.slots = h->slots.p; a.barrier = h->cgbar.p; a.hsums = h->hsall.p;
  a.n_ranks = 1; a.rank = 0; a.epoch0 = h->p2p_epoch;
  if (h->comm != nullptr)
  {
    a.n_ranks = h->n_ranks; a.rank = h->rank;
    for (int r = 0; r < h->n_ranks; ++r)
    {
      a.peer_ll1[r] = reinterpret_cast<uint4*>((char*)h->peer_base[r]+h->xbuf_ll1_off);
      a.peer_ll2[r] = reinterpret_cast<uint4*>((char*)h->peer_base[r]+h->xbuf_ll2_off);
    }
  }
  a.trace = h->cg_trace.p;
  a.trace_max = (int)(NQS_CG_TRACE_MAX*NQS_CG_TRACE_WORDS/NQS_CGP_TRACE_WORDS);
  {
    Span sp(h, TAG_ROWS);
    cudaError_t e = cudaErrorUnknown;
    if (h->cgp_coop != 0)
    { // cooperative + cluster launch: the driver guarantees (or refuses) co-residency of all clusters
      e = cgp_launch(h->sv_cpt, h->sv_defer, a, h->sv_cs, h->sv_nclusters, h->sv_nt, h->sv_smem, h->stream, true, nullptr);
      if (e == cudaSuccess) h->cgp_coop = 1;
      else
      {
        cudaGetLastError();
        if (h->cgp_coop == 1) throw Error(NQS_ERR_CUDA, std::string("cooperative launch of cg_persist_kernel failed: ")+cudaGetErrorString(e));
        h->cgp_coop = 0;
      }
    }
    if (h->cgp_coop == 0)
      NQS_CUDA(cgp_launch(h->sv_cpt, h->sv_defer, a, h->sv_cs, h->sv_nclusters, h->sv_nt, h->sv_smem, h->stream, false, nullptr));
    check_launch(h, "cg_persist_kernel");
  }
  NQS_CUDA(cudaMemcpyAsync((char*)h->pinned+PIN_CG_SNAP, h->scal.p, sizeof(CgScalars), cudaMemcpyDeviceToHost, h->stream));
  h->cg_inflight = true;
}

// after a synchronisation: scalars of the persistent solve -> stats, exchange epoch, errors
void collect_cg(nqs_handle * h, nqs_sr_stats * st, bool * nonfinite)
{
  if (!h->cg_inflight) return;
  h->cg_inflight = false;
  CgScalars s;
  std::memcpy(&s, (char*)h->pinned+PIN_CG_SNAP, sizeof(s));
  if (nonfinite) *nonfinite = s.nonfinite != 0;
  if (s.barrier_timeout || s.peer_timeout)
  {
    cudaMemsetAsync(h->cgbar.p, 0, sizeof(unsigned int), h->stream);
    cudaStreamSynchronize(h->stream);
    if (s.peer_timeout) throw Error(NQS_ERR_NCCL, "in-kernel NVLink exchange timed out after 20 s: a peer rank did not reach the same CG iteration");
    throw Error(NQS_ERR_CUDA, "grid barrier of cg_persist_kernel timed out after 20 s: its CTAs were not co-resident (another kernel held SMs "
      "and cooperative launches are unavailable)");
  }
  const int products = s.nonfinite ? 0 : s.iters+1;
  h->p2p_epoch += (unsigned int)products;
  h->timing.rows_count = products;
  if (!s.nonfinite && !s.fixed) h->cg_prev_iters = std::max(1, s.iters);
  if (st) { st->cg_iters = s.iters; st->cg_res2 = s.res2; st->cg_rhs2 = s.rhs2; }
}

void do_evolve(nqs_handle * h, const cd * dx_dev, double lr, const CgScalars * sc = nullptr, int need_done = 0)
{
  h->theta_matches_O = false; h->hidden_valid = false; h->o_pending = false;
  invalidate_tables(h);
  update_params_kernel<<<grid_for(h->P, 256, 148*4), 256, 0, h->stream>>>(h->N, h->M, h->trsymm ? -1 : h->model, h->P, dx_dev, lr, var_ptr(h), sc, need_done);   // tied variables: plain update (ref update_parameters)
  check_launch(h, "update_params_kernel");
  expand_vars(h);
  build_tables_async(h);   // the next sweep needs them anyway; their bound rides on the caller's final read-back
  NQS_CUDA(cudaMemsetAsync(h->fresh.p, 0, (size_t)h->K, h->stream)); // lnpsi0 is now the pre-update value (ref keeps it)
  // ref update_variables tail (:161-169): theta and sa re-derived for the current spins; lnpsi0 is NOT refreshed
  launch_theta(h, h->spins.p, h->spins.p, h->theta.p, h->sa.p, nullptr);
}

void do_sweeps(nqs_handle * h, int n_sweeps)
{
  NQS_REQUIRE(h->initialized, NQS_ERR_STATE, "nqs_do_mcmc_steps before nqs_initialize / nqs_warm_up");
  NQS_REQUIRE(n_sweeps >= 0, NQS_ERR_INVALID, "n_sweeps < 0");
  const long long nsteps = (long long)n_sweeps*h->N;
  if (nsteps == 0) return;
  launch_sweep(h, nsteps);
  h->flip_index = h->order_host[(h->pos+h->N-1)%h->N];   // the machine's index_ is the last visited site
}

void do_initialize(nqs_handle * h, const int8_t * spins_host)
{
  h->theta_matches_O = false; h->hidden_valid = false; h->o_pending = false;
  std::vector<int8_t> s((size_t)h->K*h->N, 1);
  if (spins_host) std::memcpy(s.data(), spins_host, s.size());
  else if (h->cfg.J > 0) // Neel, ref impl_hamiltonians.cuh:196-201
    for (long long k = 0; k < h->K; ++k)
      for (int i = 0; i < h->N; ++i)
        s[(size_t)k*h->N+i] = (i%2 == 0) ? 1 : -1;
  NQS_CUDA(cudaMemcpyAsync(h->spins.p, s.data(), s.size(), cudaMemcpyHostToDevice, h->stream));
  launch_theta(h, h->spins.p, h->spins.p, h->theta.p, h->sa.p, h->lnpsi0.p);
  NQS_CUDA(cudaMemsetAsync(h->fresh.p, 1, (size_t)h->K, h->stream));
  NQS_CUDA(cudaStreamSynchronize(h->stream));
  h->initialized = true;
}

void build_order(nqs_handle * h)
{
  std::vector<int> ring;
  const int N = h->N;
  if (h->cfg.order == NQS_ORDER_CHECKERBOARD)
  { // ring 0 -> evens -> odds -> 0, pointer advanced before use (ref impl_hamiltonians.cuh:163-180,209-210)
    for (int i = 0; i < N; i += 2) ring.push_back(i);
    for (int i = 1; i < N; i += 2) ring.push_back(i);
  }
  else
    for (int i = 0; i < N; ++i) ring.push_back(i);
  std::vector<int> ord(N);
  for (int t = 0; t < N; ++t) ord[t] = ring[(t+1)%N];
  NQS_CUDA(cudaMemcpy(h->order.p, ord.data(), sizeof(int)*N, cudaMemcpyHostToDevice));
  h->order_host = ord;
}

void build_J(nqs_handle * h)
{ // ref impl_hamiltonians.cuh:136-161
  const int L = h->N;
  std::vector<double> Jm((size_t)L*L, 0.0);
  for (int i = 0; i < L; ++i)
    for (int j = i+1; j < L; ++j)
    {
      double dist = (double)(j-i);
      if (h->cfg.pbc) dist = ((j-i) < L/2) ? (double)(j-i) : (double)(L-(j-i));
      Jm[(size_t)i*L+j] = h->cfg.J*std::pow(dist, -h->cfg.alpha);
      Jm[(size_t)j*L+i] = Jm[(size_t)i*L+j];
    }
  NQS_CUDA(cudaMemcpy(h->Jmat.p, Jm.data(), sizeof(double)*Jm.size(), cudaMemcpyHostToDevice));
}

void upload_params(nqs_handle * h, const std::vector<std::complex<double> > & v)
{
  invalidate_tables(h);
  h->theta_matches_O = false; h->hidden_valid = false; h->o_pending = false;
  NQS_CUDA(cudaMemcpyAsync(var_ptr(h), v.data(), sizeof(cd)*v.size(), cudaMemcpyHostToDevice, h->stream));
  expand_vars(h);
  NQS_CUDA(cudaStreamSynchronize(h->stream));
}
std::vector<std::complex<double> > download_params(nqs_handle * h)
{
  std::vector<std::complex<double> > v((size_t)h->P);
  NQS_CUDA(cudaMemcpyAsync(v.data(), var_ptr(h), sizeof(cd)*v.size(), cudaMemcpyDeviceToHost, h->stream));
  NQS_CUDA(cudaStreamSynchronize(h->stream));
  return v;
}

// O [K][P], the CG vectors and the reduction scratch (skipped for sampler-only handles until nqs_enable_sr)
void alloc_sr(nqs_handle * h)
{
  if (h->aO.p != nullptr) return;
  plan_cols(h);
  if (h->cols_ok)
  { // hidden-unit factors T [K][M] (+ L, FFNN) for the tensor-core GEMMs (SR setup; structured S*v)
    h->Tm.alloc((size_t)h->K*h->M);
    if (h->model == MODEL_FFNN) { h->Lm.alloc((size_t)h->K*h->M); h->vnat.alloc((size_t)h->N*h->M); }
    h->abs2.alloc((size_t)h->sc_nchunks*3*h->M);
  }
  // structured S*v: no O [K][P]; it is allocated only if nqs_log_derivs asks for it
  if (h->cfg.flags & NQS_FLAG_STRUCTURED_SV) plan_struct(h);
  else
  {
    plan_sv(h);
    // Rows too wide for the one-pass kernel (P/16 columns beyond the register file of a 16-CTA cluster: RBM alpha=4 at N=256,
    // P = 263 424): the explicit formulation would stream O TWICE per product (two-pass kernels, 2 x 34.5 GB per product and
    // rank at 8 GPUs: 106 of 138 ms per step).  The factor form of the same product, two tensor-core GEMMs on [K][N] spins and
    // [K][M] tanh(theta), needs no O at all (35 ms per step there), so it takes over -- unless the caller asked for the two-pass
    // kernels (NQS_FLAG_TWO_PASS_SV) or NQS_AUTO_STRUCTURED=0.  kernel_variant("sv") says which ran.
    const char * au = std::getenv("NQS_AUTO_STRUCTURED");
    if (!h->sv_ok && h->cols_ok && !h->trsymm && !(h->cfg.flags & (NQS_FLAG_TWO_PASS_SV | NQS_FLAG_SETUP_FROM_O | NQS_FLAG_NO_DMMA)) &&
        !(au && std::atoi(au) == 0))
    {
      h->cfg.flags |= NQS_FLAG_STRUCTURED_SV;
      plan_struct(h);
      h->variant_sv += "_auto(P/16_columns_exceed_the_one-pass_kernel)";
    }
    else h->O.alloc((size_t)h->K*(size_t)h->P);
  }
  h->aO.alloc(h->P); h->F.alloc(h->P); h->dx.alloc(h->P); h->r.alloc(h->P); h->pvec.alloc(h->P); h->z.alloc(h->P); h->t.alloc(h->P);
  h->zk.alloc(h->K); h->diag.alloc(h->P);
  const long long ctiles = (h->P+NQS_COL_THREADS-1)/NQS_COL_THREADS;
  long long nrb = (2LL*16*h->sm_count+ctiles-1)/ctiles;
  nrb = std::max<long long>(1, std::min<long long>(nrb, std::min<long long>(64, (h->K+63)/64)));
  h->nrb = (int)nrb;
  h->rows_per_block = (h->K+nrb-1)/nrb;
  h->nrb = (int)((h->K+h->rows_per_block-1)/h->rows_per_block);
  // O generated inside the first S*v of the CG (sv_fused_kernel, GEN): RBM, one-pass kernel with the software pipeline, N a
  // multiple of 16 (TMA rows of the int8 spins), factors available.  OPT-IN (NQS_SV_GEN=1), measured twice at cfg3 and slower both
  // times than writer (1.34 ms) + read pass (1.43 ms): (round 1) with the geometry of the other products every element costs two
  // shared-memory loads and two multiplications, the launch is issue-bound at 3.5 ms; (round 2) with a consumer thread count that
  // is a multiple of M (256 threads, 16-CTA clusters, 9 columns per thread) an element is one sign flip of one T value, but a
  // cluster then handles 1820 rows of only 2072 columns per CTA and the per-row DSMEM exchange dominates: ~4.5 ms (step 25.0 vs
  // 14.6 ms).  512 consumer threads (the geometry of the other products) do not fit: 17 warps x 120 registers exceed the register
  // file's allocation unit.  The separate O writer stays the default.
  { const char * ng = std::getenv("NQS_SV_GEN");
    h->gen_ok = h->sv_ok && h->sv_defer && h->cols_ok && h->model == MODEL_RBM && h->N%16 == 0 && h->M < 65535 && h->N < 32766 &&
      (ng && std::atoi(ng) != 0) &&
      (size_t)NQS_SV_MAX_SLOTS*sv_gen_slot_bytes(h->N, h->M)+NQS_SV_TAIL_BYTES <= h->smem_optin;
    if (h->gen_ok)
    { // geometry of the generating launch: by default that of the other products; if a consumer thread count NT <= 480 that is
      // a multiple of M exists, the cluster size that covers the most SMs with at most 9 columns per thread
      h->gen_cs = h->sv_cs; h->gen_cpt = h->sv_cpt; h->gen_nt = h->sv_nt; h->gen_nclusters = h->sv_nclusters;
      h->gen_pc = h->sv_pc; h->gen_rpc = h->sv_rpc; h->gen_q = (h->sv_nt%h->M == 0) ? h->sv_nt/h->M : 0;
      int step = h->M;
      while (step%32 != 0) step *= 2;
      const int nt = (step <= 480) ? 480/step*step : 0;      // (512 + 32 threads would need <= 112 registers: the allocation unit is 512 per warp)
      const char * g2 = std::getenv("NQS_SV_GEN_PLAN");
      if (nt > 0 && !(g2 && std::atoi(g2) == 0))
      {
        static const int sizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16};
        int best_cs = 0, best_cov = 0, best_cpt = 0;
        for (int cs : sizes)
        {
          const long long pc = (h->P+cs-1)/cs;
          const int cpt = (int)((pc+nt-1)/nt);
          if (cpt > 9) continue;
          const int cov = cs*(h->sm_count/cs)+(cs == h->sv_cs ? 1000 : 0);      // the cluster size of the other products if it fits
          if (cov > best_cov) { best_cov = cov; best_cs = cs; best_cpt = cpt; }
        }
        if (best_cs > 0)
        {
          h->gen_cs = best_cs; h->gen_cpt = best_cpt; h->gen_nt = nt; h->gen_pc = (h->P+best_cs-1)/best_cs;
          const int ncl = std::max(1, std::min<int>(h->sm_count/best_cs, (int)std::min<long long>(h->K, 1<<20)));
          h->gen_rpc = (h->K+ncl-1)/ncl;
          h->gen_nclusters = (int)((h->K+h->gen_rpc-1)/h->gen_rpc);
          h->gen_q = nt/h->M;
        }
      }
    } }
  plan_cgp(h);
  if (h->cgp_ok) h->variant_sv += "_persistentcg";
  h->part.alloc(std::max(std::max(std::max((size_t)h->nrb*5*h->P, (size_t)h->sv_nclusters*2*h->P), (size_t)h->sc_nchunks*4*h->P),
                         (size_t)h->gen_nclusters*2*h->P));
  h->sums.alloc((size_t)5*h->P+3);
  h->hsall.alloc(4);
  h->traw.alloc((size_t)2*h->P);
  h->slots.alloc((size_t)2*NQS_CGP_MAX_CTAS*NQS_CGP_NVALS);   // sized for the persistent kernel's grid (cg_fused_kernel uses the first 148 of each half)
  h->cgbar.alloc(1);
  NQS_CUDA(cudaMemset(h->cgbar.p, 0, sizeof(unsigned int)));
  h->scal.alloc(1);
  NQS_CUDA(cudaMemset(h->scal.p, 0, sizeof(CgScalars)));
  { const char * tr = std::getenv("NQS_CG_TRACE");
    if (tr && std::atoi(tr) != 0)
    {
      h->cg_trace.alloc((size_t)NQS_CG_TRACE_MAX*NQS_CG_TRACE_WORDS);
      NQS_CUDA(cudaMemset(h->cg_trace.p, 0, (size_t)NQS_CG_TRACE_MAX*NQS_CG_TRACE_WORDS*sizeof(unsigned long long)));
    } }
  NQS_CUDA(cudaMemset(h->dx.p, 0, sizeof(cd)*h->P)); // CG warm start is zero only at construction (ref impl_optimizer.cuh:55)
}

struct ParamFile { const char * suffix; long long off, count, row; const char * name; };
std::vector<ParamFile> param_files(const nqs_handle * h)
{ // ref save/load: RBM Dw/Da/Db (:225-232,281-286); FFNN Dw1/Dw2(=w1o)/Db1 (:931-937,985-991)
  const long long NM = (long long)h->N*h->M;
  if (h->trsymm)   // ref RBMTrSymm::save / load (:474-517): every variable in ONE file named by the prefix itself
    return {{"", 0, h->P, h->P+1, "variables"}};
  if (h->model == MODEL_RBM)
    return {{"Dw.dat", 0, NM, h->M, "w"}, {"Da.dat", NM, h->N, h->N, "a"}, {"Db.dat", NM+h->N, h->M, h->M, "b"}};
  return {{"Dw1.dat", 0, NM, h->M, "w1"}, {"Dw2.dat", NM+h->M, h->M, h->M, "w2"}, {"Db1.dat", NM, h->M, h->M, "b1"}};
}
} // namespace

// =====================================================================================================================
extern "C"
{
int32_t nqs_abi_version(void) { return NQS_B200_ABI_VERSION; }

const char * nqs_last_error(const nqs_handle * h) { return h ? h->err.c_str() : g_create_error.c_str(); }

nqs_status nqs_sr_options_default(nqs_sr_options * o)
{
  if (!o) return NQS_ERR_INVALID;
  o->lr = 1e-2;        // ref default -lr, gpu/src/LICH-train_rbm.cu:36
  o->tol = 1e-5;       // gpu/include/impl_optimizer.cuh:60
  o->max_iter = 1000;  // gpu/include/conjugate_gradient.cuh:19
  o->fixed_iters = 0;
  o->lambda = -1.0;
  o->n_mc_steps = 1;   // ref default -nms
  o->apply_update = 1;
  return NQS_OK;
}

nqs_status nqs_create(const nqs_config * cfg, nqs_handle ** out)
{
  if (out) *out = nullptr;
  nqs_handle * h = nullptr;
  nqs_status rc = guarded(nullptr, [&]()
  {
    NQS_REQUIRE(cfg && out, NQS_ERR_INVALID, "nqs_create: null argument");
    NQS_REQUIRE(cfg->abi_version == NQS_B200_ABI_VERSION, NQS_ERR_INVALID, "nqs_create: abi_version mismatch");
    NQS_REQUIRE(cfg->model >= NQS_MODEL_RBM && cfg->model <= NQS_MODEL_FFNNTRSYMM, NQS_ERR_INVALID, "nqs_create: unknown model");
    NQS_REQUIRE((cfg->model != NQS_MODEL_RBMTRSYMM && cfg->model != NQS_MODEL_FFNNTRSYMM) || cfg->n_hiddens%cfg->n_inputs == 0, NQS_ERR_INVALID,
      "nqs_create: a translation-symmetric ansatz needs n_hiddens = alpha * n_inputs (the expanded width)");
    NQS_REQUIRE(cfg->model != NQS_MODEL_RBMZ2PRSYMM || cfg->n_hiddens%4 == 0, NQS_ERR_INVALID,
      "nqs_create: the Z2/parity-symmetric RBM needs n_hiddens = 4 * alpha (the expanded width)");
    NQS_REQUIRE(cfg->n_inputs >= 1 && cfg->n_hiddens >= 1 && cfg->n_chains >= 1, NQS_ERR_INVALID, "nqs_create: sizes must be >= 1");
    NQS_REQUIRE(!(cfg->pbc && cfg->n_inputs%2 == 1), NQS_ERR_INVALID, "kL%2 == 1 (set \"isPBC\" to \"false\".)"); // ref :141-142
    NQS_REQUIRE(cfg->order == NQS_ORDER_CHECKERBOARD || cfg->order == NQS_ORDER_SEQUENTIAL, NQS_ERR_INVALID, "nqs_create: unknown order");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw Error(NQS_ERR_CUDA, std::string("no CUDA device (libnqs_b200 has no CPU fallback): ")+cudaGetErrorString(e));
    NQS_REQUIRE(cfg->device >= 0 && cfg->device < ndev, NQS_ERR_INVALID, "# error ---> dev >= nDevice"); // ref LICH-train_rbm.cu:66-70
    NQS_CUDA(cudaSetDevice(cfg->device));
    h = new nqs_handle();
    h->cfg = *cfg;
    { // callers that cannot pass flags (the reference-compatible CLI, pynqs) opt into the structured S*v through the environment
      const char * es = std::getenv("NQS_STRUCTURED_SV");
      if (es && std::atoi(es) != 0 && cfg->n_inputs <= 256 &&
          !(cfg->flags & (NQS_FLAG_SETUP_FROM_O | NQS_FLAG_TWO_PASS_SV | NQS_FLAG_NO_DMMA)))
        h->cfg.flags |= NQS_FLAG_STRUCTURED_SV;
    }
    h->N = cfg->n_inputs; h->M = cfg->n_hiddens; h->model = cfg->model;
    if (cfg->model == NQS_MODEL_RBMTRSYMM) { h->model = MODEL_RBM; h->trsymm = true; h->tied = TIED_RBM_TR; h->alpha_f = h->M/h->N; }
    if (cfg->model == NQS_MODEL_RBMZ2PRSYMM) { h->model = MODEL_RBM; h->trsymm = true; h->tied = TIED_RBM_Z2PR; h->alpha_f = h->M/4; }
    if (cfg->model == NQS_MODEL_FFNNTRSYMM) { h->model = MODEL_FFNN; h->trsymm = true; h->tied = TIED_FFNN_TR; h->alpha_f = h->M/h->N; }
    h->K = cfg->n_chains; h->Ktot = cfg->n_chains_total > 0 ? cfg->n_chains_total : cfg->n_chains; h->koff = cfg->chain_offset;
    NQS_REQUIRE(h->Ktot >= h->K, NQS_ERR_INVALID, "n_chains_total < n_chains");
    h->P = (h->model == MODEL_RBM) ? (long long)h->N*h->M+h->N+h->M : (long long)h->N*h->M+2*h->M;
    h->Pfull = h->P;
    if (h->tied == TIED_RBM_TR) h->P = (long long)h->N*h->alpha_f+1+h->alpha_f;
    if (h->tied == TIED_RBM_Z2PR) h->P = (long long)h->N*h->alpha_f+h->alpha_f;
    if (h->tied == TIED_FFNN_TR) h->P = (long long)h->N*h->alpha_f+2*h->alpha_f;
    std::memset(&h->timing, 0, sizeof(h->timing));
    cudaDeviceProp prop;
    NQS_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    NQS_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));

    for (int i = 0; i < 8; ++i) NQS_CUDA(cudaEventCreate(&h->ev[i]));

    h->ev_ok = true;
    NQS_CUDA(cudaMallocHost(&h->pinned, 4096));
    const size_t KM = (size_t)h->K*h->M, KN = (size_t)h->K*h->N;
    h->params.alloc(h->Pfull); h->theta.alloc(KM); h->tmp_theta.alloc(0);
    if (h->trsymm) { h->vars.alloc(h->P); NQS_CUDA(cudaMemset(h->vars.p, 0, sizeof(cd)*h->P)); }
    h->lnpsi0.alloc(h->K); h->lnpsi1.alloc(h->K); h->sa.alloc(h->K); h->htilda.alloc(h->K);
    h->spins.alloc(KN); h->tmp_spins.alloc(KN);
    h->Jmat.alloc((size_t)h->N*h->N); h->order.alloc(h->N);
    h->fresh.alloc(h->K);
    NQS_CUDA(cudaMemset(h->fresh.p, 0, (size_t)h->K));
    // product-form tables: the sweep keeps its state in registers (M <= 512: 16 hidden-unit slots per lane; M <= 1024: two warps
    // per chain); the local-energy kernel keeps tanh(theta) in shared memory and works up to M = 2048
    if (!(cfg->flags & NQS_FLAG_FORCE_GENERIC) && h->M <= 2048)
    { // (FNN: the same cosh 2W / sinh 2W / 2W tables serve ffnn_fast_kernels.cuh)
      int jpl = 1;
      while (32*jpl < h->M) jpl <<= 1;
      h->jpl = jpl; h->mpad = 32*jpl;
      const size_t nm = (size_t)h->N*h->mpad;
      h->ftab_a.alloc(nm); h->ftab_b.alloc(nm); h->ctab_a.alloc(nm); h->ctab_b.alloc(nm);
      if (h->model == MODEL_RBM) { h->ftab32.alloc(nm); h->f32_stats.alloc(2); NQS_CUDA(cudaMemset(h->f32_stats.p, 0, 16)); }
      h->npad32 = ((h->N+31)/32)*32;
      h->ctabT_a.alloc((size_t)h->M*h->npad32); h->ctabT_b.alloc((size_t)h->M*h->npad32); h->w2.alloc(nm); h->afac.alloc((size_t)2*h->N); h->aexp.alloc((size_t)2*h->N); h->bound.alloc(1);
    }
    NQS_CUDA(cudaMemset(h->params.p, 0, sizeof(cd)*h->Pfull));
    NQS_CUDA(cudaMemset(h->spins.p, 0, KN));   // like the reference's zero-initialised spinStates_dev_
    NQS_CUDA(cudaMemset(h->theta.p, 0, sizeof(cd)*KM));
    NQS_CUDA(cudaMemset(h->lnpsi0.p, 0, sizeof(cd)*h->K));
    NQS_CUDA(cudaMemset(h->sa.p, 0, sizeof(cd)*h->K));
    NQS_CUDA(cudaMemset(h->htilda.p, 0, sizeof(cd)*h->K));
    if (cfg->max_predrawn_steps > 0) h->uniforms.alloc((size_t)cfg->max_predrawn_steps*h->K);
    if (!(cfg->flags & NQS_FLAG_NO_SR)) alloc_sr(h);
    build_order(h);
    build_J(h);
    NQS_CUDA(cudaDeviceSynchronize());
    *out = h;
  });
  if (rc != NQS_OK && h) { nqs_destroy(h); }
  return rc;
}

void nqs_destroy(nqs_handle * h)
{
  if (!h) return;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  dump_cg_trace(h);
  for (int r = 0; r < 16; ++r)
    if (h->peer_base[r] && r != h->rank) cudaIpcCloseMemHandle(h->peer_base[r]);
  if (h->xbuf) cudaFree(h->xbuf);
  if (h->comm && g_nccl.commDestroy) g_nccl.commDestroy(h->comm);
  if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
  if (h->ev_ok) for (int i = 0; i < 8; ++i) cudaEventDestroy(h->ev[i]);

  for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
  if (h->pinned) cudaFreeHost(h->pinned);
  delete h;
}

nqs_status nqs_sync(nqs_handle * h)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]() { NQS_CUDA(cudaStreamSynchronize(h->stream)); });
}

nqs_status nqs_n_variables(const nqs_handle * h, int64_t * P)
{
  if (!h || !P) return NQS_ERR_INVALID;
  *P = h->P;
  return NQS_OK;
}

nqs_status nqs_set_params(nqs_handle * h, const nqs_cdouble * params, int64_t P)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(params && P == h->P, NQS_ERR_INVALID, "nqs_set_params: P mismatch");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    invalidate_tables(h);
    NQS_CUDA(cudaMemcpyAsync(var_ptr(h), params, sizeof(cd)*P, cudaMemcpyHostToDevice, h->stream));
    expand_vars(h);
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    h->theta_matches_O = false; h->hidden_valid = false; h->o_pending = false;
  });
}

nqs_status nqs_get_params(nqs_handle * h, nqs_cdouble * params, int64_t P)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(params && P == h->P, NQS_ERR_INVALID, "nqs_get_params: P mismatch");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaMemcpyAsync(params, var_ptr(h), sizeof(cd)*P, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_init_params_random(nqs_handle * h, uint64_t seed)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    std::mt19937_64 ran(seed);
    const int N = h->N, M = h->M;
    std::vector<std::complex<double> > v((size_t)h->P);
    const long long NM = (long long)N*M;
    if (h->tied == TIED_RBM_Z2PR)
    { // ref ctor impl_neural_quantum_state.cuh:562-579
      const int al = h->alpha_f;
      std::normal_distribution<double> randw(0, std::sqrt(1.0/(4*al+N))), randb(0, std::sqrt(1.0/(4*al)));
      for (long long i = 0; i < (long long)N*al; ++i) { const double re = 1e-1*randw(ran), im = 1e-1*randw(ran); v[i] = {re, im}; }
      for (int j = 0; j < al; ++j) { const double re = 1e-1*randb(ran), im = 1e-1*randb(ran); v[(size_t)N*al+j] = {re, im}; }
    }
    else if (h->tied == TIED_FFNN_TR)
    { // ref ctor impl_neural_quantum_state.cuh:1041-1061
      const int al = h->alpha_f;
      std::normal_distribution<double> randwi1(0, std::sqrt(1.0/((1+al)*N))), randw1o(0, std::sqrt(1.0/(al*N)));
      for (long long i = 0; i < (long long)N*al; ++i) { const double re = randwi1(ran), im = 1e-1*randwi1(ran); v[i] = {re, im}; }
      for (int j = 0; j < al; ++j) v[(size_t)N*al+j] = {0.0, 0.0};
      for (int j = 0; j < al; ++j) { const double re = randw1o(ran), im = 1e-1*randw1o(ran); v[(size_t)N*al+al+j] = {re, im}; }
    }
    else if (h->trsymm)
    { // ref ctor impl_neural_quantum_state.cuh:325-345
      const int al = h->alpha_f;
      std::normal_distribution<double> randw(0, std::sqrt(1.0/((1+al)*N))), randb(0, std::sqrt(1.0/(N*al)));
      for (long long i = 0; i < (long long)N*al; ++i) { const double re = 1e-1*randw(ran), im = 1e-1*randw(ran); v[i] = {re, im}; }
      v[(size_t)N*al] = {0.0, 0.0};
      for (int j = 0; j < al; ++j) { const double re = 1e-1*randb(ran), im = 1e-1*randb(ran); v[(size_t)N*al+1+j] = {re, im}; }
    }
    else if (h->model == MODEL_RBM)
    { // ref ctor impl_neural_quantum_state.cuh:30-48
      std::normal_distribution<double> randw(0, std::sqrt(1.0/(N+M))), randb(0, std::sqrt(1.0/M));
      for (long long i = 0; i < NM; ++i) { const double re = 1e-1*randw(ran), im = 1e-1*randw(ran); v[i] = {re, im}; }
      for (int i = 0; i < N; ++i) v[NM+i] = {0.0, 0.0};
      for (int j = 0; j < M; ++j) { const double re = 1e-1*randb(ran), im = 1e-1*randb(ran); v[NM+N+j] = {re, im}; }
    }
    else
    { // ref ctor :766-783
      std::normal_distribution<double> randwi1(0, std::sqrt(1.0/(N+M))), randw1o(0, std::sqrt(1.0/M));
      for (long long i = 0; i < NM; ++i) { const double re = randwi1(ran), im = 1e-1*randwi1(ran); v[i] = {re, im}; }
      for (int j = 0; j < M; ++j) v[NM+j] = {0.0, 0.0};
      for (int j = 0; j < M; ++j) { const double re = randw1o(ran), im = 1e-1*randw1o(ran); v[NM+M+j] = {re, im}; }
    }
    upload_params(h, v);
  });
}

nqs_status nqs_load_params(nqs_handle * h, const char * prefix)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(prefix, NQS_ERR_INVALID, "nqs_load_params: null prefix");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    std::vector<std::complex<double> > v = download_params(h);
    for (const ParamFile & f : param_files(h))
    {
      const std::string path = std::string(prefix)+f.suffix;
      std::ifstream reader(path);
      if (!reader.is_open())
      { // ref :247-251 -- not an error
        std::cout << "# --- file-path: " << path << " is not exist..." << std::endl;
        continue;
      }
      std::vector<std::complex<double> > raw;
      std::complex<double> temp;
      while (reader >> temp) raw.push_back(temp);
      if ((long long)raw.size() == f.count)
        std::copy(raw.begin(), raw.end(), v.begin()+f.off);
      else
        std::cout << "# check '" << f.name << "' size... " << std::endl; // ref :258-259
    }
    upload_params(h, v);
  });
}

nqs_status nqs_save_params(nqs_handle * h, const char * prefix, int32_t precision)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(prefix, NQS_ERR_INVALID, "nqs_save_params: null prefix");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    const std::vector<std::complex<double> > v = download_params(h);
    int idx = 0;
    for (const ParamFile & f : param_files(h))
    {
      std::ofstream writer(std::string(prefix)+f.suffix);
      NQS_REQUIRE(writer.is_open(), NQS_ERR_IO, std::string("cannot write ")+prefix+f.suffix);
      writer << std::setprecision(precision);
      for (long long n = 0; n < f.count; ++n)
      {
        writer << v[f.off+n] << " ";
        if (idx == 0 && (n+1)%f.row == 0) writer << std::endl; // W: one row per line
      }
      if (idx == 1) writer << std::endl;                        // a / w1o: trailing newline; b / b1: none
      ++idx;
    }
  });
}

nqs_status nqs_initialize(nqs_handle * h, const int8_t * spins)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]() { NQS_CUDA(cudaSetDevice(h->cfg.device)); do_initialize(h, spins); });
}

nqs_status nqs_warm_up(nqs_handle * h, int32_t n_sweeps, const int8_t * spins)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    do_initialize(h, spins);
    // the all-true accept_next_state_ quirk (ref impl_mcmc_sampler.cuh:21-22): flips site index_ everywhere, lnpsi0 stays stale
    const int site = h->flip_index;
    flip_site_all_kernel<<<grid_for(h->K*(long long)h->M, 256, 148*8), 256, 0, h->stream>>>(h->N, h->M, h->K, h->model, h->params.p,
      h->spins.p, h->theta.p, h->sa.p, site);
    check_launch(h, "flip_site_all_kernel");
    flip_site_all_finish_kernel<<<grid_for(h->K, 256, 148*4), 256, 0, h->stream>>>(h->N, h->M, h->K, h->model, h->params.p,
      h->spins.p, h->sa.p, site);
    check_launch(h, "flip_site_all_finish_kernel");
    NQS_CUDA(cudaMemsetAsync(h->fresh.p, 0, (size_t)h->K, h->stream)); // lnpsi0 still holds the un-flipped state's value
    do_sweeps(h, n_sweeps);
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_do_mcmc_steps(nqs_handle * h, int32_t n_sweeps)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    reset_phase_times(h);
    { Span t(h, TAG_SWEEP); do_sweeps(h, n_sweeps); }
    resolve_spans(h);
  });
}

nqs_status nqs_set_uniforms(nqs_handle * h, const double * u, int64_t steps)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    h->u_zc = nullptr;
    if (u == nullptr) { h->u_steps = 0; h->u_used = 0; return; }
    NQS_REQUIRE(steps > 0, NQS_ERR_INVALID, "nqs_set_uniforms: steps <= 0");
    // A PINNED (page-locked, device-mapped) host buffer is read in place by the sweep kernels: every uniform is used exactly
    // once, so staging 8 K N bytes in HBM first only adds a serial copy in front of the sweep (0.3 ms for 16.8 MB at N=128,
    // K=16384); the kernels fetch a group of proposals ahead, so the PCIe latency stays off the dependency chain.
    // NQS_UNIFORMS_ZEROCOPY=0 keeps the staging copy.  Either way `u` must stay valid and unchanged until the sweeps that consume
    // it have completed (the next synchronising call after them).
    const char * zc = std::getenv("NQS_UNIFORMS_ZEROCOPY");
    if (!(zc && std::atoi(zc) == 0))
    {
      cudaPointerAttributes at;
      if (cudaPointerGetAttributes(&at, u) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr)
        h->u_zc = static_cast<const double*>(at.devicePointer);
      else cudaGetLastError();
    }
    if (h->u_zc == nullptr)
    {
      NQS_REQUIRE((size_t)steps*h->K <= h->uniforms.n, NQS_ERR_INVALID, "nqs_set_uniforms: steps exceed nqs_config.max_predrawn_steps");
      NQS_CUDA(cudaMemcpyAsync(h->uniforms.p, u, sizeof(double)*(size_t)steps*h->K, cudaMemcpyHostToDevice, h->stream));
    }
    h->u_steps = steps; h->u_used = 0;
  });
}

nqs_status nqs_get_spins(nqs_handle * h, int8_t * spins)
{
  if (!h || !spins) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaMemcpyAsync(spins, h->spins.p, (size_t)h->K*h->N, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_get_lnpsi(nqs_handle * h, nqs_cdouble * lnpsi)
{
  if (!h || !lnpsi) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaMemcpyAsync(lnpsi, h->lnpsi0.p, sizeof(cd)*h->K, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_get_theta(nqs_handle * h, nqs_cdouble * theta)
{
  if (!h || !theta) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaMemcpyAsync(theta, h->theta.p, sizeof(cd)*(size_t)h->K*h->M, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_get_accept_log(nqs_handle * h, uint8_t * acc, int64_t steps)
{
  if (!h || !acc) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->cfg.flags & NQS_FLAG_ACCEPT_LOG, NQS_ERR_STATE, "handle created without NQS_FLAG_ACCEPT_LOG");
    NQS_REQUIRE(steps == h->acc_log_steps, NQS_ERR_INVALID, "nqs_get_accept_log: steps != proposals of the last sweep call");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaMemcpyAsync(acc, h->acc_log.p, (size_t)steps*h->K, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_forward_flip(nqs_handle * h, int32_t site, nqs_cdouble * lnpsi1)
{
  if (!h || !lnpsi1) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(site >= 0 && site < h->N, NQS_ERR_INVALID, "nqs_forward_flip: site out of range");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    launch_eloc(h, h->lnpsi1.p, site);
    h->flip_index = site;
    NQS_CUDA(cudaMemcpyAsync(lnpsi1, h->lnpsi1.p, sizeof(cd)*h->K, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_lnpsi_fixed_spins(nqs_handle * h, const int8_t * spins, nqs_cdouble * lnpsi)
{
  if (!h || !spins || !lnpsi) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaMemcpyAsync(h->tmp_spins.p, spins, (size_t)h->K*h->N, cudaMemcpyHostToDevice, h->stream));
    // ref forward(spins, lnpsi, false): theta from the argument, sa from the MEMBER spins (:119-120); chain state untouched here
    // (RBMTrSymm::forward(spins, ...) takes sa from its ARGUMENT, :401-403)
    launch_theta(h, h->tmp_spins.p, h->trsymm ? h->tmp_spins.p : h->spins.p, nullptr, nullptr, h->lnpsi1.p);
    NQS_CUDA(cudaMemcpyAsync(lnpsi, h->lnpsi1.p, sizeof(cd)*h->K, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_local_energy(nqs_handle * h, nqs_cdouble * htilda)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->initialized, NQS_ERR_STATE, "nqs_local_energy before nqs_initialize / nqs_warm_up");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    reset_phase_times(h);
    { Span t(h, TAG_ELOC); launch_eloc(h, nullptr, 0); }
    resolve_spans(h);
    h->flip_index = h->N-1; // the reference's loop ends with forward(L-1)
    if (htilda)
    {
      NQS_CUDA(cudaMemcpyAsync(htilda, h->htilda.p, sizeof(cd)*h->K, cudaMemcpyDeviceToHost, h->stream));
      NQS_CUDA(cudaStreamSynchronize(h->stream));
    }
  });
}

nqs_status nqs_log_derivs(nqs_handle * h, nqs_cdouble * O_host)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->aO.p != nullptr, NQS_ERR_STATE, "handle created with NQS_FLAG_NO_SR");
    NQS_REQUIRE(h->initialized, NQS_ERR_STATE, "nqs_log_derivs before nqs_initialize / nqs_warm_up");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    reset_phase_times(h);
    if (h->O.p == nullptr) h->O.alloc((size_t)h->K*(size_t)h->P);   // structured mode keeps no O until somebody asks for it
    { Span t(h, TAG_ODERIV); launch_oderiv(h); }
    resolve_spans(h);
    if (O_host)
    {
      NQS_CUDA(cudaMemcpyAsync(O_host, h->O.p, sizeof(cd)*(size_t)h->K*(size_t)h->P, cudaMemcpyDeviceToHost, h->stream));
      NQS_CUDA(cudaStreamSynchronize(h->stream));
    }
  });
}

nqs_status nqs_smatrix_dot(nqs_handle * h, double lambda, const nqs_cdouble * v, nqs_cdouble * Sv, nqs_cdouble * aO, double * diag)
{
  if (!h || !v || !Sv) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->aO.p != nullptr, NQS_ERR_STATE, "handle created with NQS_FLAG_NO_SR");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    const long long P = h->P;
    if (h->struct_sv && !h->hidden_valid) launch_hidden_values(h);
    sr_setup(h, false);
    NQS_CUDA(cudaMemcpyAsync(h->pvec.p, v, sizeof(cd)*P, cudaMemcpyHostToDevice, h->stream));
    const int nparts = matvec_passes(h, h->pvec.p, nullptr);
    launch_cg_fused(h, CG_MODE_DOT, nparts, lambda, h->pvec.p);
    NQS_CUDA(cudaMemcpyAsync(Sv, h->t.p, sizeof(cd)*P, cudaMemcpyDeviceToHost, h->stream));
    if (aO) NQS_CUDA(cudaMemcpyAsync(aO, h->aO.p, sizeof(cd)*P, cudaMemcpyDeviceToHost, h->stream));
    if (diag) NQS_CUDA(cudaMemcpyAsync(diag, h->diag.p, sizeof(double)*P, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_sr_step(nqs_handle * h, const nqs_sr_options * opt, nqs_sr_stats * st)
{
  if (!h || !opt) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->aO.p != nullptr, NQS_ERR_STATE, "handle created with NQS_FLAG_NO_SR");
    NQS_REQUIRE(h->initialized, NQS_ERR_STATE, "nqs_sr_step before nqs_warm_up");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    reset_phase_times(h);
    nqs_sr_stats s;
    std::memset(&s, 0, sizeof(s));
    { Span t(h, TAG_SWEEP); do_sweeps(h, opt->n_mc_steps); }
    // (Running the HBM-write-bound O writer on a side stream next to the fp64-bound local energy was tried twice -- one CTA per
    // chain, and a grid-stride writer of 2 CTAs per SM that leaves room for the local-energy CTAs: the local energy is L1/LSU
    // heavy and slows from 0.59 to 1.85 ms next to the writer, 14.10 vs 14.17 ms per step.  Not kept.)
    { Span t(h, TAG_ELOC); launch_eloc(h, nullptr, 0); h->flip_index = h->N-1; }
    { Span t(h, TAG_ODERIV);
      if (h->struct_sv) launch_hidden_values(h);
      else if (h->gen_ok) { launch_hidden_values(h); h->o_pending = true; }   // O is written by the first S*v of the CG
      else launch_oderiv(h); }
    { Span t(h, TAG_SETUP); sr_setup(h, true); }
    const double bp_before = h->bp;
    if (opt->lambda < 0)
    { // ref schedular_, impl_optimizer.cuh:72-78
      h->bp *= 0.9;
      const double lam = 100.0*h->bp;
      s.lambda = (lam > 1e-2) ? lam : 1e-2;
    }
    else s.lambda = opt->lambda;
    double hs[3];
    const double invk = 1.0/(double)h->Ktot;
    const bool persistent = cgp_usable(h);
    if (persistent)
    { // nothing below waits for the device: the energy check (ref optimizer.cuh:134-138) is made by the CG kernel itself, which
      // skips the solve -- and, through its flag, the update -- when <h> is not finite; one synchronisation ends the step
      { Span t(h, TAG_CG); cg_solve_persistent(h, s.lambda, opt->tol, opt->max_iter, opt->fixed_iters); }
      if (opt->apply_update)
      { Span t(h, TAG_UPDATE); do_evolve(h, h->dx.p, opt->lr, h->scal.p, 0); }
      NQS_CUDA(cudaMemcpyAsync((char*)h->pinned+PIN_HS, h->hsall.p, sizeof(double)*3, cudaMemcpyDeviceToHost, h->stream));
      NQS_CUDA(cudaStreamSynchronize(h->stream)); // ref cudaDeviceSynchronize, optimizer.cuh:153
      finish_tables(h);
      std::memcpy(hs, (char*)h->pinned+PIN_HS, sizeof(hs));
      resolve_spans(h);
      bool nonfinite = false;
      collect_cg(h, &s, &nonfinite);
      s.e_re = hs[0]*invk; s.e_im = hs[1]*invk;
      s.finite = nonfinite ? 0 : 1;
      if (nonfinite) { h->bp = bp_before; if (st) *st = s; return; }
    }
    else
    { // launch per iteration (two-pass / structured / O-generating S*v, or NQS_CG_PERSIST=0): the same asynchronous shape -- the
      // first batch of iterations and the update are enqueued without looking at the device; one synchronisation ends the step
      int enq = 0;
      { Span t(h, TAG_CG); enq = cg_solve_async_begin(h, s.lambda, opt->tol, opt->max_iter, opt->fixed_iters); }
      if (opt->apply_update)
      { Span t(h, TAG_UPDATE); do_evolve(h, h->dx.p, opt->lr, h->scal.p, opt->fixed_iters > 0 ? 0 : 1); }
      NQS_CUDA(cudaMemcpyAsync((char*)h->pinned+PIN_HS, h->hsall.p, sizeof(double)*3, cudaMemcpyDeviceToHost, h->stream));
      NQS_CUDA(cudaStreamSynchronize(h->stream)); // ref cudaDeviceSynchronize, optimizer.cuh:153
      finish_tables(h);
      std::memcpy(hs, (char*)h->pinned+PIN_HS, sizeof(hs));
      bool nonfinite = false;
      const bool more = cg_finish_async(h, s.lambda, opt->max_iter, opt->fixed_iters, enq, &s, &nonfinite);
      s.e_re = hs[0]*invk; s.e_im = hs[1]*invk;
      s.finite = nonfinite ? 0 : 1;
      if (nonfinite) { h->bp = bp_before; resolve_spans(h); if (st) *st = s; return; }
      if (more && opt->apply_update)
      { // the update enqueued behind the first batch saw an unconverged solve and declined: now it applies
        { Span t(h, TAG_UPDATE); do_evolve(h, h->dx.p, opt->lr); }
        NQS_CUDA(cudaStreamSynchronize(h->stream));
        finish_tables(h);
      }
      resolve_spans(h);
    }
    const double n2 = s.e_re*s.e_re+s.e_im*s.e_im;
    s.rsd = std::sqrt((hs[2]*invk-n2)/n2);
    if (st) *st = s;
  });
}

nqs_status nqs_get_sr_vectors(nqs_handle * h, nqs_cdouble * F, nqs_cdouble * dx)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->aO.p != nullptr, NQS_ERR_STATE, "handle created with NQS_FLAG_NO_SR");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    if (F) NQS_CUDA(cudaMemcpyAsync(F, h->F.p, sizeof(cd)*h->P, cudaMemcpyDeviceToHost, h->stream));
    if (dx) NQS_CUDA(cudaMemcpyAsync(dx, h->dx.p, sizeof(cd)*h->P, cudaMemcpyDeviceToHost, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_evolve(nqs_handle * h, const nqs_cdouble * dx, double lr)
{
  if (!h || !dx) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    cd * buf = h->pvec.p;
    DevBuf<cd> tmp;
    if (!buf) { tmp.alloc(h->P); buf = tmp.p; }
    NQS_CUDA(cudaMemcpyAsync(buf, dx, sizeof(cd)*h->P, cudaMemcpyHostToDevice, h->stream));
    do_evolve(h, buf, lr);
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    finish_tables(h);
  });
}

nqs_status nqs_enable_sr(nqs_handle * h)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]() { NQS_CUDA(cudaSetDevice(h->cfg.device)); alloc_sr(h); NQS_CUDA(cudaStreamSynchronize(h->stream)); });
}

nqs_status nqs_set_hamiltonian(nqs_handle * h, double hfield, double J, double alpha, int32_t pbc, int32_t order)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(!(pbc && h->N%2 == 1), NQS_ERR_INVALID, "kL%2 == 1 (set \"isPBC\" to \"false\".)"); // ref impl_hamiltonians.cuh:141-142
    NQS_REQUIRE(order == NQS_ORDER_CHECKERBOARD || order == NQS_ORDER_SEQUENTIAL, NQS_ERR_INVALID, "nqs_set_hamiltonian: unknown order");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    h->cfg.h = hfield; h->cfg.J = J; h->cfg.alpha = alpha; h->cfg.pbc = pbc; h->cfg.order = order;
    build_order(h);
    build_J(h);
    h->pos = 0;
  });
}

nqs_status nqs_sr_reset(nqs_handle * h)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->aO.p != nullptr, NQS_ERR_STATE, "nqs_sr_reset before nqs_enable_sr");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    h->bp = 1.0;
    NQS_CUDA(cudaMemsetAsync(h->dx.p, 0, sizeof(cd)*h->P, h->stream));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

nqs_status nqs_set_seed(nqs_handle * h, uint64_t seed)
{
  if (!h) return NQS_ERR_INVALID;
  h->cfg.seed = seed;
  h->step_counter = 0;
  h->yarn_state_at = -1;
  return NQS_OK;
}

nqs_status nqs_set_rng(nqs_handle * h, int32_t kind, uint64_t seed, uint64_t seed_distance)
{
  if (!h) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(kind == NQS_RNG_PHILOX || kind == NQS_RNG_YARN2, NQS_ERR_INVALID, "nqs_set_rng: unknown generator kind");
    h->rng_kind = kind;
    h->cfg.seed = seed;
    h->seed_distance = seed_distance;
    h->step_counter = 0;
    h->yarn_state_at = -1;
  });
}

nqs_status nqs_comm_get_unique_id(char id[NQS_UNIQUE_ID_BYTES])
{
  if (!id) return NQS_ERR_INVALID;
  if (!g_nccl.load()) { g_create_error = g_nccl.why; return NQS_ERR_NCCL; }
  ncclUniqueIdT u;
  const int rc = g_nccl.getUniqueId(&u);
  if (rc != 0) { g_create_error = "ncclGetUniqueId failed"; return NQS_ERR_NCCL; }
  std::memcpy(id, u.internal, NQS_UNIQUE_ID_BYTES);
  return NQS_OK;
}

nqs_status nqs_comm_init(nqs_handle * h, int32_t n_ranks, int32_t rank, const char id[NQS_UNIQUE_ID_BYTES])
{
  if (!h || !id) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, NQS_ERR_INVALID, "nqs_comm_init: bad rank");
    if (!g_nccl.load()) throw Error(NQS_ERR_NCCL, g_nccl.why);
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    ncclUniqueIdT u;
    std::memcpy(u.internal, id, NQS_UNIQUE_ID_BYTES);
    void * comm = nullptr;
    const int rc = g_nccl.commInitRank(&comm, n_ranks, u, rank);
    if (rc != 0)
      throw Error(NQS_ERR_NCCL, std::string("ncclCommInitRank failed: ")+(g_nccl.getErrorString ? g_nccl.getErrorString(rc) : "?"));
    h->comm = comm; h->n_ranks = n_ranks; h->rank = rank;
  });
}

nqs_status nqs_comm_p2p_export(nqs_handle * h, char handle_out[NQS_IPC_HANDLE_BYTES])
{
  if (!h || !handle_out) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->n_ranks > 1 && h->n_ranks <= NQS_CG_MAX_RANKS, NQS_ERR_STATE, "nqs_comm_p2p_export: call nqs_comm_init first (2..16 ranks)");
    NQS_REQUIRE(h->aO.p != nullptr, NQS_ERR_STATE, "handle created with NQS_FLAG_NO_SR");
    static_assert(sizeof(cudaIpcMemHandle_t) == NQS_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    if (h->xbuf == nullptr)
    {
      h->xbuf_data_bytes = (size_t)2*h->n_ranks*2*(size_t)h->P*sizeof(double);
      const size_t flag_bytes = (size_t)2*NQS_CG_MAX_RANKS*NQS_CGP_MAX_CTAS*sizeof(unsigned int);
      // behind the CG region: the SR-setup exchange (setup_exchange_finalize_kernel): data [2][n_ranks][5P+4] doubles + flags
      h->xbuf_setup_off = (h->xbuf_data_bytes+flag_bytes+64+255)/256*256;
      h->xbuf_setup_flag_off = h->xbuf_setup_off+(size_t)2*h->n_ranks*(5*(size_t)h->P+4)*sizeof(double);
      // behind that: the LL packet regions of cg_persist_kernel (16-byte packets; zero = no packet yet)
      const size_t slice = ((size_t)h->P+h->n_ranks-1)/h->n_ranks;
      h->xbuf_ll1_off = (h->xbuf_setup_flag_off+(size_t)2*NQS_CG_MAX_RANKS*NQS_CG_MAX_CTAS*sizeof(unsigned int)+255)/256*256;
      h->xbuf_ll2_off = h->xbuf_ll1_off+(size_t)h->n_ranks*slice*2*16;
      const size_t total = h->xbuf_ll2_off+(size_t)h->P*2*16;
      cudaError_t e = cudaMalloc(&h->xbuf, total);
      if (e != cudaSuccess) throw Error(NQS_ERR_NOMEM, std::string("cudaMalloc of the peer exchange buffer failed: ")+cudaGetErrorString(e));
      NQS_CUDA(cudaMemset(h->xbuf, 0, total));
      // header behind the flags: the grid this rank would run the persistent CG kernel on; the importers compare them
      const int hdr[4] = {0x4e515331, h->cgp_ok ? 1 : 0, h->sv_nclusters*h->sv_cs, h->sv_nt};
      NQS_CUDA(cudaMemcpy((char*)h->xbuf+h->xbuf_data_bytes+flag_bytes, hdr, sizeof(hdr), cudaMemcpyHostToDevice));
      NQS_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t ipc;
    NQS_CUDA(cudaIpcGetMemHandle(&ipc, h->xbuf));
    std::memcpy(handle_out, &ipc, NQS_IPC_HANDLE_BYTES);
  });
}

nqs_status nqs_comm_p2p_import(nqs_handle * h, const char * handles)
{
  if (!h || !handles) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->xbuf != nullptr, NQS_ERR_STATE, "nqs_comm_p2p_import before nqs_comm_p2p_export");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    for (int r = 0; r < h->n_ranks; ++r)
    {
      if (r == h->rank) { h->peer_base[r] = h->xbuf; continue; }
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, handles+(size_t)r*NQS_IPC_HANDLE_BYTES, NQS_IPC_HANDLE_BYTES);
      void * ptr = nullptr;
      cudaError_t e = cudaIpcOpenMemHandle(&ptr, ipc, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
      {
        cudaGetLastError();
        throw Error(NQS_ERR_UNSUPPORTED, std::string("cudaIpcOpenMemHandle(rank ")+std::to_string(r)+") failed: "+cudaGetErrorString(e)+
          " -- the engine keeps using ncclAllReduce");
      }
      h->peer_base[r] = ptr;
    }
    // every rank must run the persistent CG kernel (its packet exchange is a protocol of its own) or none
    const size_t flag_bytes = (size_t)2*NQS_CG_MAX_RANKS*NQS_CGP_MAX_CTAS*sizeof(unsigned int);
    h->cgp_peers_agree = true;
    for (int r = 0; r < h->n_ranks; ++r)
    {
      int hdr[4] = {0, 0, 0, 0};
      NQS_CUDA(cudaMemcpy(hdr, (char*)h->peer_base[r]+h->xbuf_data_bytes+flag_bytes, sizeof(hdr), cudaMemcpyDeviceToHost));
      if (hdr[0] != 0x4e515331 || hdr[1] != 1 || !h->cgp_ok) h->cgp_peers_agree = false;   // (the LL exchange pairs no CTAs: grids may differ)
    }
    h->p2p_ok = true;
    h->p2p_epoch = 0;
  });
}

nqs_status nqs_comm_p2p_disable(nqs_handle * h)
{
  if (!h) return NQS_ERR_INVALID;
  h->p2p_ok = false;
  return NQS_OK;
}

nqs_status nqs_get_timing(nqs_handle * h, nqs_timing * t)
{
  if (!h || !t) return NQS_ERR_INVALID;
  *t = h->timing;
  return NQS_OK;
}

nqs_status nqs_set_timing(nqs_handle * h, int32_t enabled)
{
  if (!h) return NQS_ERR_INVALID;
  h->timing_on = enabled != 0;
  return NQS_OK;
}

nqs_status nqs_event_record(nqs_handle * h, int32_t slot)
{
  if (!h || slot < 0 || slot >= 8) return NQS_ERR_INVALID;
  return guarded(h, [&]() { NQS_CUDA(cudaSetDevice(h->cfg.device)); NQS_CUDA(cudaEventRecord(h->ev[slot], h->stream)); });
}

nqs_status nqs_event_elapsed_ms(nqs_handle * h, int32_t a, int32_t b, float * ms)
{
  if (!h || !ms || a < 0 || a >= 8 || b < 0 || b >= 8) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaEventSynchronize(h->ev[b]));
    NQS_CUDA(cudaEventElapsedTime(ms, h->ev[a], h->ev[b]));
  });
}

extern "C++"
{
namespace
{
struct CkptHeader
{
  char magic[8];            // "NQSCKPT2"
  int32_t model, N, M, trsymm;
  int64_t K, Ktot, koff, P;
  int32_t pos, flip_index, cg_prev_iters, has_sr;
  uint64_t step_counter, seed;
  double bp, hfield, J, alpha;
  int32_t pbc, order;
  int32_t rng_kind, pad_;
  uint64_t seed_distance;
};
template <typename T>
void ckpt_write(std::ofstream & f, nqs_handle * h, const T * dev, size_t n)
{
  std::vector<T> buf(n);
  NQS_CUDA(cudaMemcpyAsync(buf.data(), dev, n*sizeof(T), cudaMemcpyDeviceToHost, h->stream));
  NQS_CUDA(cudaStreamSynchronize(h->stream));
  f.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)(n*sizeof(T)));
}
template <typename T>
void ckpt_read(std::ifstream & f, nqs_handle * h, T * dev, size_t n)
{
  std::vector<T> buf(n);
  f.read(reinterpret_cast<char*>(buf.data()), (std::streamsize)(n*sizeof(T)));
  NQS_REQUIRE((size_t)f.gcount() == n*sizeof(T), NQS_ERR_IO, "checkpoint file truncated");
  NQS_CUDA(cudaMemcpyAsync(dev, buf.data(), n*sizeof(T), cudaMemcpyHostToDevice, h->stream));
  NQS_CUDA(cudaStreamSynchronize(h->stream));
}
} // namespace
} // extern "C++"

nqs_status nqs_checkpoint_save(nqs_handle * h, const char * path)
{
  if (!h || !path) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_REQUIRE(h->initialized, NQS_ERR_STATE, "nqs_checkpoint_save before nqs_initialize / nqs_warm_up");
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    finish_tables(h);
    std::ofstream f(path, std::ios::binary);
    NQS_REQUIRE(f.is_open(), NQS_ERR_IO, std::string("cannot write ")+path);
    CkptHeader hd;
    std::memset(&hd, 0, sizeof(hd));
    std::memcpy(hd.magic, "NQSCKPT2", 8);
    hd.model = h->model; hd.N = h->N; hd.M = h->M; hd.trsymm = h->tied;
    hd.K = h->K; hd.Ktot = h->Ktot; hd.koff = h->koff; hd.P = h->P;
    hd.pos = h->pos; hd.flip_index = h->flip_index; hd.cg_prev_iters = h->cg_prev_iters; hd.has_sr = h->aO.p != nullptr ? 1 : 0;
    hd.step_counter = h->step_counter; hd.seed = h->cfg.seed; hd.bp = h->bp;
    hd.hfield = h->cfg.h; hd.J = h->cfg.J; hd.alpha = h->cfg.alpha; hd.pbc = h->cfg.pbc; hd.order = h->cfg.order;
    hd.rng_kind = h->rng_kind; hd.seed_distance = h->seed_distance;
    f.write(reinterpret_cast<const char*>(&hd), sizeof(hd));
    ckpt_write(f, h, var_ptr(h), (size_t)h->P);
    ckpt_write(f, h, h->spins.p, (size_t)h->K*h->N);
    ckpt_write(f, h, h->theta.p, (size_t)h->K*h->M);
    ckpt_write(f, h, h->lnpsi0.p, (size_t)h->K);
    ckpt_write(f, h, h->sa.p, (size_t)h->K);
    ckpt_write(f, h, h->fresh.p, (size_t)h->K);
    if (hd.has_sr) ckpt_write(f, h, h->dx.p, (size_t)h->P);
    f.close();
    NQS_REQUIRE(f.good(), NQS_ERR_IO, std::string("write error on ")+path);
  });
}

nqs_status nqs_checkpoint_load(nqs_handle * h, const char * path)
{
  if (!h || !path) return NQS_ERR_INVALID;
  return guarded(h, [&]()
  {
    NQS_CUDA(cudaSetDevice(h->cfg.device));
    std::ifstream f(path, std::ios::binary);
    NQS_REQUIRE(f.is_open(), NQS_ERR_IO, std::string("cannot read ")+path);
    CkptHeader hd;
    f.read(reinterpret_cast<char*>(&hd), sizeof(hd));
    NQS_REQUIRE((size_t)f.gcount() == sizeof(hd) && std::memcmp(hd.magic, "NQSCKPT2", 8) == 0, NQS_ERR_IO, "not a libnqs_b200 checkpoint");
    NQS_REQUIRE(hd.model == h->model && hd.N == h->N && hd.M == h->M && hd.trsymm == h->tied && hd.K == h->K &&
      hd.Ktot == h->Ktot && hd.koff == h->koff && hd.P == h->P, NQS_ERR_INVALID,
      "checkpoint was written by a handle of another shape (model, sizes, chains or chain offset differ)");
    NQS_CUDA(cudaStreamSynchronize(h->stream));
    invalidate_tables(h);
    h->theta_matches_O = false; h->hidden_valid = false; h->o_pending = false;
    ckpt_read(f, h, var_ptr(h), (size_t)h->P);
    expand_vars(h);
    ckpt_read(f, h, h->spins.p, (size_t)h->K*h->N);
    ckpt_read(f, h, h->theta.p, (size_t)h->K*h->M);
    ckpt_read(f, h, h->lnpsi0.p, (size_t)h->K);
    ckpt_read(f, h, h->sa.p, (size_t)h->K);
    ckpt_read(f, h, h->fresh.p, (size_t)h->K);
    if (hd.has_sr)
    {
      alloc_sr(h);
      ckpt_read(f, h, h->dx.p, (size_t)h->P);
    }
    h->pos = hd.pos; h->flip_index = hd.flip_index; h->cg_prev_iters = hd.cg_prev_iters;
    h->step_counter = hd.step_counter; h->cfg.seed = hd.seed; h->bp = hd.bp;
    h->rng_kind = hd.rng_kind; h->seed_distance = hd.seed_distance; h->yarn_state_at = -1;   // the yarn2 state is rebuilt from the counters
    if (hd.hfield != h->cfg.h || hd.J != h->cfg.J || hd.alpha != h->cfg.alpha || hd.pbc != h->cfg.pbc || hd.order != h->cfg.order)
    { // the Hamiltonian travels with the state
      h->cfg.h = hd.hfield; h->cfg.J = hd.J; h->cfg.alpha = hd.alpha; h->cfg.pbc = hd.pbc; h->cfg.order = hd.order;
      build_order(h);
      build_J(h);
    }
    h->u_steps = 0; h->u_used = 0;
    h->initialized = true;
    NQS_CUDA(cudaStreamSynchronize(h->stream));
  });
}

const char * nqs_kernel_variant(const nqs_handle * h, const char * stage)
{
  if (!h || !stage) return "";
  if (!std::strcmp(stage, "sweep")) return h->variant_sweep.c_str();
  if (!std::strcmp(stage, "eloc")) return h->variant_eloc.c_str();
  if (!std::strcmp(stage, "theta")) return h->variant_theta.c_str();
  if (!std::strcmp(stage, "sv")) return h->variant_sv.c_str();
  return "";
}
} // extern "C"
