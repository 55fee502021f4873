// Host-side engine behind the C ABI of include/nqs_b200.h.  One nqs_handle owns the chains of one GPU: parameters,
// chain state (spins, theta, lnpsi0, sa), the O matrix and the CG vectors, all resident in HBM for the life of the handle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/nqs_b200.h"
#include "device_math.cuh"
#include "fast_kernels.cuh"
#include "ffnn_fast_kernels.cuh"
#include "sweep_f32.cuh"
#include "rows_umma.cuh"
#include "cols_umma.cuh"
#include "yarn2.cuh"

namespace nqs
{
struct Error: public std::runtime_error
{
  nqs_status code;
  Error(nqs_status c, const std::string & m): std::runtime_error(m), code(c) {}
};

#define NQS_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  throw nqs::Error(NQS_ERR_CUDA, std::string(#call)+": "+cudaGetErrorString(e_)+" ("+__FILE__+":"+std::to_string(__LINE__)+")"); } while (0)
#define NQS_REQUIRE(cond, code, msg) do { if (!(cond)) throw nqs::Error(code, msg); } while (0)

struct CgScalars;
// tied-variable (symmetric) ansaetze: sampled as their expanded plain network, optimised in the tied variables
enum { TIED_NONE = 0, TIED_RBM_TR = 1, TIED_RBM_Z2PR = 2, TIED_FFNN_TR = 3 };

template <typename T>
struct DevBuf
{
  T * p = nullptr;
  size_t n = 0;
  void alloc(size_t count)
  {
    free();
    if (count == 0) return;
    cudaError_t e = cudaMalloc(&p, count*sizeof(T));
    if (e != cudaSuccess)
      throw Error(NQS_ERR_NOMEM, "cudaMalloc of "+std::to_string(count*sizeof(T))+" bytes failed: "+cudaGetErrorString(e));
    n = count;
  }
  void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
  ~DevBuf() { free(); }
};
} // namespace nqs

struct nqs_handle
{
  nqs_config cfg;
  int N = 0, M = 0, model = 0;
  long long K = 0, Ktot = 0, koff = 0, P = 0;
  // translation-symmetric RBM (NQS_MODEL_RBMTRSYMM): `model` is MODEL_RBM for every sampler kernel, `params` holds the EXPANDED
  // network [wf | af | bf] (Pfull entries), `vars` the P = N*alpha+1+alpha variables the optimiser moves
  bool trsymm = false;                    // any tied-variable ansatz (the name is the first one built)
  int tied = 0;                           // TIED_*: which one
  int alpha_f = 0;
  long long Pfull = 0;
  nqs::DevBuf<nqs::cd> vars;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  size_t smem_optin = 0;

  // model + chain state
  nqs::DevBuf<nqs::cd> params, theta, lnpsi0, lnpsi1, sa, htilda, tmp_theta;
  nqs::DevBuf<int8_t> spins, tmp_spins;
  nqs::DevBuf<double> Jmat, uniforms, sjs;
  nqs::DevBuf<int> order;
  std::vector<int> order_host;            // host copy of the site ring (the machine's index_ after a sweep is read from it)
  nqs::DevBuf<unsigned char> acc_log, fresh;
  // specialised RBM path: flip tables rebuilt after every parameter change (fast_kernels.cuh)
  nqs::DevBuf<nqs::FlipTab> ftab_a, ftab_b;
  nqs::DevBuf<float4> ftab32;             // fp32 copy of the flip tables (sweep_f32.cuh)
  nqs::DevBuf<unsigned long long> f32_stats;   // [2] proposals / proposals decided by the fp64 path
  nqs::DevBuf<nqs::CoshTab> ctab_a, ctab_b, ctabT_a, ctabT_b;
  nqs::DevBuf<nqs::cd> w2, aexp;
  nqs::DevBuf<double> afac, bound;
  bool tables_valid = false;
  double theta_bound = 0.0;
  int jpl = 0, mpad = 0, npad32 = 0;
  long long u_steps = 0, u_used = 0;      // pre-drawn feed: proposals available / consumed
  const double * u_zc = nullptr;          // the feed is read IN PLACE from the caller's pinned host buffer (device alias), else null
  long long acc_log_steps = 0;
  int pos = 0;                            // next position in the site ring
  int flip_index = 0;                     // the machine's index_ (ref impl_neural_quantum_state.cuh:19)
  unsigned long long step_counter = 0;    // proposals done so far (RNG counter)
  // trng::yarn2 stream (yarn2.cuh, nqs_set_rng): counter-addressed by (seed, seed_distance, global chain, step_counter)
  int rng_kind = 0;                       // nqs_rng
  unsigned long long seed_distance = 0;
  nqs::DevBuf<uint32_t> yarn_tab;         // g^i tables of the output map
  nqs::DevBuf<uint2> yarn_state;          // [K] engine state after `yarn_state_at` draws
  long long yarn_state_at = -1;           // -1: rebuild from the counters at the next fill
  nqs::DevBuf<double> yarn_u;             // [steps][K] uniforms of the sweep launch in flight
  bool initialized = false;
  bool theta_matches_O = false;           // O was written from the current spins / theta / params (structured SR setup allowed)

  // SR
  nqs::DevBuf<nqs::cd> O, aO, F, dx, r, pvec, z, t, zk;
  nqs::DevBuf<double> diag, part, sums, traw, slots, hsall;
  nqs::DevBuf<nqs::CgScalars> scal;
  nqs::DevBuf<unsigned int> cgbar;         // grid-barrier counter of cg_fused_kernel (zero between launches)
  double bp = 1.0;                        // lambda schedule state (ref bp_, optimizer.cuh:176)
  int cg_prev_iters = 4;                  // iterations the previous CG solve needed (how many to enqueue before polling)
  int nrb = 1;                            // row blocks of the column passes
  long long rows_per_block = 0;
  void * pinned = nullptr;                // small pinned staging area for scalar read-backs
  bool cg_attr_set = false;               // cg_fused_kernel's carve-out preference was set on this handle's device
  int cg_coop = -1;                       // -1 unknown, 1 = cooperative launches work on this device, 0 = plain launches + barrier time-out
  // one-pass S*v (sv_fused.cuh): cluster size, columns per thread, threads, TMA slots, clusters, rows per cluster
  bool sv_ok = false;
  int sv_cs = 0, sv_cpt = 0, sv_nt = 0, sv_nslot = 0, sv_nclusters = 0, sv_defer = 0, sv_depth = 0;
  size_t sv_smem = 0, sv_slot_bytes = 0;
  long long sv_pc = 0, sv_rpc = 0;
  // persistent CG (cg_persist.cuh): the whole solve as one launch of the sv_fused clusters
  bool cgp_ok = false;                    // planned for this handle (single GPU, or every rank planned the same grid)
  bool cgp_peers_agree = true;            // multi-GPU: all ranks export the same persistent grid (checked at p2p import)
  int cgp_coop = -1;                      // cooperative + cluster launch: -1 untried, 1 works, 0 refused (plain cluster launch, barrier time-out)
  bool cg_check_finite = false;           // the INIT launch of the launch-per-iteration solve checks <h> itself (nqs_sr_step)
  int cg_async_enq = 0;                   // launch-per-iteration solve enqueued without polling: iterations in the queue (0 = none pending)
  bool cg_inflight = false;               // a persistent solve was launched and its scalars are not read back yet
  bool bound_inflight = false;            // theta_bound of freshly built tables is on its way to pinned memory
  // structured S*v (sv_struct.cuh, NQS_FLAG_STRUCTURED_SV): hidden-unit factors T (and L, FFNN), chain chunks of the column GEMM
  bool struct_sv = false, hidden_valid = false;
  nqs::DevBuf<nqs::cd> Tm, Lm, vnat;
  nqs::DevBuf<int8_t> bq;                 // rows_umma.cuh: int8 digit planes of the B operand (W, or the W block of v), UMMA tile order
  nqs::DevBuf<double> bscale;             // their per-column power-of-two scales [2M]
  nqs::DevBuf<unsigned long long> tmaxb;  // cols_umma.cuh: bit patterns of max_k |T_kj|_inf per hidden unit [M]
  int cols_umma = 0;                      // 1: the O^H z / SR-setup GEMM runs on tcgen05 (int8 UMMA), 0: fp64 DMMA
  int rows_umma = -1;                     // -1 undecided, 0 fp64 DMMA rows kernel, 1 tcgen05 int8 (Ozaki) rows kernel
  int gen_cs = 0, gen_cpt = 0, gen_nt = 0, gen_nclusters = 0, gen_q = 0;   // launch geometry of the O-generating S*v (sv_fused.cuh, GEN)
  long long gen_pc = 0, gen_rpc = 0;
  bool gen_ok = false, o_pending = false; // O is written by the first S*v of the CG instead of a separate writer / it still has to be
  nqs::DevBuf<double> abs2;               // [chunks][3M] sums of |T|^2, |L|^2 of the SR setup GEMM
  bool cols_ok = false;                   // spin_cols_dmma_kernel planned (N <= 256): structured S*v and the SR setup GEMM
  int sc_variant = 0, sc_nchunks = 0, sc_colgroups = 0;
  long long sc_rows_per_chunk = 0;

  // multi-GPU
  void * comm = nullptr;                  // ncclComm_t
  int n_ranks = 1, rank = 0;
  // in-kernel all-reduce of the CG partials over NVLink peer memory (cg_fused.cuh): receive buffer [2][n_ranks][2P] doubles +
  // flags [2][16] on every rank, mapped into every peer with cudaIpc
  void * xbuf = nullptr;
  size_t xbuf_data_bytes = 0;
  void * peer_base[16] = {nullptr};
  bool p2p_ok = false;
  unsigned int p2p_epoch = 0;
  size_t xbuf_ll1_off = 0, xbuf_ll2_off = 0;             // LL packet regions of the persistent CG kernel's reduce-scatter / all-gather
  size_t xbuf_setup_off = 0, xbuf_setup_flag_off = 0;   // SR-setup exchange region of xbuf (data [2][n_ranks][5P+4], flags)
  unsigned int setup_epoch = 0;
  nqs::DevBuf<unsigned long long> cg_trace;   // NQS_CG_TRACE=1: per-launch time stamps of cg_fused_kernel (diagnostics)
  long long cg_trace_n = 0;

  // bookkeeping
  std::string err;
  bool timing_on = false;
  nqs_timing timing;
  cudaEvent_t ev[8];                      // user slots of nqs_event_record
  bool ev_ok = false;
  struct Span { int tag, b, e; };
  std::vector<cudaEvent_t> evpool;        // events of the timed spans, resolved after the call's final sync
  size_t ev_used = 0;
  std::vector<Span> spans;
  std::string variant_sweep = "generic", variant_eloc = "generic", variant_theta = "generic", variant_sv = "two_pass";
};
