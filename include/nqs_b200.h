/*
 * nqs_b200.h -- C ABI of libnqs_b200.so, the B200-native (sm_100a) variational-Monte-Carlo engine for
 * neural-network quantum states.  It is the drop-in boundary for ONE hot path of
 * dkkim1005/Neural_Network_Quantum_State: Metropolis single-spin-flip sweep -> long-range TFI local energy ->
 * log-derivatives O -> stochastic-reconfiguration conjugate gradient -> parameter update, for the complex RBM and
 * the one-hidden-layer complex FNN.  "ref:" comments cite the reference interface each entry point replaces
 * (paths relative to the reference root; all reference arithmetic on this path is thrust::complex<double>).
 *
 * Conventions
 *   - plain C, no C++/torch types; complex numbers are interleaved {re, im} doubles (== std::complex<double>,
 *     thrust::complex<double>, numpy complex128, cuDoubleComplex);
 *   - every pointer argument is a HOST pointer unless its name ends in _dev; the library owns all device memory;
 *   - one handle == the chains of ONE GPU (one process per GPU in multi-GPU runs); a handle is not thread-safe,
 *     like the reference objects (single host thread, default stream);
 *   - every call returns nqs_status and never exits the process (the reference calls exit(1), gpu/include/common.cuh:12-17);
 *     nqs_last_error() gives the message;
 *   - there is NO CPU fallback: without a CUDA device nqs_create fails with NQS_ERR_CUDA.
 *   - index notation as in the reference: i = visible site (N), j = hidden unit (M), k = chain (K), P = #variables.
 */
#ifndef NQS_B200_H
#define NQS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NQS_B200_ABI_VERSION 1

typedef struct nqs_handle nqs_handle;

typedef struct nqs_cdouble { double re, im; } nqs_cdouble;

typedef enum nqs_status {
  NQS_OK = 0,
  NQS_ERR_INVALID = 1,    /* bad argument / size mismatch (ref: std::length_error, std::invalid_argument) */
  NQS_ERR_CUDA = 2,       /* CUDA runtime failure or no device (ref: CHECK_ERROR -> exit(1)) */
  NQS_ERR_NOMEM = 3,
  NQS_ERR_IO = 4,
  NQS_ERR_STATE = 5,      /* call order violated (e.g. sweep before initialize) */
  NQS_ERR_NCCL = 6,
  NQS_ERR_NONFINITE = 7,  /* <H> is not finite (ref: gpu/include/optimizer.cuh:134-138 prints and returns) */
  NQS_ERR_UNSUPPORTED = 8
} nqs_status;

/* NQS_MODEL_RBMTRSYMM: the translation-symmetric RBM (ref: RBMTrSymm<T>(nInputs, alpha, nChains), gpu/include/
 * neural_quantum_state.cuh:62-105, impl :301-538, kernels :1487-1553; driver gpu/src/LICH-train_rbmtrsymm.cu).  Variables are
 * [w (f*N+i, alpha filters) | a (1) | b (alpha)], P = N*alpha + 1 + alpha; the sampler works on the expanded RBM
 * wf[i][f*N+j] = w[f][(i+j)%N], bf[f*N+j] = b[f], af[i] = a[0].  nqs_config.n_hiddens is the EXPANDED width alpha*N
 * (a multiple of n_inputs); parameter files: ONE file `prefix` holding all P variables (ref :474-482). */
/* NQS_MODEL_RBMZ2PRSYMM -- ref RBMZ2PrSymm<T> (:106-147, impl :540-745, kernels :1556-1618): Z2- and parity-symmetric RBM,
 * variables [w (i*alpha+f) | b (alpha)], P = N*alpha + alpha, no visible bias; expanded RBM of n_hiddens = 4*alpha units
 * wf[i][4f+{0,1,2,3}] = {w[i][f], -w[i][f], w[N-1-i][f], -w[N-1-i][f]}, bf[4f+j] = b[f].
 * NQS_MODEL_FFNNTRSYMM -- ref FFNNTrSymm<T> (:197-237, impl :1019-1223, kernels :1693-1750): translation-symmetric FNN, variables
 * [wi1 (f*N+i) | b1 (alpha) | w1o (alpha)], P = N*alpha + 2*alpha; expanded FNN of n_hiddens = alpha*N units
 * W1[i][f*N+j] = wi1[f][(i+j)%N], b1f[f*N+j] = b1[f], w1of[f*N+j] = w1o[f].  Both: ONE parameter file named by the prefix. */
typedef enum nqs_model { NQS_MODEL_RBM = 0, NQS_MODEL_FFNN = 1, NQS_MODEL_RBMTRSYMM = 2, NQS_MODEL_RBMZ2PRSYMM = 3,
                         NQS_MODEL_FFNNTRSYMM = 4 } nqs_model;
/* site visiting order of one sweep */
typedef enum nqs_order {
  NQS_ORDER_CHECKERBOARD = 0, /* 2,4,..,1,3,..,0  ref: gpu/include/impl_hamiltonians.cuh:163-180,209-210 (LITFIChain)   */
  NQS_ORDER_SEQUENTIAL = 1    /* 1,2,..,N-1,0     ref: gpu/include/impl_meas.cuh:12-21,33-34 (Sampler4SpinHalf, pynqs) */
} nqs_order;

/* ref: RBM<T>(nInputs,nHiddens,nChains) gpu/include/neural_quantum_state.cuh:21 + LITFIChain(machine,L,h,J,alpha,isPBC,
 * seedNumber,seedDistance,prefix) gpu/include/hamiltonians.cuh:52-55 + StochasticReconfigurationCG<T>(K,P)
 * gpu/include/optimizer.cuh:117. */
typedef struct nqs_config {
  int32_t abi_version;     /* NQS_B200_ABI_VERSION */
  int32_t model;           /* nqs_model */
  int32_t n_inputs;        /* N = L */
  int32_t n_hiddens;       /* M */
  int64_t n_chains;        /* K_loc: chains owned by this handle (this GPU) */
  int64_t n_chains_total;  /* K over all ranks (0 -> n_chains) */
  int64_t chain_offset;    /* global id of local chain 0; keys the internal RNG so results do not depend on the #GPUs */
  double h;                /* transverse field  (ref driver: h = -cos(theta), gpu/src/LICH-train_rbm.cu:91) */
  double J;                /* coupling          (ref driver: J =  sin(theta)) */
  double alpha;            /* J_ij = J |i-j|^-alpha */
  int32_t pbc;             /* ref isPBC (distance rule gpu/include/impl_hamiltonians.cuh:146; L must be even) */
  int32_t order;           /* nqs_order */
  uint64_t seed;           /* ref seedNumber: key of the internal counter RNG, or seed of the yarn2 stream (nqs_set_rng) */
  int32_t device;          /* CUDA ordinal (ref: -dev) */
  int32_t flags;           /* NQS_FLAG_* */
  int64_t max_predrawn_steps; /* capacity (in proposals per chain) of the device buffer for nqs_set_uniforms; 0 -> none */
} nqs_config;

#define NQS_FLAG_NO_SR        1  /* sampler only (pynqs use): do not allocate O [K][P] nor the CG vectors */
#define NQS_FLAG_ACCEPT_LOG   2  /* keep accept masks of the most recent nqs_do_mcmc_steps/nqs_warm_up call (tests) */
#define NQS_FLAG_FORCE_GENERIC 4 /* use the generic (direct log cosh) kernels even where a specialised one exists */
#define NQS_FLAG_SETUP_FROM_O 16 /* SR setup sums by a pass over O (reference structure) instead of from the factors (spins, tanh theta) */
#define NQS_FLAG_TWO_PASS_SV  8  /* S*v as two streaming passes over O (reference structure) instead of the one-pass cluster kernel */
#define NQS_FLAG_STRUCTURED_SV 32 /* S*v from the factors of O (spins, tanh theta) as two fp64 tensor-core GEMMs; O [K][P] is neither
                                   * written nor allocated unless nqs_log_derivs asks for it (N <= 256).  Same results to rounding. */
#define NQS_FLAG_NO_DMMA      64 /* theta = S W + b with the scalar-FMA kernel instead of the fp64 tensor-core GEMM (A/B switch) */

/* statistics of one SR iteration.  ref: the row printed by propagate, gpu/include/optimizer.cuh:156-159 */
typedef struct nqs_sr_stats {
  double e_re, e_im;     /* <h> per site = conj(conjHavg) */
  double rsd;            /* sqrt((<|h|^2> - |<h>|^2)/|<h>|^2) */
  double lambda;         /* regulariser used */
  int32_t cg_iters;      /* S*p products inside the CG loop (the initial residual product is not counted) */
  int32_t finite;        /* 0 -> <h> not finite, nothing was updated */
  double cg_res2;        /* final |r|^2 */
  double cg_rhs2;        /* |F|^2 */
} nqs_sr_stats;

typedef struct nqs_sr_options {
  double lr;             /* ref deltaTau / -lr */
  double tol;            /* ref 1e-5  gpu/include/impl_optimizer.cuh:60 */
  int32_t max_iter;      /* ref 1000  gpu/include/conjugate_gradient.cuh:19 */
  int32_t fixed_iters;   /* >0: run exactly this many CG iterations ignoring tol (benchmarks / parity); 0 -> reference rule */
  double lambda;         /* <0: reference schedule max(100*0.9^p, 1e-2), p = 1,2,.. (impl_optimizer.cuh:72-78); >=0: use this */
  int32_t n_mc_steps;    /* ref nMCSteps / -nms: sweeps before measuring */
  int32_t apply_update;  /* 1: evolve(dx, lr) as the reference; 0: solve only (dx stays in the handle) */
} nqs_sr_options;

/* ---- lifetime ------------------------------------------------------------------------------------------------- */
nqs_status nqs_create(const nqs_config * cfg, nqs_handle ** out);
void nqs_destroy(nqs_handle * h);
/* message of the last failing call on this handle (h == NULL: last failing nqs_create of this thread) */
const char * nqs_last_error(const nqs_handle * h);
int32_t nqs_abi_version(void);
nqs_status nqs_sync(nqs_handle * h); /* cudaStreamSynchronize of the handle's stream */
/* The reference builds the ansatz first and hands it to the Hamiltonian/sampler and to the optimizer later
 * (gpu/src/LICH-train_rbm.cu:90-101); these three let a host mirror that order on one handle:
 *   nqs_set_hamiltonian  ref: LITFIChain ctor (h, J, alpha, isPBC) gpu/include/impl_hamiltonians.cuh:118-183 (J matrix, site ring)
 *   nqs_set_seed         ref: seedNumber of BaseParallelSampler (impl_mcmc_sampler.cuh:6-15); restarts the proposal counter,
 *                        keeps the generator kind
 *   nqs_enable_sr        ref: StochasticReconfigurationCG ctor (impl_optimizer.cuh:45-64): allocates O [K][P] + CG vectors on a
 *                        handle created with NQS_FLAG_NO_SR (no-op otherwise) */
nqs_status nqs_set_hamiltonian(nqs_handle * h, double hfield, double J, double alpha, int32_t pbc, int32_t order);
nqs_status nqs_set_seed(nqs_handle * h, uint64_t seed);
/* ref: TRNGWrapper<FloatType, trng::yarn2>(seedNumber, seedDistance, nChains) gpu/include/trng4cuda.cuh:41-54 + get_uniformDist
 * :62-65.  NQS_RNG_YARN2: chain k (GLOBAL id, so the stream does not depend on the number of GPUs) draws the uniforms of
 * trng::yarn2 after seed(seedNumber); jump(2ul*seedDistance*k), one uniform01_dist<double> draw per proposal -- the published
 * TRNG4 algorithm restated in csrc/yarn2.cuh (the library itself is absent offline: stream parity is unpinned).
 * NQS_RNG_PHILOX (default of nqs_create): Philox4x32-10 keyed by (seed, global chain, proposal), seed_distance ignored.
 * Restarts the proposal counter like nqs_set_seed.  A feed given by nqs_set_uniforms takes precedence over either generator. */
typedef enum nqs_rng { NQS_RNG_PHILOX = 0, NQS_RNG_YARN2 = 1 } nqs_rng;
nqs_status nqs_set_rng(nqs_handle * h, int32_t kind, uint64_t seed, uint64_t seed_distance);
nqs_status nqs_enable_sr(nqs_handle * h);
/* a NEW optimizer object in the reference starts with bp_ = 1 (lambda schedule) and dx = 0 (CG warm start),
 * gpu/include/impl_optimizer.cuh:45-64 */
nqs_status nqs_sr_reset(nqs_handle * h);

/* ---- parameters.  Layout = reference `variables_`: RBM [W (i*M+j) | a | b]  (gpu/include/impl_neural_quantum_state.cuh:33-38),
 *      FFNN [W1 (i*M+j) | b1 | w1o] (:772-776).  P = N*M+N+M resp. N*M+2M. ------------------------------------------------ */
nqs_status nqs_n_variables(const nqs_handle * h, int64_t * P);
nqs_status nqs_set_params(nqs_handle * h, const nqs_cdouble * params, int64_t P);
nqs_status nqs_get_params(nqs_handle * h, nqs_cdouble * params, int64_t P);
/* reference init law (ctor, :30-48 / :766-783) from a GIVEN seed (the reference seeds from the clock) */
nqs_status nqs_init_params_random(nqs_handle * h, uint64_t seed);
/* ref: RBM::load(prefix)/save(prefix, precision=10) :225-232,281-286 (Dw/Da/Db .dat), FFNN :931-937,985-991 (Dw1/Dw2/Db1).
 * A missing or wrong-sized file is NOT an error: the reference message is printed to stdout and the values are kept. */
nqs_status nqs_load_params(nqs_handle * h, const char * prefix);
nqs_status nqs_save_params(nqs_handle * h, const char * prefix, int32_t precision);

/* ---- sampler -------------------------------------------------------------------------------------------------- */
/* ref: Ansatz::initialize(lnpsi_dev, spins_dev) :67-91.  spins[K_loc*N] in {+1,-1}; NULL -> Neel if J>0 else all up
 * (LITFIChain::initialize_, gpu/include/impl_hamiltonians.cuh:192-204).  Computes theta, sa, lnpsi0. */
nqs_status nqs_initialize(nqs_handle * h, const int8_t * spins);
/* ref: BaseParallelSampler::warm_up(n) gpu/include/impl_mcmc_sampler.cuh:18-25 = initialize + the all-true
 * accept_next_state_ quirk (flips the machine's current flip index, 0 after construction, lnpsi0 left stale) + n sweeps. */
nqs_status nqs_warm_up(nqs_handle * h, int32_t n_sweeps, const int8_t * spins);
/* ref: do_mcmc_steps(n) :28-39; one sweep = N single-site proposals per chain. */
nqs_status nqs_do_mcmc_steps(nqs_handle * h, int32_t n_sweeps);
/* Pre-drawn uniforms u[steps][K_loc] replacing TRNGWrapper::get_uniformDist (gpu/include/trng4cuda.cuh:62-65): proposal t
 * (counted from this call) of chain k uses u[t][k].  u == NULL -> the handle's generator (nqs_set_rng; Philox4x32-10 by default).
 * u is consumed asynchronously: it must stay valid and unchanged until the sweeps that use it have completed (any later
 * synchronising call).  Pageable memory is staged through HBM (steps <= nqs_config.max_predrawn_steps); a page-locked
 * buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory) is read in place over PCIe by the sweep kernels, with no
 * staging copy and no limit on steps (NQS_UNIFORMS_ZEROCOPY=0 restores the copy). */
nqs_status nqs_set_uniforms(nqs_handle * h, const double * u, int64_t steps);
nqs_status nqs_get_spins(nqs_handle * h, int8_t * spins);                  /* ref: get_spinStates() (real part) */
nqs_status nqs_get_lnpsi(nqs_handle * h, nqs_cdouble * lnpsi);             /* ref: BaseParallelSampler::get_lnpsi() */
nqs_status nqs_get_theta(nqs_handle * h, nqs_cdouble * theta);             /* y_kj [K_loc][M] (tests) */
nqs_status nqs_get_accept_log(nqs_handle * h, uint8_t * acc, int64_t steps);/* [steps][K_loc], needs NQS_FLAG_ACCEPT_LOG */
/* ref: Ansatz::forward(int flipIdx, lnpsi_dev) :93-104 -- lnpsi of every chain with site `site` flipped (tests, ratios) */
nqs_status nqs_forward_flip(nqs_handle * h, int32_t site, nqs_cdouble * lnpsi1);
/* ref: Ansatz::forward(spins_dev, lnpsi_dev, saveSpinStates=false) :107-129 as used by pynqs get_lnpsi_for_fixed_spins
 * (gpu/src/pywrapping_sampler.cu:88-99).  The plain-RBM visible-bias quirk (sa from the member spins) is kept. */
nqs_status nqs_lnpsi_fixed_spins(nqs_handle * h, const int8_t * spins, nqs_cdouble * lnpsi);

/* ---- measurement + optimisation ------------------------------------------------------------------------------------ */
/* ref: get_htilda(htilda_dev) -> LITFIChain::get_htilda_ gpu/include/impl_hamiltonians.cuh:220-241.  htilda may be NULL. */
nqs_status nqs_local_energy(nqs_handle * h, nqs_cdouble * htilda);
/* ref: get_lnpsiGradients(O_dev) -> RBM::backward :146-154 / FFNN::backward :858-865 (GPU layout: W block transposed).
 * Fills the handle's O [K_loc][P]; O_host may be NULL (8.7 GB at N=128,M=256,K=16384 -- tests only). */
nqs_status nqs_log_derivs(nqs_handle * h, nqs_cdouble * O_host);
/* ref: SMatrixForCG::set_lnpsiGradients + dot, gpu/include/functor_for_CG.cuh:91-127: out = S v with the handle's O and the
 * given lambda (also recomputes <O>, diag).  aO/diag may be NULL. */
nqs_status nqs_smatrix_dot(nqs_handle * h, double lambda, const nqs_cdouble * v, nqs_cdouble * Sv, nqs_cdouble * aO, double * diag);
/* ref: one iteration of StochasticReconfigurationCG::propagate, gpu/include/optimizer.cuh:127-165. */
nqs_status nqs_sr_step(nqs_handle * h, const nqs_sr_options * opt, nqs_sr_stats * stats);
nqs_status nqs_sr_options_default(nqs_sr_options * opt);
nqs_status nqs_get_sr_vectors(nqs_handle * h, nqs_cdouble * F, nqs_cdouble * dx); /* either may be NULL */
/* ref: sampler.evolve(dx_dev, lr) -> update_variables :156-170 (FFNN :867-878) */
nqs_status nqs_evolve(nqs_handle * h, const nqs_cdouble * dx, double lr);

/* ---- multi-GPU: one handle per rank, chains sharded, NCCL all-reduce of SR sums and of O^H z per CG iteration --------- */
#define NQS_UNIQUE_ID_BYTES 128
nqs_status nqs_comm_get_unique_id(char id[NQS_UNIQUE_ID_BYTES]);                      /* rank 0, then broadcast by the host */
nqs_status nqs_comm_init(nqs_handle * h, int32_t n_ranks, int32_t rank, const char id[NQS_UNIQUE_ID_BYTES]);
/* Optional fast path for the per-CG-iteration exchange: the all-reduce of O_loc^H (O_loc v) is done inside the CG kernel over
 * NVLink peer memory.  After nqs_comm_init every rank exports the handle of its receive buffer, the host gathers the handles
 * of all ranks (rank order, NQS_IPC_HANDLE_BYTES each) and every rank imports them.  If the import fails
 * (NQS_ERR_UNSUPPORTED: no peer mapping between the processes) the engine keeps using ncclAllReduce. */
#define NQS_IPC_HANDLE_BYTES 64
nqs_status nqs_comm_p2p_export(nqs_handle * h, char handle_out[NQS_IPC_HANDLE_BYTES]);
nqs_status nqs_comm_p2p_import(nqs_handle * h, const char * handles /* [n_ranks][NQS_IPC_HANDLE_BYTES] */);
nqs_status nqs_comm_p2p_disable(nqs_handle * h); /* back to ncclAllReduce (all ranks must switch together) */

/* ---- introspection for benchmarks ----------------------------------------------------------------------------------- */
typedef struct nqs_timing {
  float sweep_ms, eloc_ms, oderiv_ms, setup_ms, cg_ms, update_ms; /* CUDA-event times of the phases of the last timed call */
  float rows_ms, cols_ms;   /* sums over CG iterations of the per-launch durations of the two O passes (O.v rows / O^H z columns) */
  int32_t rows_count, cols_count;
  int64_t kernel_launches;  /* kernels launched by this handle since creation */
} nqs_timing;
nqs_status nqs_get_timing(nqs_handle * h, nqs_timing * t);
nqs_status nqs_set_timing(nqs_handle * h, int32_t enabled);
/* CUDA events on the handle's stream for callers that time several calls as one region (slot 0..7) */
nqs_status nqs_event_record(nqs_handle * h, int32_t slot);
nqs_status nqs_event_elapsed_ms(nqs_handle * h, int32_t slot_begin, int32_t slot_end, float * ms); /* synchronises slot_end */
/* name of the kernel variant in use for stage "sweep" | "eloc" | "theta" | "sv" ("generic", "rbm_regs_j8", "fused_cs8_cpt8", ...) */
const char * nqs_kernel_variant(const nqs_handle * h, const char * stage);

/* Full-state checkpoint of one handle (one rank's shard) as a binary sidecar file.  The reference saves the parameters only
 * (text, 10 digits: sampler.save() every 100 iterations, gpu/include/optimizer.cuh:154-155,163), so a restarted run re-warms its
 * chains and restarts the lambda schedule; the sidecar carries what that loses: variables (exact doubles), spins, theta, lnpsi0,
 * sa, the per-chain "lnpsi0 is current" flags, site-ring position, the machine's index_, the generator (kind, seed, seedDistance, draw counter), bp_ of the lambda
 * schedule (optimizer.cuh:176) and the CG warm start dx (impl_optimizer.cuh:55).  A handle of the same (model, N, M, chains,
 * chain offset) that loads it continues BIT-IDENTICALLY.  Errors: NQS_ERR_IO (file), NQS_ERR_INVALID (shape mismatch). */
nqs_status nqs_checkpoint_save(nqs_handle * h, const char * path);
nqs_status nqs_checkpoint_load(nqs_handle * h, const char * path);

#ifdef __cplusplus
}
#endif
#endif /* NQS_B200_H */
