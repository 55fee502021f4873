// C++ host side above the C ABI (include/nqs_b200.h): the reference's ansatz / sampler / optimizer interface for the VMC hot
// path, same class names, method names, argument meaning and error behaviour, so that a reference driver compiles against
// this header by swapping its four includes.  Everything here only forwards to libnqs_b200.so; there is
// no arithmetic on the host and no CPU fallback.
//
//   spinhalf::RBM<double>, spinhalf::FFNN<double>      ref gpu/include/neural_quantum_state.cuh:17-59,151-194
//   spinhalf::RBMTrSymm<double>                         ref gpu/include/neural_quantum_state.cuh:62-105
//   spinhalf::RBMZ2PrSymm<double>, FFNNTrSymm<double>   ref gpu/include/neural_quantum_state.cuh:106-147,197-237
//   spinhalf::LITFIChain<Traits>                        ref gpu/include/hamiltonians.cuh:43-75 + mcmc_sampler.cuh:16-37
//   Sampler4SpinHalf<Traits>                            ref gpu/include/meas.cuh:11-28, impl_meas.cuh:5-41
//   StochasticReconfigurationCG<double>                 ref gpu/include/optimizer.cuh:112-181
//
// Differences a caller can see (all forced by the C-ABI boundary):
//   * pointer arguments are HOST pointers to std::complex<double> (the reference passes device pointers to thrust::complex);
//   * the device is chosen with nqs_host::set_device(dev) before constructing an ansatz (the reference calls cudaSetDevice);
//   * the uniform random numbers are the reference's: one trng::yarn2 engine per chain, seed(seedNumber) + jump(2*seedDistance*k),
//     restated in csrc/yarn2.cuh (TRNG4 itself is absent offline); NQS_RNG=philox in the environment selects the
//     engine's counter generator keyed by `seedNumber` instead (`seedDistance` is then ignored).
#pragma once
#include <chrono>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/nqs_b200.h"

namespace nqs_host
{
inline int & current_device() { static int dev = 0; return dev; }
inline void set_device(const int dev) { current_device() = dev; }

struct Error: public std::runtime_error
{
  nqs_status status;
  Error(nqs_status s, const std::string & m): std::runtime_error(m), status(s) {}
};

inline void check(const nqs_handle * h, const nqs_status rc, const char * what)
{
  if (rc == NQS_OK) return;
  const char * msg = nqs_last_error(h);
  throw Error(rc, std::string(what)+": "+(msg ? msg : ""));
}

inline nqs_cdouble * cptr(std::complex<double> * p) { return reinterpret_cast<nqs_cdouble*>(p); }
inline const nqs_cdouble * cptr(const std::complex<double> * p) { return reinterpret_cast<const nqs_cdouble*>(p); }

// common part of RBM<T> / FFNN<T>: one nqs_handle == ansatz + its chains on one GPU
class Ansatz
{
public:
  Ansatz(const int model, const int nInputs, const int nHiddens, const int nChains):
    knInputs(nInputs), knHiddens(nHiddens), knChains(nChains)
  {
    nqs_config cfg;
    cfg.abi_version = NQS_B200_ABI_VERSION;
    cfg.model = model;
    cfg.n_inputs = nInputs; cfg.n_hiddens = nHiddens; cfg.n_chains = nChains; cfg.n_chains_total = 0; cfg.chain_offset = 0;
    cfg.h = 0.0; cfg.J = 0.0; cfg.alpha = 0.0; cfg.pbc = 0; cfg.order = NQS_ORDER_CHECKERBOARD;
    cfg.seed = 0; cfg.device = current_device();
    cfg.flags = NQS_FLAG_NO_SR;          // O and the CG vectors are allocated when an optimizer first uses the sampler
    cfg.max_predrawn_steps = 0;
    // NQS_STRUCTURED_SV=1 in the environment opts plain RBM / FNN handles into the structured S*v; the tied-variable ansaetze have
    // no such form (their rows of O are not outer products), so the switch is hidden from the library while such a handle is made
    const bool tied = model != NQS_MODEL_RBM && model != NQS_MODEL_FFNN;
    const char * envs = tied ? std::getenv("NQS_STRUCTURED_SV") : nullptr;
    const std::string saved = envs ? envs : "";
    if (envs) unsetenv("NQS_STRUCTURED_SV");
    const nqs_status rc = nqs_create(&cfg, &h_);
    if (envs) setenv("NQS_STRUCTURED_SV", saved.c_str(), 1);
    if (rc != NQS_OK)
    {
      const char * msg = nqs_last_error(nullptr);
      throw Error(rc, std::string("nqs_create: ")+(msg ? msg : ""));
    }
    // the reference ctor draws the parameters from a clock-seeded generator (impl_neural_quantum_state.cuh:30-31)
    const uint64_t seed = (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();
    check(h_, nqs_init_params_random(h_, seed), "nqs_init_params_random");
    int64_t P = 0;
    check(h_, nqs_n_variables(h_, &P), "nqs_n_variables");
    knVariables = (int)P;
  }
  Ansatz(const Ansatz &) = delete;
  Ansatz & operator=(const Ansatz &) = delete;
  ~Ansatz() { nqs_destroy(h_); }

  void initialize(std::complex<double> * lnpsi, const int8_t * spinStates = nullptr)
  {
    check(h_, nqs_initialize(h_, spinStates), "initialize");
    if (lnpsi) check(h_, nqs_get_lnpsi(h_, cptr(lnpsi)), "get_lnpsi");
  }
  // lnpsi of every chain with site `spinFlipIndex` flipped
  void forward(const int spinFlipIndex, std::complex<double> * lnpsi) { check(h_, nqs_forward_flip(h_, spinFlipIndex, cptr(lnpsi)), "forward(int)"); }
  // amplitudes of given configurations; saveSpinStates=true also makes them the chain state (== initialize)
  void forward(const int8_t * spinStates, std::complex<double> * lnpsi, const bool saveSpinStates = true)
  {
    if (saveSpinStates) initialize(lnpsi, spinStates);
    else check(h_, nqs_lnpsi_fixed_spins(h_, spinStates, cptr(lnpsi)), "forward(spins)");
  }
  void backward(std::complex<double> * lnpsiGradients)
  {
    check(h_, nqs_enable_sr(h_), "enable_sr");
    check(h_, nqs_log_derivs(h_, cptr(lnpsiGradients)), "backward");
  }
  void update_variables(const std::complex<double> * derivativeLoss, const double learningRate)
  { check(h_, nqs_evolve(h_, cptr(derivativeLoss), learningRate), "update_variables"); }
  void save(const std::string prefix, const int precision = 10) const { check(h_, nqs_save_params(h_, prefix.c_str(), precision), "save"); }
  void load(const std::string prefix) { check(h_, nqs_load_params(h_, prefix.c_str()), "load"); }
  void copy_to(Ansatz & other) const
  {
    if (other.knVariables != knVariables) throw std::length_error("copy_to: different numbers of variables");
    std::vector<std::complex<double> > v((size_t)knVariables);
    check(h_, nqs_get_params(h_, cptr(v.data()), knVariables), "get_params");
    check(other.h_, nqs_set_params(other.h_, cptr(v.data()), knVariables), "set_params");
  }
  std::vector<int8_t> get_spinStates() const
  {
    std::vector<int8_t> s((size_t)knChains*knInputs);
    check(h_, nqs_get_spins(h_, s.data()), "get_spinStates");
    return s;
  }
  int get_nChains() const { return knChains; }
  int get_nInputs() const { return knInputs; }
  int get_nHiddens() const { return knHiddens; }
  int get_nVariables() const { return knVariables; }
  nqs_handle * handle() const { return h_; }

private:
  const int knInputs, knHiddens, knChains;
  int knVariables = 0;
  nqs_handle * h_ = nullptr;
};
} // namespace nqs_host

namespace spinhalf
{
template <typename FloatType>
class RBM: public nqs_host::Ansatz
{
  static_assert(sizeof(FloatType) == sizeof(double), "libnqs_b200 computes in fp64 only");
public:
  RBM(const int nInputs, const int nHiddens, const int nChains): nqs_host::Ansatz(NQS_MODEL_RBM, nInputs, nHiddens, nChains) {}
};

template <typename FloatType>
class FFNN: public nqs_host::Ansatz
{
  static_assert(sizeof(FloatType) == sizeof(double), "libnqs_b200 computes in fp64 only");
public:
  FFNN(const int nInputs, const int nHiddens, const int nChains): nqs_host::Ansatz(NQS_MODEL_FFNN, nInputs, nHiddens, nChains) {}
};

// ref: RBMTrSymm<T>(nInputs, alpha, nChains), gpu/include/neural_quantum_state.cuh:62-105: `alpha` filters of N tied weights each,
// nVariables = N*alpha + 1 + alpha; save / load take the FILE path (one file with every variable, impl :474-517)
template <typename FloatType>
class RBMTrSymm: public nqs_host::Ansatz
{
  static_assert(sizeof(FloatType) == sizeof(double), "libnqs_b200 computes in fp64 only");
public:
  RBMTrSymm(const int nInputs, const int alpha, const int nChains):
    nqs_host::Ansatz(NQS_MODEL_RBMTRSYMM, nInputs, alpha*nInputs, nChains), kAlpha(alpha) {}
  int get_alpha() const { return kAlpha; }
private:
  const int kAlpha;
};

// ref: RBMZ2PrSymm<T>(nInputs, alpha, nChains), gpu/include/neural_quantum_state.cuh:106-147: Z2- and parity-symmetric RBM,
// nVariables = N*alpha + alpha (4 hidden units per filter); save / load take the FILE path (impl :680-722)
template <typename FloatType>
class RBMZ2PrSymm: public nqs_host::Ansatz
{
  static_assert(sizeof(FloatType) == sizeof(double), "libnqs_b200 computes in fp64 only");
public:
  RBMZ2PrSymm(const int nInputs, const int alpha, const int nChains):
    nqs_host::Ansatz(NQS_MODEL_RBMZ2PRSYMM, nInputs, 4*alpha, nChains), kAlpha(alpha) {}
  int get_alpha() const { return kAlpha; }
private:
  const int kAlpha;
};

// ref: FFNNTrSymm<T>(nInputs, alpha, nChains), gpu/include/neural_quantum_state.cuh:197-237: translation-symmetric FNN,
// nVariables = N*alpha + 2*alpha; save / load take the FILE path (impl :1167-1209)
template <typename FloatType>
class FFNNTrSymm: public nqs_host::Ansatz
{
  static_assert(sizeof(FloatType) == sizeof(double), "libnqs_b200 computes in fp64 only");
public:
  FFNNTrSymm(const int nInputs, const int alpha, const int nChains):
    nqs_host::Ansatz(NQS_MODEL_FFNNTRSYMM, nInputs, alpha*nInputs, nChains), kAlpha(alpha) {}
  int get_alpha() const { return kAlpha; }
private:
  const int kAlpha;
};

// ref: LITFIChain<TraitsClass> (gpu/include/hamiltonians.cuh:43-75) with the BaseParallelSampler interface (mcmc_sampler.cuh:21-29)
template <typename TraitsClass>
class LITFIChain
{
  using AnsatzType = typename TraitsClass::AnsatzType;
  using FloatType = typename TraitsClass::FloatType;
public:
  LITFIChain(AnsatzType & machine, const int L, const FloatType h, const FloatType J, const double alpha, const bool isPBC,
    const unsigned long seedNumber, const unsigned long seedDistance, const std::string prefix = "./"):
    machine_(machine), kprefix(prefix)
  {
    if (L != machine.get_nInputs())
      throw std::length_error("machine.get_nInputs() is not the same as L!");
    const nqs_status rc = nqs_set_hamiltonian(machine.handle(), h, J, alpha, isPBC ? 1 : 0, NQS_ORDER_CHECKERBOARD);
    if (rc == NQS_ERR_INVALID) throw std::invalid_argument(nqs_last_error(machine.handle()));
    nqs_host::check(machine.handle(), rc, "nqs_set_hamiltonian");
    // ref: BaseParallelSampler(..., seedNumber, seedDistance) -> TRNGWrapper<FloatType, trng::yarn2> rng_ (mcmc_sampler.cuh:34)
    const char * env = std::getenv("NQS_RNG");
    const int kind = (env && !std::strcmp(env, "philox")) ? NQS_RNG_PHILOX : NQS_RNG_YARN2;
    nqs_host::check(machine.handle(), nqs_set_rng(machine.handle(), kind, (uint64_t)seedNumber, (uint64_t)seedDistance), "nqs_set_rng");
  }
  void warm_up(const int nMCSteps = 100) { nqs_host::check(hd(), nqs_warm_up(hd(), nMCSteps, nullptr), "warm_up"); }
  void do_mcmc_steps(const int nMCSteps = 1) { nqs_host::check(hd(), nqs_do_mcmc_steps(hd(), nMCSteps), "do_mcmc_steps"); }
  std::vector<std::complex<double> > get_lnpsi()
  {
    std::vector<std::complex<double> > v((size_t)machine_.get_nChains());
    nqs_host::check(hd(), nqs_get_lnpsi(hd(), nqs_host::cptr(v.data())), "get_lnpsi");
    return v;
  }
  void get_htilda(std::complex<double> * htilda) { nqs_host::check(hd(), nqs_local_energy(hd(), nqs_host::cptr(htilda)), "get_htilda"); }
  void get_lnpsiGradients(std::complex<double> * lnpsiGradients) { machine_.backward(lnpsiGradients); }
  int get_nChains() const { return machine_.get_nChains(); }
  void evolve(const std::complex<double> * trueGradients, const double learningRate) { machine_.update_variables(trueGradients, learningRate); }
  void save() const { machine_.save(kprefix); }
  nqs_handle * handle() const { return machine_.handle(); }
private:
  nqs_handle * hd() const { return machine_.handle(); }
  AnsatzType & machine_;
  const std::string kprefix;
};
} // namespace spinhalf

// ref: Sampler4SpinHalf<TraitsClass> (gpu/include/meas.cuh:11-28, impl_meas.cuh:5-41): the Hamiltonian-free sampler behind the
// measurement programs and pynqs' PySampler -- sequential site ring 1,2,..,N-1,0, random +-1 initial spins (the reference draws
// them inside psi.initialize(lnpsi) from a clock-seeded generator, neural_quantum_state.cuh:239-249; here from seedNumber so that
// a run is reproducible), the BaseParallelSampler interface otherwise.
template <typename TraitsClass>
class Sampler4SpinHalf
{
  using AnsatzType = typename TraitsClass::AnsatzType;
public:
  Sampler4SpinHalf(AnsatzType & psi, const unsigned long seedNumber, const unsigned long seedDistance): psi_(psi), seed_(seedNumber)
  {
    nqs_host::check(psi.handle(), nqs_set_hamiltonian(psi.handle(), 0.0, 0.0, 0.0, 0, NQS_ORDER_SEQUENTIAL), "nqs_set_hamiltonian");
    const char * env = std::getenv("NQS_RNG");
    const int kind = (env && !std::strcmp(env, "philox")) ? NQS_RNG_PHILOX : NQS_RNG_YARN2;
    nqs_host::check(psi.handle(), nqs_set_rng(psi.handle(), kind, (uint64_t)seedNumber, (uint64_t)seedDistance), "nqs_set_rng");
  }
  void warm_up(const int nMCSteps = 100)
  {
    std::mt19937_64 ran(seed_*0x9E3779B97F4A7C15ull+0x5EEDull);
    std::vector<int8_t> spins((size_t)psi_.get_nChains()*psi_.get_nInputs());
    for (auto & s : spins) s = (ran()&1ull) ? 1 : -1;
    nqs_host::check(hd(), nqs_warm_up(hd(), nMCSteps, spins.data()), "warm_up");
  }
  void do_mcmc_steps(const int nMCSteps = 1) { nqs_host::check(hd(), nqs_do_mcmc_steps(hd(), nMCSteps), "do_mcmc_steps"); }
  std::vector<std::complex<double> > get_lnpsi()
  {
    std::vector<std::complex<double> > v((size_t)psi_.get_nChains());
    nqs_host::check(hd(), nqs_get_lnpsi(hd(), nqs_host::cptr(v.data())), "get_lnpsi");
    return v;
  }
  std::vector<int8_t> get_quantumStates() const { return psi_.get_spinStates(); }
  int get_nInputs() const { return psi_.get_nInputs(); }
  int get_nChains() const { return psi_.get_nChains(); }
private:
  nqs_handle * hd() const { return psi_.handle(); }
  AnsatzType & psi_;
  const unsigned long seed_;
};

// ref: StochasticReconfigurationCG<FloatType> (gpu/include/optimizer.cuh:112-181): the whole iteration body runs on the device
// inside nqs_sr_step; this class keeps the reference's loop, stop rules and stdout.
template <typename FloatType>
class StochasticReconfigurationCG
{
public:
  StochasticReconfigurationCG(const int nChains, const int nVariables): knChains(nChains), knVariables(nVariables) {}
  template <typename SamplerType>
  void propagate(SamplerType & sampler, const int nIteration, const int nMCSteps, const FloatType deltaTau, const FloatType RSDcutoff,
    const int nrec = 100)
  {
    nqs_handle * h = sampler.handle();
    int64_t P = 0;
    nqs_host::check(h, nqs_n_variables(h, &P), "nqs_n_variables");
    if (P != knVariables || sampler.get_nChains() != knChains)
      throw std::length_error("StochasticReconfigurationCG: (nChains, nVariables) do not match the sampler");
    nqs_host::check(h, nqs_enable_sr(h), "nqs_enable_sr");
    if (!attached_)
    { // a new optimizer object starts its lambda schedule and its CG warm start from scratch (impl_optimizer.cuh:45-64)
      nqs_host::check(h, nqs_sr_reset(h), "nqs_sr_reset");
      attached_ = true;
    }
    nqs_sr_options opt;
    nqs_sr_options_default(&opt);
    opt.lr = deltaTau;
    opt.n_mc_steps = nMCSteps;
    std::cout << "# of loop\t" << "<H>" << std::endl << std::setprecision(7);
    for (int n = 0; n < nIteration; ++n)
    {
      nqs_sr_stats st;
      nqs_host::check(h, nqs_sr_step(h, &opt, &st), "nqs_sr_step");
      if (!st.finite)
      {
        std::cout << "# \"Havg\" has non-value type. We stop here." << std::endl;
        return;
      }
      if (n%nrec == (nrec-1))
        sampler.save();
      std::cout << std::setw(5) << (n+1) << std::setw(16) << st.e_re << std::setw(16) << st.rsd << std::endl << std::flush;
      if (st.rsd < RSDcutoff)
      {
        std::cout << "# We got a converged solution." << std::endl;
        sampler.save();
        break;
      }
    }
  }
private:
  const int knChains, knVariables;
  bool attached_ = false;
};
