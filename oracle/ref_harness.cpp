// oracle/_ref/libnqs_ref.so -- the REFERENCE'S OWN CPU implementation behind a small C ABI (TEST INFRASTRUCTURE).
//
// This translation unit #includes the reference headers where they lie under /root/reference/cpu/include
// (nothing is copied into this repository) and instantiates, unmodified:
//   spinhalf::RBM<double>, spinhalf::FFNN<double>,     cpu/include/neural_quantum_state.hpp, impl_neural_quantum_state.hpp
//   spinhalf::RBMTrSymm<double>, spinhalf::FFNNTrSymm<double>
//   BaseParallelSampler<...>                            cpu/include/mcmc_sampler.hpp, impl_mcmc_sampler.hpp
//   SMatrixForCG<double>, ConjugateGradient<double>     cpu/include/functor_for_CG.hpp, conjugate_gradient.hpp
// What the CPU tree does NOT have is written here as a thin shim that restates the GPU tree:
//   LITFIChainCPU  <- gpu/include/impl_hamiltonians.cuh:118-259 (J matrix, Neel init, checkerboard ring, E_loc / L)
//   sr_step        <- gpu/include/optimizer.cuh:125-166 with the GPU solver settings tol = 1e-5, maxIter = 1000
//                     (gpu/include/impl_optimizer.cuh:60, gpu/include/conjugate_gradient.cuh:19)
// TRNG4 is replaced by oracle/shim/trng (pre-drawn uniform feed).  Private members of the reference classes are
// read through `#define private public` AFTER all standard headers were included; the reference files themselves
// are untouched.  Used by tests/ (to pin oracle/nqs_oracle.py), tests/golden/make_golden.py and bench.py's
// `--impl reference` / cpu_baseline legs.  Never linked into the product library.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <complex>
#include <cstring>
#include <exception>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#include <omp.h>

#include "trng/feed.hpp"

#define private public
#define protected public
#include "mcmc_sampler.hpp"
#include "neural_quantum_state.hpp"
#include "functor_for_CG.hpp"
#include "conjugate_gradient.hpp"
#undef private
#undef protected

typedef std::complex<double> cdouble;

namespace
{
template <typename Ansatz>
struct Traits
{
  using AnsatzType = Ansatz;
  using FloatType = double;
};

// CPU restatement of gpu/include/hamiltonians.cuh:43-75 + impl_hamiltonians.cuh:118-259 on top of the reference's CPU
// BaseParallelSampler (CRTP hooks named as in cpu/include/hamiltonians.hpp:14-38).
template <typename TraitsClass>
class LITFIChainCPU: public BaseParallelSampler<LITFIChainCPU, TraitsClass>
{
  USING_OF_BASE_PARALLEL_SAMPLER(LITFIChainCPU, TraitsClass);
  using AnsatzType = typename TraitsClass::AnsatzType;
public:
  LITFIChainCPU(AnsatzType & machine, const int L, const double h, const double J, const double alpha, const bool isPBC,
    const int orderKind):
    BaseParallelSampler<LITFIChainCPU, TraitsClass>(machine.get_nInputs(), machine.get_nChains(), 1ul, 0ul),
    machine_(machine), kL(L), knChains(machine.get_nChains()), kh(h), kJ(J), Jmatrix_(L*L, 0.0), list_(L), useCustomInit_(false)
  {
    if (kL != machine.get_nInputs())
      throw std::length_error("machine.get_nInputs() is not the same as L!");
    if (isPBC && kL%2 == 1)
      throw std::invalid_argument("kL%2 == 1 (set \"isPBC\" to \"false\".)");
    for (int i=0; i<kL; ++i)
      for (int j=i+1; j<kL; ++j)
      {
        double dist = (j-i);
        if (isPBC)
          dist = (((j-i)<kL/2) ? (j-i) : kL-(j-i));
        Jmatrix_[i*kL+j] = J*std::pow(dist, -alpha);
        Jmatrix_[j*kL+i] = Jmatrix_[i*kL+j];
      }
    for (int i=0; i<kL; ++i)
      list_[i].set_item(i);
    int idx0 = 0;
    if (orderKind == 0)
    { // checkerboard ring, impl_hamiltonians.cuh:163-180
      for (int i=0; i<kL; i+=2) { list_[idx0].set_nextptr(&list_[i]); idx0 = i; }
      for (int i=1; i<kL; i+=2) { list_[idx0].set_nextptr(&list_[i]); idx0 = i; }
    }
    else
    { // sequential ring of Sampler4SpinHalf, gpu/include/impl_meas.cuh:12-21
      for (int i=0; i<kL; ++i) { list_[idx0].set_nextptr(&list_[i]); idx0 = i; }
    }
    list_[idx0].set_nextptr(&list_[0]);
    idxptr_ = &list_[0];
  }

  void set_initial_spins(const double * spins)
  {
    customInit_.assign(spins, spins+static_cast<size_t>(kL)*knChains);
    useCustomInit_ = true;
  }

  void get_htilda(cdouble * htilda)
  {
    const cdouble * s = machine_.get_spinStates();
    // 1/2 sum_ij s_i J_ij s_j   (impl_hamiltonians.cuh:226-231, kernel :871-887)
    for (int k=0; k<knChains; ++k)
    {
      double acc = 0;
      for (int i=0; i<kL; ++i)
      {
        double sj = 0;
        for (int j=0; j<kL; ++j)
          sj += Jmatrix_[i*kL+j]*s[k*kL+j].real();
        acc += sj*s[k*kL+i].real();
      }
      htilda[k] = 0.5*acc;
    }
    // transverse field (impl_hamiltonians.cuh:233-238, kernel :857-869)
    for (int i=0; i<kL; ++i)
    {
      machine_.forward(i, &lnpsi1_[0]);
      for (int k=0; k<knChains; ++k)
        htilda[k] += kh*std::exp(lnpsi1_[k]-lnpsi0_[k]);
    }
    for (int k=0; k<knChains; ++k)
      htilda[k] = (1.0/kL)*htilda[k]; // :240
  }
  const std::vector<cdouble> & tracked_lnpsi() const { return lnpsi0_; }
  void get_lnpsiGradients(cdouble * g) { machine_.backward(g); }
  void evolve(const cdouble * dx, const double lr) { machine_.update_variables(dx, lr); }

  void initialize(cdouble * lnpsi)
  {
    std::vector<cdouble> spins(static_cast<size_t>(kL)*knChains, cdouble(1.0, 0.0));
    if (useCustomInit_)
      for (size_t n=0; n<spins.size(); ++n)
        spins[n] = customInit_[n];
    else if (kJ > 0) // Neel, impl_hamiltonians.cuh:196-201
      for (int k=0; k<knChains; ++k)
        for (int i=0; i<kL; ++i)
          spins[k*kL+i] = ((i%2 == 0) ? 1.0 : -1.0);
    machine_.initialize(lnpsi, spins.data());
  }
  void sampling(cdouble * lnpsi)
  {
    idxptr_ = idxptr_->next_ptr();
    machine_.forward(idxptr_->get_item(), lnpsi);
  }
  void accept_next_state(const std::vector<bool> & updateList) { machine_.spin_flip(updateList); }

  AnsatzType & machine_;
  const int kL, knChains;
  const double kh, kJ;
  std::vector<double> Jmatrix_;
  std::vector<OneWayLinkedIndex<> > list_;
  OneWayLinkedIndex<> * idxptr_;
  std::vector<double> customInit_;
  bool useCustomInit_;
};

struct SRStatsC
{
  double e_re, e_im, rsd, lambda;
  int cg_iters, finite;
};

// counts Mat.dot calls: one per CG iteration + 1 for the initial residual
template <typename Mat>
struct CountingMatrix
{
  Mat & m; int ndot;
  explicit CountingMatrix(Mat & m_): m(m_), ndot(0) {}
  void dot(const cdouble * a, cdouble * b) { ++ndot; m.dot(a, b); }
  void applyPrecond(const cdouble * r, cdouble * x) const { m.applyPrecond(r, x); }
};

struct CtxBase
{
  virtual ~CtxBase() {}
  virtual void set_params(const cdouble * v) = 0;
  virtual void get_params(cdouble * v) = 0;
  virtual void load(const char * prefix) = 0;
  virtual void save(const char * prefix, int prec) = 0;
  virtual void set_initial_spins(const double * s) = 0;
  virtual void warm_up(int n) = 0;
  virtual void do_mcmc_steps(int n) = 0;
  virtual void get_lnpsi(cdouble * out) = 0;
  virtual void get_spins(double * out) = 0;
  virtual void get_y(cdouble * out) = 0;
  virtual void forward_flip(int idx, cdouble * out) = 0;
  virtual void get_htilda(cdouble * out) = 0;
  virtual void get_gradients(cdouble * out) = 0;
  virtual void smatrix_set(const cdouble * O, double lambda, cdouble * aO, double * diag) = 0;
  virtual void smatrix_dot(const cdouble * v, cdouble * out) = 0;
  virtual void smatrix_diag(double * out) = 0;
  virtual void sr_step(int nms, double lr, int maxIter, double tol, double lambdaOverride, SRStatsC * st, cdouble * F, cdouble * dx) = 0;
  virtual void evolve(const cdouble * dx, double lr) = 0;
  trng::uniform_feed feed;
  int N, M, K, P;
};

template <typename Ansatz>
struct Ctx: public CtxBase
{
  using T = Traits<Ansatz>;
  std::unique_ptr<Ansatz> machine;
  std::unique_ptr<LITFIChainCPU<T> > sampler;
  std::unique_ptr<SMatrixForCG<double> > smat;
  std::vector<cdouble> ht, O, aO, F, dx, ones;
  double bp;

  Ctx(int N_, int M_, int K_, double h, double J, double alpha, int pbc, int orderKind): bp(1.0)
  {
    N = N_; M = M_; K = K_;
    feed.nChains = K;
    trng::current_feed() = &feed; // engines capture the pointer in seed()
    machine.reset(new Ansatz(N, M, K));
    sampler.reset(new LITFIChainCPU<T>(*machine, N, h, J, alpha, pbc != 0, orderKind));
    trng::current_feed() = nullptr;
    P = machine->get_nVariables();
    smat.reset(new SMatrixForCG<double>(K, P));
    ht.resize(K); aO.resize(P); F.resize(P); dx.assign(P, cdouble(0, 0)); ones.assign(K, cdouble(1, 0));
  }
  void set_params(const cdouble * v) override { std::copy(v, v+P, machine->variables_.begin()); }
  void get_params(cdouble * v) override { std::copy(machine->variables_.begin(), machine->variables_.end(), v); }
  void load(const char * prefix) override;
  void save(const char * prefix, int prec) override;
  void set_initial_spins(const double * s) override { sampler->set_initial_spins(s); }
  void warm_up(int n) override { sampler->warm_up(n); }
  void do_mcmc_steps(int n) override { sampler->do_mcmc_steps(n); }
  void get_lnpsi(cdouble * out) override { std::copy(sampler->tracked_lnpsi().begin(), sampler->tracked_lnpsi().end(), out); }
  void get_spins(double * out) override
  {
    const cdouble * s = machine->get_spinStates();
    for (size_t n=0; n<static_cast<size_t>(K)*N; ++n) out[n] = s[n].real();
  }
  void get_y(cdouble * out) override { std::copy(machine->y_.begin(), machine->y_.end(), out); }
  void forward_flip(int idx, cdouble * out) override { machine->forward(idx, out); }
  void get_htilda(cdouble * out) override { sampler->get_htilda(out); }
  void get_gradients(cdouble * out) override { sampler->get_lnpsiGradients(out); }
  void smatrix_set(const cdouble * Oin, double lambda, cdouble * aOout, double * diag) override
  {
    O.assign(Oin, Oin+static_cast<size_t>(K)*P);
    smat->set_lnpsiGradients(O.data(), lambda);
    std::copy(smat->avglnpsiGradients_.begin(), smat->avglnpsiGradients_.end(), aOout);
    std::copy(smat->diag_.begin(), smat->diag_.end(), diag);
  }
  void smatrix_dot(const cdouble * v, cdouble * out) override { smat->dot(v, out); }
  void smatrix_diag(double * out) override { std::copy(smat->diag_.begin(), smat->diag_.end(), out); }
  void evolve(const cdouble * d, double lr) override { sampler->evolve(d, lr); }

  // body of StochasticReconfigurationCG::propagate, gpu/include/optimizer.cuh:127-165 (one iteration)
  void sr_step(int nms, double lr, int maxIter, double tol, double lambdaOverride, SRStatsC * st, cdouble * Fout, cdouble * dxout) override
  {
    const cdouble oneOverTotalMeas = 1.0/static_cast<double>(K), kzero(0, 0);
    O.resize(static_cast<size_t>(K)*P);
    sampler->do_mcmc_steps(nms);
    sampler->get_htilda(ht.data());
    sampler->get_lnpsiGradients(O.data());
    for (int k=0; k<K; ++k)
      ht[k] = std::conj(ht[k]);
    const cdouble conjHavg = oneOverTotalMeas.real()*std::accumulate(ht.begin(), ht.end(), kzero);
    st->e_re = conjHavg.real(); st->e_im = -conjHavg.imag(); st->finite = 1; st->cg_iters = 0; st->rsd = 0; st->lambda = 0;
    if (!std::isfinite(conjHavg.real()))
    {
      st->finite = 0;
      return;
    }
    blas::gemv(P, K, oneOverTotalMeas, O.data(), ones.data(), kzero, aO.data());
    blas::gemv(P, K, oneOverTotalMeas, O.data(), ht.data(), kzero, F.data());
    for (int i=0; i<P; ++i)
      F[i] = std::conj(F[i]-conjHavg*aO[i]); // SR__FStep2__, gpu/include/impl_optimizer.cuh:82-96
    // schedular_, gpu/include/impl_optimizer.cuh:72-78
    double lambda = lambdaOverride;
    if (lambdaOverride < 0)
    {
      bp *= 0.9;
      lambda = 100.0*bp;
      lambda = ((lambda > 1e-2) ? lambda : 1e-2);
    }
    st->lambda = lambda;
    smat->set_lnpsiGradients(O.data(), lambda);
    ConjugateGradient<double> cg(P, tol, maxIter);
    CountingMatrix<SMatrixForCG<double> > cm(*smat);
    cg.solve(cm, F.data(), dx.data());
    st->cg_iters = cm.ndot-1;
    sampler->evolve(dx.data(), lr);
    double h2 = 0;
    for (int k=0; k<K; ++k)
      h2 += std::norm(ht[k]);
    st->rsd = std::sqrt((h2/K-std::norm(conjHavg))/std::norm(conjHavg));
    if (Fout) std::copy(F.begin(), F.end(), Fout);
    if (dxout) std::copy(dx.begin(), dx.end(), dxout);
  }
};

template <> void Ctx<spinhalf::RBM<double> >::load(const char * prefix)
{ // same three files as the GPU RBM::load(prefix), gpu/include/impl_neural_quantum_state.cuh:281-286
  const std::string p(prefix);
  machine->load(spinhalf::RBMDataType::W, p+"Dw.dat");
  machine->load(spinhalf::RBMDataType::V, p+"Da.dat");
  machine->load(spinhalf::RBMDataType::H, p+"Db.dat");
}
template <> void Ctx<spinhalf::RBM<double> >::save(const char * prefix, int prec)
{
  const std::string p(prefix);
  machine->save(spinhalf::RBMDataType::W, p+"Dw.dat", prec);
  machine->save(spinhalf::RBMDataType::V, p+"Da.dat", prec);
  machine->save(spinhalf::RBMDataType::H, p+"Db.dat", prec);
}
template <> void Ctx<spinhalf::FFNN<double> >::load(const char * prefix)
{ // gpu/include/impl_neural_quantum_state.cuh:985-991
  const std::string p(prefix);
  machine->load(spinhalf::FFNNDataType::W1, p+"Dw1.dat");
  machine->load(spinhalf::FFNNDataType::W2, p+"Dw2.dat");
  machine->load(spinhalf::FFNNDataType::B1, p+"Db1.dat");
}
template <> void Ctx<spinhalf::FFNN<double> >::save(const char * prefix, int prec)
{
  const std::string p(prefix);
  machine->save(spinhalf::FFNNDataType::W1, p+"Dw1.dat", prec);
  machine->save(spinhalf::FFNNDataType::W2, p+"Dw2.dat", prec);
  machine->save(spinhalf::FFNNDataType::B1, p+"Db1.dat", prec);
}
// the tied-variable ansaetze of the CPU tree keep every variable in ONE file named by the path itself
// (cpu/include/impl_neural_quantum_state.hpp:515-548, 1159-1192), like their GPU counterparts
template <> void Ctx<spinhalf::RBMTrSymm<double> >::load(const char * prefix) { machine->load(std::string(prefix)); }
template <> void Ctx<spinhalf::RBMTrSymm<double> >::save(const char * prefix, int prec) { machine->save(std::string(prefix), prec); }
template <> void Ctx<spinhalf::FFNNTrSymm<double> >::load(const char * prefix) { machine->load(std::string(prefix)); }
template <> void Ctx<spinhalf::FFNNTrSymm<double> >::save(const char * prefix, int prec) { machine->save(std::string(prefix), prec); }
} // namespace

#define REF_TRY(stmt) try { stmt; return 0; } catch (const std::exception & e) { std::cerr << "# ref_harness: " << e.what() << std::endl; return 1; }

extern "C"
{
// model: 0 = RBM, 1 = FFNN (CPU gradient layout is natural i*M+j), 2 = RBMTrSymm, 4 = FFNNTrSymm (for these two
// M is the number of filters alpha; numbering as nqs_model of include/nqs_b200.h -- the CPU tree has no RBMZ2PrSymm).
// order: 0 = checkerboard, 1 = sequential.
void * ref_create(int model, int N, int M, int K, double h, double J, double alpha, int pbc, int order)
{
  try
  {
    if (model == 0) return new Ctx<spinhalf::RBM<double> >(N, M, K, h, J, alpha, pbc, order);
    if (model == 1) return new Ctx<spinhalf::FFNN<double> >(N, M, K, h, J, alpha, pbc, order);
    if (model == 2) return new Ctx<spinhalf::RBMTrSymm<double> >(N, M, K, h, J, alpha, pbc, order);
    if (model == 4) return new Ctx<spinhalf::FFNNTrSymm<double> >(N, M, K, h, J, alpha, pbc, order);
  }
  catch (const std::exception & e) { std::cerr << "# ref_harness: " << e.what() << std::endl; }
  return nullptr;
}
void ref_destroy(void * c) { delete static_cast<CtxBase*>(c); }
int ref_n_variables(void * c) { return static_cast<CtxBase*>(c)->P; }
int ref_set_uniforms(void * c, const double * u, long steps)
{ // u[steps][K] must stay alive while the sampler draws; the draw counter is NOT reset (engines keep counting).
  CtxBase * x = static_cast<CtxBase*>(c);
  x->feed.u = u; x->feed.steps = steps;
  return 0;
}
int ref_set_params(void * c, const cdouble * v) { REF_TRY(static_cast<CtxBase*>(c)->set_params(v)) }
int ref_get_params(void * c, cdouble * v) { REF_TRY(static_cast<CtxBase*>(c)->get_params(v)) }
int ref_load(void * c, const char * prefix) { REF_TRY(static_cast<CtxBase*>(c)->load(prefix)) }
int ref_save(void * c, const char * prefix, int prec) { REF_TRY(static_cast<CtxBase*>(c)->save(prefix, prec)) }
int ref_set_initial_spins(void * c, const double * s) { REF_TRY(static_cast<CtxBase*>(c)->set_initial_spins(s)) }
int ref_warm_up(void * c, int n) { REF_TRY(static_cast<CtxBase*>(c)->warm_up(n)) }
int ref_do_mcmc_steps(void * c, int n) { REF_TRY(static_cast<CtxBase*>(c)->do_mcmc_steps(n)) }
int ref_get_lnpsi(void * c, cdouble * out) { REF_TRY(static_cast<CtxBase*>(c)->get_lnpsi(out)) }
int ref_get_spins(void * c, double * out) { REF_TRY(static_cast<CtxBase*>(c)->get_spins(out)) }
int ref_get_y(void * c, cdouble * out) { REF_TRY(static_cast<CtxBase*>(c)->get_y(out)) }
int ref_forward_flip(void * c, int idx, cdouble * out) { REF_TRY(static_cast<CtxBase*>(c)->forward_flip(idx, out)) }
int ref_get_htilda(void * c, cdouble * out) { REF_TRY(static_cast<CtxBase*>(c)->get_htilda(out)) }
int ref_get_gradients(void * c, cdouble * out) { REF_TRY(static_cast<CtxBase*>(c)->get_gradients(out)) }
int ref_smatrix_set(void * c, const cdouble * O, double lambda, cdouble * aO, double * diag) { REF_TRY(static_cast<CtxBase*>(c)->smatrix_set(O, lambda, aO, diag)) }
int ref_smatrix_dot(void * c, const cdouble * v, cdouble * out) { REF_TRY(static_cast<CtxBase*>(c)->smatrix_dot(v, out)) }
int ref_smatrix_diag(void * c, double * out) { REF_TRY(static_cast<CtxBase*>(c)->smatrix_diag(out)) }
int ref_evolve(void * c, const cdouble * dx, double lr) { REF_TRY(static_cast<CtxBase*>(c)->evolve(dx, lr)) }
// lambdaOverride < 0: use the reference schedule.  stats = {e_re, e_im, rsd, lambda, cg_iters, finite}
int ref_sr_step(void * c, int nms, double lr, int maxIter, double tol, double lambdaOverride, double * stats6, cdouble * F, cdouble * dx)
{
  try
  {
    SRStatsC st;
    static_cast<CtxBase*>(c)->sr_step(nms, lr, maxIter, tol, lambdaOverride, &st, F, dx);
    stats6[0] = st.e_re; stats6[1] = st.e_im; stats6[2] = st.rsd; stats6[3] = st.lambda; stats6[4] = st.cg_iters; stats6[5] = st.finite;
    return 0;
  }
  catch (const std::exception & e) { std::cerr << "# ref_harness: " << e.what() << std::endl; return 1; }
}
void ref_set_threads(int omp_threads) { omp_set_num_threads(omp_threads); }
int ref_max_threads() { return omp_get_max_threads(); }
}
