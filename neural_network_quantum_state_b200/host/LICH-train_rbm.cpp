// LICH-train_rbm-gpu: trains a complex RBM on the long-range transverse-field Ising chain by stochastic reconfiguration.
// Drop-in for the reference program of the same name (ref gpu/src/LICH-train_rbm.cu:14-119): same options and defaults,
// same parameter-file prefix  <path>/RBMLICH-L{L}NH{nh}A{alpha}T{theta}V{ver}D{w,a,b}.dat, same stdout table.
// The hot path runs in libnqs_b200.so (hand-written sm_100a kernels) through the classes of nqs_host.hpp.
// -DNQS_DRIVER_RBMTRSYMM builds LICH-train_rbmtrsymm-gpu (ref gpu/src/LICH-train_rbmtrsymm.cu: the translation-symmetric RBM on
// the periodic chain, the CMake default target of the reference); -DNQS_DRIVER_RBMZ2PRSYMM / -DNQS_DRIVER_FFNNTRSYMM build the drivers of
// the other two tied ansaetze (ref gpu/src/LICH-train_rbmz2prsymm.cu, gpu/src/LICH-train_ffnntrsymm.cu).  -DNQS_DRIVER_FFNN builds LICH-train_ffnn-gpu: the same driver for the one-hidden-layer FNN (the reference ships the
// ansatz, gpu/src/CH-train_ffnn.cu:75, but no LICH driver for it; prefix FFNNLICH-..., files Dw1/Dw2/Db1).
#include <chrono>
#include <cmath>
#include <iostream>
#include <string>
#include "argsparse.hpp"
#include "nqs_host.hpp"

using namespace spinhalf;
using nqs_host::argsparse;
using nqs_host::pair_t;

template <typename FloatType>
static std::string remove_zeros_in_str(const FloatType val)
{ // "2.000000" -> "2", "0.785398" -> "0.785398"
  std::string s = std::to_string(val);
  s.erase(s.find_last_not_of('0')+1, std::string::npos);
  s.erase(s.find_last_not_of('.')+1, std::string::npos);
  return s;
}

int main(int argc, char * argv[])
{
#if defined(NQS_DRIVER_RBMTRSYMM)
  // ref gpu/src/LICH-train_rbmtrsymm.cu:14-110: -nf filters instead of -nh, periodic chain, nwarm 500 / rsd 1e-3 by default,
  // ONE variables file <path>/RBMTrSymmLICH-L{L}NF{nf}A{alpha}T{theta}V{ver}
  using Machine = RBMTrSymm<double>;
  const std::string tag = "RBMTrSymmLICH-L", what = "RBMTrSymm";
#elif defined(NQS_DRIVER_RBMZ2PRSYMM)
  // ref gpu/src/LICH-train_rbmz2prsymm.cu:14-113: -nf filters, OPEN chain, nwarm 100 / rsd 1e-3 by default, ONE variables file
  // <path>/RBMZ2PrSymmLICH-L{L}NF{nf}A{alpha}T{theta}V{ver}  (its help text says "RBMTrSymm", :21)
  using Machine = RBMZ2PrSymm<double>;
  const std::string tag = "RBMZ2PrSymmLICH-L", what = "RBMTrSymm";
#elif defined(NQS_DRIVER_FFNNTRSYMM)
  // ref gpu/src/LICH-train_ffnntrsymm.cu:14-113: -nf filters, periodic chain, nwarm 100 / rsd 1e-3 by default, ONE variables file
  using Machine = FFNNTrSymm<double>;
  const std::string tag = "FFNNTrSymmLICH-L", what = "FFNNTrSymm";
#elif defined(NQS_DRIVER_FFNN)
  using Machine = FFNN<double>;
  const std::string tag = "FFNNLICH-L", what = "FFNN";
#else
  using Machine = RBM<double>;
  const std::string tag = "RBMLICH-L", what = "RBM";
#endif
  const std::vector<pair_t> options = {
#if defined(NQS_DRIVER_RBMTRSYMM) || defined(NQS_DRIVER_RBMZ2PRSYMM) || defined(NQS_DRIVER_FFNNTRSYMM)
    {"L", "# of lattice sites"}, {"nf", "# of filters"}, {"ns", "# of spin samples for parallel Monte-Carlo"},
#else
    {"L", "# of lattice sites"}, {"nh", "# of hidden nodes"}, {"ns", "# of spin samples for parallel Monte-Carlo"},
#endif
    {"niter", "# of iterations to train "+what}, {"alpha", "exponent in the two-body interaction: J_{i,j} ~ 1/|i-j|^{alpha}"},
    {"theta", "J = sin(theta), h = -cos(theta)"}, {"ver", "version"}, {"nwarm", "# of MCMC steps for warming-up"},
    {"nms", "# of MCMC steps for sampling spins"}, {"dev", "device number"}, {"lr", "learning_rate"},
    {"rsd", "cutoff value of the energy deviation per energy (convergence criterion)"},
    {"path", "directory to load and save files"}, {"seed", "seed of the parallel random number generator"},
    {"ifprefix", "prefix of the file to load data"}};
#if defined(NQS_DRIVER_RBMTRSYMM) || defined(NQS_DRIVER_RBMZ2PRSYMM) || defined(NQS_DRIVER_FFNNTRSYMM)
  const std::vector<pair_t> defaults = {
#if defined(NQS_DRIVER_RBMTRSYMM)
    {"nwarm", "500"},
#else
    {"nwarm", "100"},
#endif
    {"nms", "1"}, {"lr", "1e-2"}, {"rsd", "1e-3"}, {"path", "."}, {"seed", "0"}, {"ifprefix", "None"}};
  const char * width_opt = "nf";
  const std::string width_tag = "NF";
#if defined(NQS_DRIVER_RBMZ2PRSYMM)
  const bool isPBC = false;
#else
  const bool isPBC = true;
#endif
#else
  const std::vector<pair_t> defaults = {
    {"nwarm", "100"}, {"nms", "1"}, {"lr", "1e-2"}, {"path", "."}, {"seed", "0"}, {"ifprefix", "None"}};
  const char * width_opt = "nh";
  const std::string width_tag = "NH";
  const bool isPBC = false;
#endif
  argsparse parser(argc, argv, options, defaults);

  const int L = parser.find<int>("L"), nChains = parser.find<int>("ns"), nWarmup = parser.find<int>("nwarm"),
    nMonteCarloSteps = parser.find<int>("nms"), deviceNumber = parser.find<int>("dev"), nIterations = parser.find<int>("niter");
  const double lr = parser.find<double>("lr"), RSDcutoff = parser.find<double>("rsd");
  const unsigned long long seed = parser.find<unsigned long long>("seed");
  const std::string path = parser.find<>("path")+"/", Lstr = parser.find<>("L"), ifprefix = parser.find<>("ifprefix");
  const auto nhArr = parser.mfind<int>(width_opt);
  const auto alphaArr = parser.mfind<double>("alpha");
  const auto verArr = parser.mfind<int>("ver");
  const auto thetaArr = parser.mfind<double>("theta");
  parser.print(std::cout);

  nqs_host::set_device(deviceNumber);
  struct SamplerTraits { using AnsatzType = Machine; using FloatType = double; };
  const unsigned long nBlocks = (unsigned long)nIterations*(unsigned long)nMonteCarloSteps*(unsigned long)L*(unsigned long)nChains;

  try
  {
    for (const auto & ver : verArr)
      for (const auto & nh : nhArr)
        for (const auto & alpha : alphaArr)
          for (const auto & theta : thetaArr)
          {
            Machine machine(L, nh, nChains);
            const double J = std::sin(theta), h = -std::cos(theta);
            const std::string prefix = path+tag+Lstr+width_tag+std::to_string(nh)+"A"+remove_zeros_in_str(alpha)+"T"+
              remove_zeros_in_str(theta)+"V"+std::to_string(ver);
            const std::string prefix0 = (ifprefix.compare("None")) ? path+ifprefix : prefix;
            machine.load(prefix0);
            LITFIChain<SamplerTraits> sampler(machine, L, h, J, alpha, isPBC, seed, nBlocks, prefix);
            const auto start = std::chrono::system_clock::now();
            sampler.warm_up(nWarmup);
            StochasticReconfigurationCG<double> iTimePropagator(nChains, machine.get_nVariables());
            iTimePropagator.propagate(sampler, nIterations, nMonteCarloSteps, lr, RSDcutoff);
            machine.save(prefix);
            const std::chrono::duration<double> elapsed_seconds = std::chrono::system_clock::now()-start;
            std::cout << "# elapsed time: " << elapsed_seconds.count() << "(sec)" << std::endl;
          }
  }
  catch (const nqs_host::Error & e)
  { // the reference prints "# ERROR --- FILE:.., LINE:.." and exits 1 on any CUDA failure (gpu/include/common.cuh:12-17)
    if (e.status == NQS_ERR_INVALID && std::string(e.what()).find("dev >= nDevice") != std::string::npos)
      std::cerr << "# error ---> dev(" << deviceNumber << ") >= # of devices" << std::endl;
    else
      std::cerr << "# ERROR --- " << e.what() << std::endl;
    return 1;
  }
  return 0;
}
