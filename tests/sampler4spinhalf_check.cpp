// Check program of tests/test_cli_host.py::test_cpp_sampler4spinhalf: the reference's measurement-side calling sequence
// (gpu/src/meas_*_rbmtrsymm.cu: ansatz.load(path); Sampler4SpinHalf smp(psi, seed, dist); smp.warm_up(n); smp.do_mcmc_steps(m);
// smp.get_lnpsi(); smp.get_quantumStates(); psi.forward(spins, lnpsi, false)) through host/nqs_host.hpp.
// usage: sampler4spinhalf_check <variables file> N alpha K   -> per chain: spins, tracked lnpsi, forward(spins) of a second instance
#include <cstdio>
#include <cstdlib>
#include "../neural_network_quantum_state_b200/host/nqs_host.hpp"

int main(int argc, char ** argv)
{
  if (argc != 5) return 2;
  const int N = std::atoi(argv[2]), alpha = std::atoi(argv[3]), K = std::atoi(argv[4]);
  try
  {
    struct Traits { using AnsatzType = spinhalf::RBMTrSymm<double>; using FloatType = double; };
    spinhalf::RBMTrSymm<double> psi(N, alpha, K), psi1(N, alpha, K);
    psi.load(argv[1]);
    psi.copy_to(psi1);
    Sampler4SpinHalf<Traits> smp(psi, 3ul, 1000ul);
    smp.warm_up(6);
    smp.do_mcmc_steps(3);
    const auto ln = smp.get_lnpsi();
    const auto s = smp.get_quantumStates();
    std::vector<std::complex<double> > fixed((size_t)K);
    psi1.forward(s.data(), fixed.data(), false);
    for (int k = 0; k < K; ++k)
    {
      for (int i = 0; i < N; ++i) std::printf("%d ", (int)s[(size_t)k*N+i]);
      std::printf("%.17g %.17g %.17g %.17g\n", ln[k].real(), ln[k].imag(), fixed[k].real(), fixed[k].imag());
    }
  }
  catch (const std::exception & e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
  return 0;
}
