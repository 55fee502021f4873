#!/usr/bin/env python
"""bench.py -- the VMC hot path on the workload BASELINE.json's metric is quoted on.

  python bench.py --gpus N --steps K --warmup W            # this repository's CUDA engine (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation (oracle/_ref), rank 0 only

One "step" = one iteration of the reference's StochasticReconfigurationCG::propagate (gpu/include/optimizer.cuh:127-165):
nms=1 Metropolis sweep (N proposals per chain) -> local energy -> O -> SR setup -> preconditioned CG to tol 1e-5 with the
reference's lambda schedule -> parameter update.  Workload = cfg3 of BASELINE.json: complex RBM alpha=2 (M=2N), long-range TFI
chain N=128 (alpha_LR=2, theta=pi/4, OBC), 16384 chains in total, sharded over the ranks (strong scaling; RNG keyed by global
chain id).  Parameters: the reference init law from a fixed numpy seed ("synthetic"), chains warmed up before timing.

value = VMC samples/s = (chains processed by all ranks per step) / (device time per step, max over ranks, CUDA events on the
engine's stream).  The O matrix alone is 8.7 GB (>> 126 MB L2), so no L2 flush is needed between steps.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    # name: (model, N, M, K_total)
    "cfg1": ("rbm", 16, 16, 512),
    "cfg2": ("rbm", 64, 128, 4096),
    "cfg3": ("rbm", 128, 256, 16384),
    "cfg4": ("ffnn", 128, 512, 8192),
    "cfg5": ("rbm", 256, 1024, 65536),     # 276 GB of O: explicit-O mode needs >= 2 GPUs (138 GB per rank), 8 GPUs -> 34.5 GB per rank
}
THETA_H = math.pi / 4
H_FIELD, J_COUP, ALPHA_LR = -math.cos(THETA_H), math.sin(THETA_H), 2.0
FP64_DMMA_PEAK_TFLOPS = 37.1   # mma.sync m8n8k4 f64, all 148 SMs, measured (scripts/micro/dmma_bench.cu); nominal 37-40
METRIC = "vmc_samples_per_s"
UNIT = "samples/s"


def synthetic_params(model: str, N: int, M: int, cfg_id: int) -> np.ndarray:
    """Reference init law (gpu/include/impl_neural_quantum_state.cuh:30-48 / :766-783) from numpy seed 20261018+cfg."""
    from neural_network_quantum_state_b200.init import reference_init
    return reference_init(model, N, M, np.random.default_rng(20261018 + cfg_id))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 3.0:     # first sample = NVML is initialised: nothing of it is timed
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def committed_traffic(kernel: str, cfg_name: str, world: int):
    """dram bytes (read + write) per pass over O of the dominant kernel from the committed ncu --set full capture of exactly this
    (config, number of GPUs) -- profiles/roofline_traffic.json, keyed "<cfg>/<world>gpu" -- else None: a capture of another
    shard size says nothing about this run."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get("%s/%dgpu" % (cfg_name, world), {}).get(kernel)
        except Exception:
            return None
    return None


def multi_gpu_parity(world: int, rank: int, local_rank: int) -> dict:
    """N > 1 only, untimed: the chain-sharded engine (NCCL SR sums + in-kernel NVLink exchange) must reproduce the single-GPU
    trajectory of the same chains (the RNG is keyed by the global chain id) and leave bit-identical parameters on every rank.
    The same check as tests/test_gpu_multi.py, run here because the driver's GPU test box has one GPU and the scaling run has N."""
    import torch.distributed as dist
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.dist import ShardPlan, bootstrap_comm, enable_p2p
    from neural_network_quantum_state_b200.init import reference_init
    model, N, M, K = "rbm", 24, 48, 1000          # 1000 chains: uneven shards for world = 3, 7, ...
    params = reference_init(model, N, M, np.random.default_rng(5))
    plan = ShardPlan(K, world, rank)
    e = Engine(model, N, M, h=H_FIELD, J=J_COUP, alpha=ALPHA_LR, seed=11, device=local_rank, **plan.engine_kwargs())
    e.set_params(params)
    bootstrap_comm(e, world, rank, p2p=False)
    p2p = enable_p2p(e, world)
    e.warm_up(30)
    traj = []
    for _ in range(5):
        st = e.sr_step(n_mc_steps=1, lr=0.05)
        traj.append((st.e_mean.real, st.rsd, st.cg_iters))
    final = e.get_params()
    variant = e.kernel_variant("sv")
    e.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, final.tobytes())
    out = None
    if rank == 0:
        identical = all(g == gathered[0] for g in gathered)
        e1 = Engine(model, N, M, K, H_FIELD, J_COUP, ALPHA_LR, seed=11, device=local_rank)
        e1.set_params(params)
        e1.warm_up(30)
        e_err, it_ok = 0.0, True
        for k in range(5):
            st = e1.sr_step(n_mc_steps=1, lr=0.05)
            e_err = max(e_err, abs(st.e_mean.real - traj[k][0]) / max(abs(st.e_mean.real), 1e-300))
            it_ok = it_ok and (st.cg_iters == traj[k][2])
        want = e1.get_params()
        e1.close()
        p_err = float(np.abs(final - want).max() / np.abs(want).max())
        out = {"ok": bool(identical and it_ok and e_err < 1e-10 and p_err < 1e-7), "ranks_bit_identical": bool(identical),
               "energy_max_rel_err_vs_1gpu": e_err, "params_max_rel_err_vs_1gpu": p_err, "cg_iters_equal": bool(it_ok),
               "in_kernel_exchange": bool(p2p), "sv_variant": variant,
               "case": "rbm N=24 M=48, 1000 chains sharded over %d ranks, 30 warm-up sweeps + 5 SR steps" % world}
    flag = [out]
    dist.broadcast_object_list(flag, src=0)
    if not flag[0]["ok"]:
        raise RuntimeError("multi-GPU parity check failed: %s" % json.dumps(flag[0]))
    return flag[0]


def gpu_reference_leg(model: str, N: int, M: int, K: int, cfg_id: int, device: int) -> dict:
    """The reference's OWN CUDA driver (gpu/src/LICH-train_rbm.cu, unmodified, compiled for sm_100 by baseline/Makefile) on this
    GPU, same parameter files, same uniforms (Philox TRNG shim), next to this engine: ms per SR step from the arrival times of the
    rows it prints (8 iterations after 100 warm-up sweeps, the first 3 left out) and the energy trajectories side by side.  A second leg next to the CPU reference arm,
    not a replacement for it."""
    import tempfile
    from baseline import ref_cuda
    from neural_network_quantum_state_b200 import Engine
    if model != "rbm":
        return {"unavailable": "the reference ships no LICH-train_ffnn driver (SURVEY 0.2)"}
    if not ref_cuda.available():
        return {"unavailable": "baseline/_ref/LICH-train_rbm-gpu-ref missing (make -C baseline in the development container)"}
    theta = float(ref_cuda.THETA_STR)
    h, J = -math.cos(theta), math.sin(theta)
    # 100 warm-up sweeps as in the main run: from the identical Neel start fewer leave zero-variance columns in O, whose 0/0 in
    # the preconditioner turns BOTH programs' parameters into NaN at the second iteration (seen with 10 sweeps)
    seed, nwarm, n_a, n_b = 20261018, 100, 3, 8
    with tempfile.TemporaryDirectory() as tmp:
        prefix = ref_cuda.prefix_for(tmp, N, M)
        e = Engine(model, N, M, K, h, J, ALPHA_LR, seed=seed, device=device)
        e.set_params(synthetic_params(model, N, M, cfg_id))
        e.save(prefix, 17)
        e.load(prefix)
        rb = ref_cuda.run(N, M, K, n_b, nwarm, seed, tmp, device=device)
        e.warm_up(nwarm)
        ours = [e.sr_step(n_mc_steps=1, lr=1e-2).e_mean.real for _ in range(n_b)]
        e.close()
    # the driver flushes one row per iteration: the time between the arrival of row n_a and row n_b spans iterations n_a+1..n_b
    # (the first n_a are left out as its warm-up: cuBLAS / Thrust first-use costs)
    t = rb["row_times_s"]
    per_iter = [(t[i] - t[i - 1]) * 1e3 for i in range(1, len(t))]
    if len(t) < n_b:
        return {"unavailable": "the reference driver printed %d of %d iteration rows: %s" % (len(t), n_b, rb["stdout_tail"][-200:])}
    ms = (t[n_b - 1] - t[n_a - 1]) / (n_b - n_a) * 1e3
    diff = max(abs(a - b) / max(abs(b), 1e-300) for a, b in zip(ours, rb["energies"])) if rb["energies"] else None
    return {"ms_per_step": ms, "value": K / (ms * 1e-3), "unit": UNIT, "steps_timed": n_b - n_a,
            "ms_per_iteration": [round(x, 3) for x in per_iter],
            "energies_reference": rb["energies"], "energies_engine": ours, "energy_max_rel_diff": diff,
            "how": "unmodified gpu/src/LICH-train_rbm.cu, nvcc -arch=sm_100, TRNG4 -> Philox shim (same uniforms as the engine); "
                   "one run of 8 iterations after 100 warm-up sweeps, each iteration timed by the arrival of the row the driver "
                   "prints and flushes at its end; ms_per_step = mean of iterations 4..8; energies printed with 7 digits"}


def sweep_roofline(model, N, M, K_loc, sweep_ms, hbm_peak_gbs, variant):
    """SURVEY 8d asks for BOTH bounds of the state-resident sweep and for a statement of which one binds.
    HBM: B_sw = theta in + out (2*K*M*16) + spins in + out (2*K*N) + lnpsi0/sa in + out (4*K*16) + the flip tables once (L2-resident
    afterwards).  fp64: K*N*M flip factors per sweep at 6 fp64 instructions each in the product-form kernel against
    the measured scalar-DFMA issue rate (half of FP64_DMMA_PEAK_TFLOPS flop/s = thread-instructions/s)."""
    if sweep_ms <= 0:
        return None
    mpad = 32 * max(1, 1 << max(0, (M - 1).bit_length() - 5))
    b_sw = 2.0 * K_loc * M * 16 + 2.0 * K_loc * N + 4.0 * K_loc * 16 + 3.0 * N * mpad * 16
    factors = float(K_loc) * N * M
    fast = variant.startswith("rbm_regs") or variant.startswith("ffnn_")
    ipf = 6.0 if variant.startswith("rbm_regs") else None
    hbm = b_sw / (sweep_ms * 1e-3) / 1e9
    out = {"hbm": {"algorithmic_bytes": b_sw, "achieved_gbs": hbm, "peak_gbs": hbm_peak_gbs, "frac": hbm / hbm_peak_gbs},
           "factor_evals_per_s": factors / (sweep_ms * 1e-3)}
    peak_inst = FP64_DMMA_PEAK_TFLOPS * 1e12 / 2.0
    if ipf is not None:
        ach = factors * ipf / (sweep_ms * 1e-3)
        out["fp64"] = {"fp64_instr_per_factor": ipf, "achieved_tinstr_s": ach / 1e12, "peak_tinstr_s": peak_inst / 1e12,
                       "frac": ach / peak_inst,
                       "peak_source": "scalar DFMA issue rate = measured fp64 peak / 2 (scripts/micro/dmma_bench.cu)"}
    out["binds"] = ("fp64 issue: the chain state is resident on chip, so HBM sees only B_sw per sweep (fraction above, small by "
                    "construction) and the time goes to dependent fp64 instructions at %s; ncu: profiles/r2_sweep_*" %
                    ("8 warps per SM (255 registers per thread)" if fast else "one warp per chain with log cosh / exp / sincos per factor"))
    return out


# =====================================================================================================================
# reference arm / cpu baseline: the reference's own CPU implementation (oracle/_ref/libnqs_ref.so) on a bounded sample
# =====================================================================================================================
def run_reference_cpu(cfg_name: str, steps: int, warmup: int, k_sample: int, n_warm_sweeps: int):
    os.environ.setdefault("OPENBLAS_NUM_THREADS", str(os.cpu_count() or 1))
    from oracle import ref_cpu
    model, N, M, _ = CONFIGS[cfg_name]
    cfg_id = int(cfg_name[3:])
    params = synthetic_params(model, N, M, cfg_id)
    cores = os.cpu_count() or 1
    lib = ref_cpu.lib()

    def make(threads):
        lib.ref_set_threads(threads)
        r = ref_cpu.RefSampler(model, N, M, k_sample, H_FIELD, J_COUP, ALPHA_LR, False)
        r.set_params(params)
        rng = np.random.default_rng(99)
        U = rng.random(((n_warm_sweeps + warmup + steps + 1) * N, k_sample))
        r.set_uniforms(U)
        # random initial spins: with the reference's identical (Neel) start a 256-chain sample needs ~100 sweeps (minutes of CPU
        # time) before its S matrix stops being degenerate (zero-variance columns -> 0/0 in the preconditioner)
        r.set_initial_spins((2 * rng.integers(0, 2, size=(k_sample, N)) - 1).astype(np.float64))
        r.warm_up(n_warm_sweeps)
        return r

    # the reference's OpenMP loops do not always scale: probe one whole SR step (sweep + E_loc + O + CG)
    # over OpenMP threads {1, 2, 4, ..., cores} (BLAS on all cores), then over BLAS threads with the best OpenMP count
    grid = sorted({min(cores, 1 << q) for q in range(0, 12)} | {cores})
    blas_ctl = ref_cpu.set_blas_threads(cores)
    probes = []

    def probe(omp, blas):
        if blas_ctl:
            ref_cpu.set_blas_threads(blas)
        r = make(omp)
        t0 = time.perf_counter()
        r.sr_step(1, 1e-2)
        dt = time.perf_counter() - t0
        r.close()
        probes.append({"omp": omp, "blas": blas if blas_ctl else int(os.environ.get("OPENBLAS_NUM_THREADS", "1")), "step_s": dt})
        return dt

    best_threads, best_t = 1, None
    for th in grid:
        dt = probe(th, cores)
        if best_t is None or dt < best_t:
            best_threads, best_t = th, dt
    blas_threads = cores if blas_ctl else int(os.environ.get("OPENBLAS_NUM_THREADS", "1"))
    if blas_ctl:
        for bt in grid[:-1]:
            dt = probe(best_threads, bt)
            if dt < best_t:
                blas_threads, best_t = bt, dt
        ref_cpu.set_blas_threads(blas_threads)
    r = make(best_threads)
    cg = []
    for _ in range(warmup):
        r.sr_step(1, 1e-2)
    t0 = time.perf_counter()
    for _ in range(steps):
        cg.append(r.sr_step(1, 1e-2)["cg_iters"])
    dt = (time.perf_counter() - t0) / max(steps, 1)
    r.close()
    return {"value": k_sample / dt, "ms_per_step": dt * 1e3, "cores": max(best_threads, blas_threads), "host_cores": cores,
            "omp_threads": best_threads, "blas_threads": blas_threads, "cg_iters": cg, "step_s_probe": best_t,
            "thread_scan": probes, "k_sample": k_sample,
            "sample": "%d of %d chains (random initial spins), same N=%d M=%d, %d warm-up sweeps, %d warm-up + %d timed SR steps; reference CPU "
                      "headers + OpenBLAS 0.3.15 (MKL/TRNG4 unavailable offline), long-range Hamiltonian shim" %
                      (k_sample, CONFIGS[cfg_name][3], N, M, n_warm_sweeps, warmup, steps)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--k-total", type=int, default=0, help="diagnostics: override the number of chains of the configuration")
    ap.add_argument("--nwarm", type=int, default=100, help="warm-up sweeps before the benchmark (reference default -nwarm)")
    ap.add_argument("--lr", type=float, default=1e-2)
    ap.add_argument("--cpu-sample-chains", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--force-generic", action="store_true")
    ap.add_argument("--cg-fixed-iters", type=int, default=0, help="diagnostics: run exactly this many CG iterations per step")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: ncclAllReduce per CG iteration instead of the in-kernel exchange")
    ap.add_argument("--two-pass-sv", action="store_true", help="S*v as two streaming passes over O (reference structure)")
    ap.add_argument("--no-structured-extra", action="store_true",
                    help="skip the second, separately reported measurement of the same step with NQS_FLAG_STRUCTURED_SV (1 GPU only)")
    ap.add_argument("--no-gpu-reference", action="store_true", help="skip the same-box run of the reference's own CUDA driver (1 GPU, RBM)")
    ap.add_argument("--no-parity-check", action="store_true", help="multi-GPU: skip the (untimed) sharded-vs-single-GPU parity check")
    ap.add_argument("--production-iters", type=int, default=100,
                    help="extra block: steps at the lambda floor 1e-2 with this many CG iterations (the regime of a converged run, "
                         "SURVEY 9.2); 0 = skip")
    ap.add_argument("--production-steps", type=int, default=3)
    ap.add_argument("--structured-sv", action="store_true",
                    help="S*v from the factors of O as two fp64 tensor-core GEMMs (no O matrix); roofline is then the fp64 tensor pipe")
    args = ap.parse_args()
    assert args.warmup >= 0 and args.steps >= 1
    # the contract is ONE JSON line on stdout: libraries that print there (NCCL's version banner, torchrun notices) are sent
    # to stderr for the duration of the run and the line is written to the saved descriptor at the end
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    model, N, M, K_total = CONFIGS[args.config]
    if args.k_total > 0:
        K_total = args.k_total
    cfg_id = int(args.config[3:])
    P = N * M + N + M if model == "rbm" else N * M + 2 * M
    config = {"workload": "%s: complex %s M=%d, long-range TFI chain N=%d (alpha=2, theta=pi/4, OBC), %d chains total, "
                          "step = sweep(nms=1) + E_loc + O + SR setup + PCG(tol 1e-5, reference lambda schedule) + update"
                          % (args.config, model.upper(), M, N, K_total),
              "N": N, "M": M, "K_total": K_total, "P": P, "sharding": "chains/%d" % world, "cg_exchange": None,
              "l2": "working set (O = %.2f GB per rank) >> 126 MB L2: no flush needed" % (K_total / world * P * 16 / 1e9),
              "nwarm_sweeps": args.nwarm, "lr": args.lr, "rng": "in-kernel Philox4x32-10 (value) / pre-drawn host uniforms (e2e)"}

    if args.structured_sv:
        config["l2"] = ("structured S*v keeps no O: per step theta + T (%.0f MB each per rank) are rewritten and re-read and the "
                        "GEMM partials (~20 MB) rewritten per product; together > 126 MB L2, no flush" % (K_total / world * M * 16 / 1e6))
        config["workload"] += "; S*v from the factors of O (NQS_FLAG_STRUCTURED_SV)"
    # ------------------------------------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import ref_cpu
        if not ref_cpu.available():
            emit({"impl": "reference", "unavailable": "oracle/_ref/libnqs_ref.so missing (run make -C oracle)"})
            return 0
        res = run_reference_cpu(args.config, args.steps, args.warmup, args.cpu_sample_chains, n_warm_sweeps=5)
        # this arm runs a bounded SAMPLE of the workload: say so in its own config (chains actually run, start, warm-up)
        config = dict(config)
        config.update({"K_total": res["k_sample"], "K_total_workload": K_total, "sharding": "host cores (rank 0 only)",
                       "cg_exchange": "none (CPU)", "nwarm_sweeps": 5, "initial_spins": "random (not Neel)",
                       "rng": "pre-drawn numpy uniforms (TRNG4 shim)",
                       "l2": "n/a (CPU)",
                       "workload": config["workload"].replace("%d chains total" % K_total,
                                                              "%d-chain sample of the %d-chain workload" % (res["k_sample"], K_total))})
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "reference",
                                 "sample": res["sample"], "host_cores": res["host_cores"], "omp_threads": res["omp_threads"],
                                 "blas_threads": res["blas_threads"], "thread_scan": res["thread_scan"]},
                "cg_iters_per_step": res["cg_iters"],
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # --------------------------------------------------------------------------------------------------- native arm
    import torch
    import torch.distributed as dist
    from neural_network_quantum_state_b200 import Engine

    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the engine has no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    assert K_total % world == 0
    K_loc = K_total // world
    if not args.structured_sv:
        need = K_loc * P * 16 + 4e9
        assert need < torch.cuda.mem_get_info()[0], \
            "%s needs %.0f GB of HBM per rank for O: use more GPUs or --structured-sv" % (args.config, need / 1e9)
    e = Engine(model, N, M, K_loc, H_FIELD, J_COUP, ALPHA_LR, pbc=False, seed=20261018, device=local_rank,
               n_chains_total=K_total, chain_offset=rank * K_loc, max_predrawn_steps=N, force_generic=args.force_generic,
               two_pass_sv=args.two_pass_sv, structured_sv=args.structured_sv)
    e.set_params(synthetic_params(model, N, M, cfg_id))
    p2p = False
    if world > 1:
        from neural_network_quantum_state_b200.dist import bootstrap_comm, enable_p2p
        bootstrap_comm(e, world, rank, p2p=False)          # NCCL communicator (SR-setup all-reduce)
        if not args.no_p2p:
            p2p = enable_p2p(e, world)                      # per-CG-iteration exchange inside the CG kernel over NVLink
    parity = multi_gpu_parity(world, rank, local_rank) if (world > 1 and not args.no_parity_check) else None
    e.warm_up(args.nwarm)

    def barrier():
        e.sync()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return e.sr_step(n_mc_steps=1, lr=args.lr, fixed_iters=args.cg_fixed_iters)

    def restart():
        """Same optimisation iterations in every measured loop (lambda decays from step to step and the CG iteration count grows
        with it): parameters, lambda schedule and CG warm start reset, chains re-warmed, W untimed steps."""
        e.set_params(synthetic_params(model, N, M, cfg_id))
        e.sr_reset()
        e.warm_up(args.nwarm)
        for _ in range(args.warmup):
            step()

    # ---- (1) the timed region: K steps, no per-kernel instrumentation, CUDA events on the engine's stream around the loop
    # nvidia-smi is started BEFORE the warm-up steps: its NVML initialisation (enumerates every GPU of the box) can stall CUDA
    # calls of the running processes for tens of ms, which must not land inside the timed region; it then samples every 100 ms
    clocks = ClockSampler(local_rank)
    clocks.start()
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = e.get_timing()["kernel_launches"]
    cg_iters, energies = [], []
    e.event_record(0)
    for _ in range(args.steps):
        st = step()
        cg_iters.append(st.cg_iters)
        energies.append(st.e_mean.real)
    e.event_record(1)
    barrier()
    ms_total = e.event_elapsed_ms(0, 1)
    clock_info = clocks.stop()
    launches = e.get_timing()["kernel_launches"] - launches0
    if world > 1:
        tt = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    value = K_total / (ms_per_step * 1e-3)
    config["cg_exchange"] = "none (1 GPU)" if world == 1 else ("in-kernel all-reduce over NVLink peer memory" if p2p else "ncclAllReduce")

    # ---- (2) the same K steps again with a CUDA-event pair around every phase and every S*v launch (phase times, roofline).
    # The ~40 extra event records per step cost ~0.15 ms (they break back-to-back launches), which is why they stay out of (1).
    restart()
    e.set_timing(True)
    barrier()
    phase = {k: 0.0 for k in ("sweep_ms", "eloc_ms", "oderiv_ms", "setup_ms", "cg_ms", "update_ms", "rows_ms", "cols_ms")}
    counts = {"rows_count": 0, "cols_count": 0}
    cg_iters_inst = []
    e.event_record(2)
    for _ in range(args.steps):
        st = step()
        cg_iters_inst.append(st.cg_iters)
        t = e.get_timing()
        for k in phase:
            phase[k] += t[k]
        for k in counts:
            counts[k] += t[k]
    e.event_record(3)
    barrier()
    ms_per_step_inst = e.event_elapsed_ms(2, 3) / args.steps
    e.set_timing(False)

    # ---- production regime: lambda has decayed to its floor 1e-2 (schedule step >= 88) and the solve needs ~100 products per
    # step -- the headline above measures schedule steps W+1 .. W+K, where lambda is still 53 -> 7 and 6-11 suffice
    production = None
    if args.production_iters > 0:
        restart()
        barrier()
        fin = []
        e.event_record(4)
        for _ in range(args.production_steps):
            st = e.sr_step(n_mc_steps=1, lr=args.lr, lam=1e-2, fixed_iters=args.production_iters)
            fin.append(bool(st.finite))
        e.event_record(5)
        barrier()
        pms = e.event_elapsed_ms(4, 5) / args.production_steps
        if world > 1:
            tt = torch.tensor([pms], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            pms = float(tt.item())
        nprod = args.production_iters + 1
        production = {"value": K_total / (pms * 1e-3), "unit": UNIT, "ms_per_step": pms, "steps": args.production_steps,
                      "lambda": 1e-2, "cg_iterations": args.production_iters, "finite": fin,
                      "o_passes_per_step": nprod,
                      "hbm_gbs_whole_step": nprod * K_loc * P * 16.0 / (pms * 1e-3) / 1e9,
                      "note": "same workload, lambda fixed at the schedule's floor and the solve run for a fixed %d iterations; "
                              "hbm_gbs_whole_step charges the WHOLE step (sweep, E_loc, O writer, setup, update included) to the "
                              "algorithmic bytes of the passes over O" % args.production_iters}

    # ---- end-to-end through the public API with HOST buffers: pinned uniforms H2D every step, spins + lnpsi + stats D2H
    e2e = None
    if not args.no_e2e:
        rng = np.random.default_rng(1234 + rank)
        n_e2e = args.steps                  # the SAME optimisation iterations as the device-timed loop (same lambda / CG counts)
        # this step's inputs wait in PINNED host memory (one buffer per step, filled before the clock starts, as a producer
        # thread would); results are read back into host numpy arrays
        u_bufs = [torch.empty((N, K_loc), dtype=torch.float64, pin_memory=True) for _ in range(n_e2e)]
        for ub in u_bufs:
            ub.numpy()[...] = rng.random((N, K_loc))
        # page-locked result buffers (a user who reads the samples back every step keeps them around)
        spins_out = torch.empty((K_loc, N), dtype=torch.int8, pin_memory=True).numpy()
        lnpsi_out = torch.empty((K_loc,), dtype=torch.complex128, pin_memory=True).numpy()
        restart()
        e2e_iters = []
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e.set_uniforms(u_bufs[i].numpy())   # pinned: read in place over PCIe by the sweep (inside the timed region)
            st = step()
            e2e_iters.append(st.cg_iters)
            spins = e.get_spinStates(out=spins_out)          # D2H: what a pynqs user reads back
            lnpsi = e.get_lnpsi(out=lnpsi_out)
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e.set_uniforms(None)
        e2e = {"value": K_total / dt, "unit": UNIT, "h2d_bytes_per_step": int(u_bufs[0].numpy().nbytes) * world,
               "d2h_bytes_per_step": int(spins.nbytes + lnpsi.nbytes + 56) * world, "ms_per_step": dt * 1e3, "steps": n_e2e,
               "cg_iters_per_step": e2e_iters,
               "api": "Engine.set_uniforms + Engine.sr_step + get_spinStates + get_lnpsi (C ABI, host buffers)"}

    if rank != 0:
        e.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: the pass(es) over O inside every CG iteration (HBM-bound)
    peak, peak_src = measured_peak_gbs()
    bytes_per_launch = K_loc * P * 16.0     # one read of this rank's O [K_loc][P] complex fp64
    sv_variant = e.kernel_variant("sv")
    if sv_variant.startswith("structured"):
        dom = other = None
    elif sv_variant.startswith("fused"):
        # one-pass cluster kernel: O is read from HBM once per S*v (the reference and the two-pass kernels read it twice).
        # Persistent CG: ONE launch per solve runs all products (passes over O) of the solve AND the vector updates between them,
        # so the launch duration is charged with the whole solve; rows_count then counts products, not launches.
        persistent = sv_variant.endswith("_persistentcg")
        dom = "cg_persist_kernel" if persistent else "sv_fused_kernel"
        dom_ms, n_timed = phase["rows_ms"] / max(counts["rows_count"], 1), counts["rows_count"]
        other = None
    else:
        dom = "matvec_cols_partial_kernel" if phase["cols_ms"] >= phase["rows_ms"] else "matvec_rows_kernel"
        is_cols = dom.startswith("matvec_cols")
        dom_ms = (phase["cols_ms"] / max(counts["cols_count"], 1)) if is_cols else (phase["rows_ms"] / max(counts["rows_count"], 1))
        n_timed = counts["cols_count"] if is_cols else counts["rows_count"]
        other = {"kernel": "matvec_rows_kernel" if is_cols else "matvec_cols_partial_kernel",
                 "avg_launch_ms": (phase["rows_ms"] / max(counts["rows_count"], 1)) if is_cols
                 else (phase["cols_ms"] / max(counts["cols_count"], 1))}
    if dom is None:
        # structured S*v: the two GEMMs (rows: z = O v, cols: O^H z) are fp64 tensor-core work, 2*K_loc*N*2M flops each
        flops = 2.0 * K_loc * N * 2 * M
        r_ms = phase["rows_ms"] / max(counts["rows_count"], 1)
        c_ms = phase["cols_ms"] / max(counts["cols_count"], 1)
        dom, dom_ms, n_timed = ("spin_cols_dmma_kernel", c_ms, counts["cols_count"]) if c_ms >= r_ms else \
                               ("spin_rows_dmma_kernel", r_ms, counts["rows_count"])
        ach = flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": dom, "variant": sv_variant, "achieved": ach, "peak": FP64_DMMA_PEAK_TFLOPS,
                    "unit": "TFLOP/s", "frac": ach / FP64_DMMA_PEAK_TFLOPS,
                    "peak_source": "fp64 DMMA peak measured on this pool's B200 with scripts/micro/dmma_bench.cu (MEASURED_PEAKS.json "
                                   "holds only the bf16 figure)",
                    "traffic": committed_traffic(dom, args.config, world), "algorithmic_flops_per_launch": flops, "avg_launch_ms": dom_ms,
                    "launches_timed": n_timed,
                    "other_pass": {"kernel": "spin_rows_dmma_kernel" if dom.startswith("spin_cols") else "spin_cols_dmma_kernel",
                                   "avg_launch_ms": r_ms if dom.startswith("spin_cols") else c_ms},
                    "note": "S*v = two real-by-complex GEMMs with the +-1 spin matrix (mma.sync m8n8k4 f64); no pass over O"}
        other = None
        achieved = None
    else:
        achieved = bytes_per_launch / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    if achieved is not None:
        roofline = {"bound": "hbm", "kernel": dom, "variant": sv_variant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                  "frac": achieved / peak, "peak_source": peak_src, "traffic": committed_traffic(dom, args.config, world),
                  "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_ms": dom_ms, "launches_timed": n_timed,
                  "note": "achieved = K_loc*P*16 B (O read once) / CUDA-event duration of the launch; SURVEY 8d's two-pass figure "
                          "B_cg = 2*K_loc*P*16 per S*v is met with %s pass(es) over O" % ("1" if other is None else "2")}
        if dom == "cg_persist_kernel":
            prods = counts["rows_count"] / float(args.steps)
            roofline.update({
                "algorithmic_bytes_per_launch": bytes_per_launch * prods, "avg_launch_ms": phase["rows_ms"] / args.steps,
                "launches_timed": args.steps, "products_per_launch": prods, "ms_per_product": dom_ms,
                "traffic": (lambda t: None if t is None else t * prods)(committed_traffic(dom, args.config, world)),
                "note": "one launch = one whole CG solve: (iterations + 1) passes over O of K_loc*P*16 B each (O read once per pass) "
                        "plus the vector updates, grid barriers and (multi-GPU) NVLink exchanges between them; achieved = bytes of "
                        "all passes / CUDA-event duration of the launch, i.e. the non-streaming phases are charged to the kernel"})
    if other is not None:
        roofline["other_pass"] = other
    sweep_ms = phase["sweep_ms"] / args.steps
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "clocks": clock_info, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "production_regime": production, "multi_gpu_parity": parity,
            "sr_step_ms": ms_per_step - sweep_ms,
            "sweep": {"ms": sweep_ms, "samples_per_s": K_total / (sweep_ms * 1e-3) if sweep_ms > 0 else None,
                      "proposals_per_s": K_total * N / (sweep_ms * 1e-3) if sweep_ms > 0 else None,
                      "kernel": e.kernel_variant("sweep"),
                      "roofline": sweep_roofline(model, N, M, K_loc, sweep_ms, peak, e.kernel_variant("sweep"))},
            "phase_ms_per_step": {k: v / args.steps for k, v in phase.items()},
            "cg_iters_per_step": cg_iters,
            "cg_ms_per_iter": (phase["cg_ms"] / max(sum(cg_iters_inst) + len(cg_iters_inst), 1)),
            "instrumented_pass": {"ms_per_step": ms_per_step_inst, "cg_iters_per_step": cg_iters_inst,
                                  "note": "phase_ms_per_step, sweep and roofline durations come from a second pass over the same K steps "
                                          "with CUDA events around every phase and S*v launch; value / ms_per_step from the first, "
                                          "uninstrumented pass"},
            "energy_per_site": energies}
    if world == 1 and not args.structured_sv and not args.no_structured_extra:
        # the same step with S*v formed from the factors of O on the fp64 tensor cores (no O matrix): reported NEXT TO the
        # headline, which stays on the explicit-O formulation the reference and the north star are stated on
        try:
            e2 = Engine(model, N, M, K_loc, H_FIELD, J_COUP, ALPHA_LR, pbc=False, seed=20261018, device=local_rank,
                        n_chains_total=K_total, chain_offset=0, structured_sv=True)
            ph2 = {k: 0.0 for k in phase}
            cnt2 = {"rows_count": 0, "cols_count": 0}
            it2, en2 = [], []
            ms2 = None
            for instrumented in (False, True):      # as above: headline pass without per-kernel events, then the instrumented one
                e2.set_params(synthetic_params(model, N, M, cfg_id))
                e2.sr_reset()
                e2.warm_up(args.nwarm)
                for _ in range(args.warmup):
                    e2.sr_step(n_mc_steps=1, lr=args.lr, fixed_iters=args.cg_fixed_iters)
                e2.set_timing(instrumented)
                e2.sync()
                e2.event_record(0)
                for _ in range(args.steps):
                    st = e2.sr_step(n_mc_steps=1, lr=args.lr, fixed_iters=args.cg_fixed_iters)
                    if instrumented:
                        t = e2.get_timing()
                        for k in ph2:
                            ph2[k] += t[k]
                        for k in cnt2:
                            cnt2[k] += t[k]
                    else:
                        it2.append(st.cg_iters)
                        en2.append(st.e_mean.real)
                e2.event_record(1)
                e2.sync()
                if not instrumented:
                    ms2 = e2.event_elapsed_ms(0, 1) / args.steps
            e2.set_timing(False)
            flops = 2.0 * K_loc * N * 2 * M
            r_ms = ph2["rows_ms"] / max(cnt2["rows_count"], 1)
            c_ms = ph2["cols_ms"] / max(cnt2["cols_count"], 1)
            line["structured_sv"] = {
                "value": K_total / (ms2 * 1e-3), "unit": UNIT, "ms_per_step": ms2, "variant": e2.kernel_variant("sv"),
                "phase_ms_per_step": {k: v / args.steps for k, v in ph2.items()}, "cg_iters_per_step": it2,
                "energy_per_site": en2,
                "gemm": {"flops_per_launch": flops, "rows_ms": r_ms, "cols_ms": c_ms,
                         "rows_tflops": flops / (r_ms * 1e-3) / 1e12 if r_ms > 0 else None,
                         "cols_tflops": flops / (c_ms * 1e-3) / 1e12 if c_ms > 0 else None,
                         "fp64_dmma_peak_tflops": FP64_DMMA_PEAK_TFLOPS,
                         "tensor_path": "tcgen05 kind::i8 (7 int8 digit planes per fp64 operand: 7x the listed flops as int8 ops)"
                                        if "umma" in e2.kernel_variant("sv") else "mma.sync m8n8k4 f64 (DMMA)",
                         "tflops_note": "fp64-EQUIVALENT rate = flops of the fp64 GEMM / time; on the int8 path it may exceed the DMMA peak"},
                "note": "opt-in NQS_FLAG_STRUCTURED_SV: O^H(O v) as two real-by-complex GEMMs on the factors of O (tcgen05 int8 UMMA with an "
                        "error-free digit split from 8192 chains per rank, fp64 DMMA below); O is never written; same energies as the "
                        "headline run to rounding"}
            e2.close()
        except Exception as ex:
            line["structured_sv"] = {"value": None, "error": str(ex)}
    if world == 1 and not args.no_gpu_reference and not args.structured_sv:
        # the engine's O (and the reference's two copies of it) must fit together: release ours first
        e.close()
        try:
            line["gpu_reference"] = gpu_reference_leg(model, N, M, K_total, cfg_id, local_rank)
        except Exception as ex:
            line["gpu_reference"] = {"unavailable": "failed: %s" % str(ex)[-300:]}
    if not args.no_cpu_baseline and world == 1:
        try:
            res = run_reference_cpu(args.config, steps=2, warmup=1, k_sample=args.cpu_sample_chains, n_warm_sweeps=5)
            line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "reference",
                                    "sample": res["sample"], "host_cores": res["host_cores"], "omp_threads": res["omp_threads"],
                                    "blas_threads": res["blas_threads"], "ms_per_step": res["ms_per_step"],
                                    "cg_iters_per_step": res["cg_iters"], "thread_scan": res["thread_scan"]}
        except Exception as ex:  # the checker being absent must not hide the GPU number
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % ex}
    emit(line)
    e.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
