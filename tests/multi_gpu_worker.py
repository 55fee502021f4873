"""Worker of tests/test_gpu_multi.py (run under torchrun, one rank per GPU): the chain-sharded engine with the in-kernel NVLink
exchange (or ncclAllReduce with --no-p2p) must reproduce the single-GPU trajectory -- the RNG is keyed by the global chain id."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from neural_network_quantum_state_b200 import Engine
    from neural_network_quantum_state_b200.dist import ShardPlan, bootstrap_comm, enable_p2p
    from neural_network_quantum_state_b200.init import reference_init
    use_p2p = "--no-p2p" not in sys.argv
    structured = "--structured" in sys.argv
    out_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model, N, M, K = "rbm", 24, 48, 1000          # 1000 chains: uneven shards for world = 3, 7, ...
    h, J = -math.cos(math.pi / 4), math.sin(math.pi / 4)
    params = reference_init(model, N, M, np.random.default_rng(5))
    plan = ShardPlan(K, world, rank)
    e = Engine(model, N, M, h=h, J=J, alpha=2.0, seed=11, device=local, structured_sv=structured, **plan.engine_kwargs())
    e.set_params(params)
    bootstrap_comm(e, world, rank, p2p=False)
    p2p = enable_p2p(e, world) if use_p2p else False
    e.warm_up(30)
    res = []
    for _ in range(5):
        st = e.sr_step(n_mc_steps=1, lr=0.05)
        res.append([st.e_mean.real, st.e_mean.imag, st.rsd, st.lam, st.cg_iters])
    final = e.get_params()
    gathered = [None] * world
    dist.all_gather_object(gathered, final.view(np.float64).tolist())
    if rank == 0:
        same = all(g == gathered[0] for g in gathered)   # replicated state must be BIT-identical on all ranks
        json.dump({"world": world, "p2p": bool(p2p), "steps": res, "params_re_im": gathered[0], "ranks_identical": same,
                   "sv": e.kernel_variant("sv")},
                  open(out_path, "w"))
    e.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
