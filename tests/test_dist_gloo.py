"""N>1 host logic on CPU (gloo, world_size 2): the shard plan, the global-chain-id RNG keying and the all-reduce protocol
(5P+3 doubles per SR iteration, P complex per CG iteration) give the SAME trajectory as one rank holding every chain.
The compute here is the oracle (test infrastructure); the product path on GPUs enqueues the same two all-reduces through NCCL
inside libnqs_b200.so (tests/test_gpu_multi.py checks that on the box)."""
import math
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, N, M, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from oracle import nqs_oracle as o
    from neural_network_quantum_state_b200.dist import ShardPlan, sr_allreduce_layout
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = ShardPlan(K, world, rank)
    sizes = []

    def reduce(x):
        x = np.ascontiguousarray(x)
        view = x.view(np.float64) if np.iscomplexobj(x) else x.astype(np.float64)
        t = torch.from_numpy(view.copy())
        dist.all_reduce(t)
        sizes.append(t.numel())
        r = t.numpy()
        return r.view(np.complex128).reshape(x.shape) if np.iscomplexobj(x) else r.reshape(x.shape)

    params, h, J, alpha = _problem(N, M)
    m = o.RBM(N, M, plan.n_local)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, h, J, alpha, False, o.UniformSource(plan.n_local, seed=77, chain_offset=plan.offset))
    s.warm_up(6)
    sr = o.StochasticReconfigurationCG(plan.n_local, m.P, reduce=reduce, n_total=K)
    E = []
    for _ in range(3):
        st = sr.step(s, 1, 0.05)
        E.append(st.e_mean)
    lay = sr_allreduce_layout(m.P)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), params=m.variables, E=np.array(E), spins=m.spins,
             offset=plan.offset, n_local=plan.n_local, cg_reduce_doubles=2 * m.P, setup_doubles=lay["count"],
             sizes=np.array(sizes))
    dist.destroy_process_group()


def _problem(N, M):
    rng = np.random.default_rng(5)
    sw = math.sqrt(1.0 / (N + M))
    W = 0.8 * (rng.normal(0, sw, (N, M)) + 1j * rng.normal(0, sw, (N, M)))
    a = 0.3 * (rng.normal(size=N) + 1j * rng.normal(size=N))
    b = 0.2 * (rng.normal(size=M) + 1j * rng.normal(size=M))
    return np.concatenate([W.ravel(), a, b]), -math.cos(math.pi / 4), math.sin(math.pi / 4), 2.0


def test_shard_plan_covers_all_chains():
    from neural_network_quantum_state_b200.dist import ShardPlan
    for K, world in ((16384, 8), (10, 3), (7, 7), (65536, 8)):
        spans = [(ShardPlan(K, world, r).offset, ShardPlan(K, world, r).n_local) for r in range(world)]
        assert spans[0][0] == 0
        for (o0, n0), (o1, _) in zip(spans, spans[1:]):
            assert o0 + n0 == o1
        assert spans[-1][0] + spans[-1][1] == K
    with pytest.raises(ValueError):
        ShardPlan(2, 3, 0)


def test_two_rank_gloo_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp
    from oracle import nqs_oracle as o
    K, N, M, world = 48, 10, 12, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, K, N, M, str(tmp_path)), nprocs=world, join=True)
    params, h, J, alpha = _problem(N, M)
    m = o.RBM(N, M, K)
    m.variables = params.copy()
    s = o.LITFIChainSampler(m, h, J, alpha, False, o.UniformSource(K, seed=77))
    s.warm_up(6)
    sr = o.StochasticReconfigurationCG(K, m.P)
    E = [sr.step(s, 1, 0.05).e_mean for _ in range(3)]
    shards = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    spins = np.concatenate([sh["spins"] for sh in shards])
    assert np.array_equal(spins, m.spins)                      # chains independent of the number of ranks
    for sh in shards:
        assert np.abs(sh["params"] - m.variables).max() < 1e-10 * np.abs(m.variables).max()
        assert np.abs(sh["E"] - np.array(E)).max() < 1e-12
    assert np.array_equal(shards[0]["params"], shards[1]["params"])   # replicated state stays bit-identical
    # the only vector all-reduces are 2P doubles (CG) and <= 2P-sized pieces of the 5P+3 setup buffer
    assert int(shards[0]["sizes"].max()) <= 2 * m.P
