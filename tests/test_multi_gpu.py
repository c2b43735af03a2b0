"""N > 1: the population is sharded across ranks; collectives are an all-gather of F and an all-reduce of
[partial rank-mu sum | partial mean | best sample].
 - GPU (needs >= 2 devices): torchrun + NCCL, checked against the single-GPU run.
 - CPU: the same decomposition with world_size 2 over gloo, arithmetic done by the oracle (checker only)."""
import os
import subprocess
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_population_matches_single_gpu():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("ok ") == 8


def _korali_experiment(objective="Rosenbrock", n=64, pop=512, gens=12, mirrored=False, seed=21):
    import korali_b200 as korali
    e = korali.Experiment()
    e["Random Seed"] = seed
    e["Problem"]["Type"] = "Optimization"
    e["Problem"]["Objective Function"] = objective
    for i in range(n):
        e["Variables"][i]["Name"] = "X%d" % i
        e["Variables"][i]["Initial Value"] = 0.2
        e["Variables"][i]["Initial Standard Deviation"] = 0.8
    e["Solver"]["Type"] = "Optimizer/CMAES"
    e["Solver"]["Population Size"] = pop
    e["Solver"]["Mirrored Sampling"] = mirrored
    e["Solver"]["Termination Criteria"]["Max Generations"] = gens
    e["Console Output"]["Verbosity"] = "Silent"
    e["File Output"]["Enabled"] = False
    return e


def _korali_run(devices, **kw):
    import korali_b200 as korali
    e = _korali_experiment(**kw)
    k = korali.Engine()
    k["Conduit"]["Type"] = "Device"
    if devices is not None:
        k["Conduit"]["Devices"] = devices
    k.run(e)
    return e


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_engine_runs_several_experiments_concurrently_on_several_devices():
    """k.run([e1, ..., e5]) with k["Conduit"]["Devices"] = 2 (engine.cpp:98-111 interleaves experiments by coroutine switches on one
    core): experiment j runs on device j mod 2, the two device threads run side by side. An experiment never spans devices in this
    mode, so every result is BITWISE the result of running that experiment alone."""
    import korali_b200 as korali
    seeds = [3, 5, 8, 13, 21]
    alone = [_korali_run(None, seed=s_, n=40, pop=256, gens=15) for s_ in seeds]
    es = [_korali_experiment(seed=s_, n=40, pop=256, gens=15) for s_ in seeds]
    k = korali.Engine()
    k["Conduit"]["Type"] = "Device"
    k["Conduit"]["Devices"] = 2
    k.run(es)
    for a, b in zip(alone, es):
        assert b["Current Generation"] == 15
        assert a["Results"]["Best Sample"]["F(x)"] == b["Results"]["Best Sample"]["F(x)"]
        assert a["Solver"]["Current Mean"] == b["Solver"]["Current Mean"]
        assert a["Solver"]["Sigma"] == b["Solver"]["Sigma"]
    assert len({e["Results"]["Best Sample"]["F(x)"] for e in es}) == len(seeds)   # different seeds, different runs


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("devices", [2, [1, 0]])
def test_engine_devices_shards_the_population_behind_the_korali_api(devices):
    """k["Conduit"]["Devices"] = G: one process, one handle + host thread per device, kcma_comm_init_all — the Korali script is the
    only thing that changes (the reference picks its Distributed conduit the same way, distributed.cpp.base:13-90). Same Philox
    samples as on one device, so the runs agree up to the all-reduce summation order."""
    for mirrored in (False, True):
        e1 = _korali_run(None, mirrored=mirrored)
        e2 = _korali_run(devices, mirrored=mirrored)
        assert e1["Current Generation"] == e2["Current Generation"] == 12
        assert e1["Solver"]["Model Evaluation Count"] == e2["Solver"]["Model Evaluation Count"] == 12 * 512
        b1, b2 = e1["Results"]["Best Sample"]["F(x)"], e2["Results"]["Best Sample"]["F(x)"]
        assert abs(b1 - b2) <= 1e-9 * abs(b1)
        m1, m2 = np.array(e1["Solver"]["Current Mean"]), np.array(e2["Solver"]["Current Mean"])
        assert np.abs(m1 - m2).max() <= 1e-8 * np.abs(m1).max()
        assert abs(e1["Solver"]["Sigma"] - e2["Solver"]["Sigma"]) <= 1e-8 * e1["Solver"]["Sigma"]


@pytest.mark.gpu
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_engine_devices_rejects_python_models():
    with pytest.raises(RuntimeError, match="device objective"):
        _korali_run(2, objective=lambda s: s.__setitem__("F(x)", -sum(x * x for x in s["Parameters"])), n=4, pop=8, gens=2)


def _gloo_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from korali_b200 import _lib
    from oracle import oracle as O
    N, lam, mu = 12, 64, 32
    for mirrored in (0, 1):
        rng = np.random.default_rng(3)                       # every rank draws the same global population
        mean, sigma = rng.standard_normal(N), 0.7
        y = rng.standard_normal((lam // (2 if mirrored else 1), N))
        if mirrored:
            y = np.repeat(y, 2, axis=0); y[1::2] *= -1
        x = mean + sigma * y
        w = np.log(mu + 0.5) - np.log(np.arange(mu) + 1.0); w /= w.sum()
        lo, hi = _lib.shard_range(lam, mirrored, rank, world)
        # local objective on the shard, all-gather(F)
        f_local = torch.from_numpy(O.objective("NegRosenbrock", x[lo:hi]))
        parts = [torch.empty(hi - lo, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, f_local)
        f = torch.cat(parts).numpy()
        idx = O.sort_index(f)                                 # identical on every rank
        sel = [(r, int(i)) for r, i in enumerate(idx[:mu]) if lo <= i < hi]
        t = np.array([x[i] - mean for _, i in sel]).reshape(-1, N)
        wl = np.array([w[r] for r, _ in sel])
        part = np.zeros(N * N + N)
        if len(sel):
            part[:N * N] = O.rank_mu(t, wl).ravel()
            part[N * N:] = (wl[:, None] * np.array([x[i] for _, i in sel])).sum(0)
        red = torch.from_numpy(part)
        dist.all_reduce(red)                                  # all-reduce([P | mean])
        full_p = O.rank_mu(x[idx[:mu].astype(int)] - mean, w)
        full_m = (w[:, None] * x[idx[:mu].astype(int)]).sum(0)
        err_p = np.abs(red.numpy()[:N * N].reshape(N, N) - full_p).max() / np.abs(full_p).max()
        err_m = np.abs(red.numpy()[N * N:] - full_m).max() / np.abs(full_m).max()
        q.put((rank, mirrored, float(err_p), float(err_m), len(sel)))
    dist.destroy_process_group()


def test_sharded_update_decomposition_over_gloo():
    """world_size 2 on CPU: shard ranges + all-gather(F) + all-reduce of the partial rank-mu sums reproduce the
    single-rank update (host-side logic of the N > 1 path; the oracle stands in for the kernels)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, mirrored, ep, em, nsel in res:
        assert ep < 1e-13 and em < 1e-13, (rank, mirrored, ep, em)
    for mirrored in (0, 1):
        assert sum(r[4] for r in res if r[1] == mirrored) == 32   # every selected sample is owned by exactly one rank
