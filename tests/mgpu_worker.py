"""Worker of tests/test_multi_gpu.py (launched with torch.distributed.run, one process per GPU): a population sharded
over WORLD_SIZE ranks must reproduce the single-GPU run: same Philox samples (counters are global sample indices),
identical ranking, and mean / paths / sigma / C equal up to the all-reduce summation order."""
import faulthandler
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from korali_b200 import _lib  # noqa: E402


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def main():
    faulthandler.enable()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [dict(n=64, population_size=512, objective="NegRosenbrock", initial_value=0.2, initial_stddev=0.8, seed=21),
             dict(n=130, population_size=1024, objective="NegEllipsoid", mirrored_sampling=1, initial_value=3.0, initial_stddev=1.0, seed=5),
             dict(n=40, population_size=256, objective="NegSphere", diagonal_covariance=1, initial_value=1.0, initial_stddev=1.0, seed=9),
             # gradient step of the mean: every rank adds its share of sum_i w_i step/sqrt(N) grad_i to the partial mean (all-reduced)
             dict(n=48, population_size=256, objective="NegSphereSin2", initial_value=1.5, initial_stddev=1.0, seed=13,
                  use_gradient_information=1, gradient_step_size=0.02),
             # bounds + resampling rounds: the infeasible counter and the decision to go on resampling are global (all-reduced)
             dict(n=12, population_size=96, objective="NegSphere", initial_value=4.0, initial_stddev=2.5, seed=31, lower_bound=-5.0,
                  upper_bound=5.0, max_infeasible_resamplings=10**9, mirrored_sampling=1),
             # discrete variables: the mutations are keyed by the GLOBAL sample index, so the sharding must not change them
             dict(n=24, population_size=128, objective="NegEllipsoid", initial_value=2.2, initial_stddev=1.5, seed=17, mirrored_sampling=1,
                  granularity=np.array([1.0, 0.0, 0.5, 0.0] * 6)),
             # constraint path (viability regime, then the regime switch): the correction loop of handleConstraints is sequential
             # in the population order, so every rank draws and corrects the whole population and only the model evaluations are
             # sharded (all-gather of F) - the run is BITWISE the one-GPU run
             dict(n=20, population_size=64, viability_population_size=64, objective="NegSphereSin2", constraint_family="HalfSpace",
                  n_constraints=4, constraint_shift=np.array([1.0, 1.0, -1.0, -1.0]), lower_bound=-10.0, upper_bound=10.0,
                  initial_value=np.array([4.0, 4.0, -2.0, -2.0] + [0.0] * 16), initial_stddev=1.0, is_sigma_bounded=1, seed=1337)]
    verbose = os.environ.get("KCMA_TEST_VERBOSE")
    for case in cases:
        if verbose:
            print("rank", rank, "case", case["objective"], "create", flush=True)
        s = _lib.Solver(device=local, rank=rank, nranks=world, **case)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(_lib.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        s.comm_init(bytes(uid.cpu().tolist()))
        ref = _lib.Solver(device=local, **case) if rank == 0 else None
        constrained = "n_constraints" in case
        saw_viability = saw_switch = False
        for g in range(40 if ("granularity" in case or constrained) else 12):   # discrete: long enough for the masking matrix to switch on
            if verbose:
                print("rank", rank, "gen", g, flush=True)
            s.run_generation()
            if constrained:
                v = s.scalar("Is Viability Regime")
                saw_viability |= v != 0
                saw_switch |= saw_viability and v == 0
            if ref is not None and constrained:
                ref.run_generation()
                for k in ["Value Vector", "Current Mean", "Evolution Path", "Conjugate Evolution Path", "Covariance Matrix", "Viability Boundaries",
                          "Normal Constraint Approximation"]:
                    assert np.array_equal(s.get(k), ref.get(k)), (case["objective"], g, k)
                assert np.array_equal(s.get_index("Sorting Index"), ref.get_index("Sorting Index")), g
                for k in ["Sigma", "Best Ever Value", "Infeasible Sample Count", "Resampled Parameter Count", "Covariance Matrix Adaptation Count",
                          "Constraint Evaluation Count", "Is Viability Regime", "Global Success Rate"]:
                    assert s.scalar(k) == ref.scalar(k), (g, k, s.scalar(k), ref.scalar(k))
            elif ref is not None:
                ref.run_generation()
                if g == 0:   # identical samples (global Philox counters) -> identical F and ranking, bit for bit
                    assert np.array_equal(s.get("Value Vector"), ref.get("Value Vector")), (case["objective"], g)
                    assert np.array_equal(s.get_index("Sorting Index"), ref.get_index("Sorting Index")), (case["objective"], g)
                    tol = 1e-13
                else:        # the all-reduce sums C in a different order: last-bit differences feed back through the eigenvectors
                    tol = 1e-9 if g < 12 else 1e-6   # (and keep growing over a long free run)
                assert relerr(s.get("Value Vector"), ref.get("Value Vector")) < tol, (case["objective"], g)
                for k in ["Current Mean", "Evolution Path", "Conjugate Evolution Path", "Covariance Matrix"]:
                    e = relerr(s.get(k), ref.get(k))
                    assert e < tol, (case["objective"], g, k, e)
                assert abs(s.scalar("Sigma") - ref.scalar("Sigma")) < tol * ref.scalar("Sigma")
                assert abs(s.scalar("Best Ever Value") - ref.scalar("Best Ever Value")) <= tol * abs(ref.scalar("Best Ever Value"))
                if g == 0:
                    assert s.scalar("Infeasible Sample Count") == ref.scalar("Infeasible Sample Count"), case["objective"]
        # every rank holds the same replicated state
        c = torch.tensor(s.get("Covariance Matrix"), device="cuda")
        c0 = c.clone(); dist.broadcast(c0, 0)
        assert torch.equal(c, c0), "replicated covariance differs between ranks"
        lo, hi = int(s.scalar("Shard Begin")), int(s.scalar("Shard End"))
        assert (lo, hi) == _lib.shard_range(case["population_size"], case.get("mirrored_sampling", 0), rank, world)
        if "granularity" in case:
            assert s.scalar("Number Of Discrete Mutations") > 0
        if constrained:
            assert saw_viability and saw_switch, "the case must run through the viability regime and leave it"
        s.close()
        if rank == 0:
            print("ok", case["objective"], "world", world, flush=True)
    # error path: a non-finite F(x) on ONE rank must fail the generation on EVERY rank (the flag rides in the all-reduce) instead of
    # leaving the other ranks blocked in a collective
    from korali_b200._abi import KcmaError
    s = _lib.Solver(device=local, rank=rank, nranks=world, n=8, population_size=64, objective="External", keep_population=1,
                    initial_value=1.0, initial_stddev=1.0, seed=3)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(_lib.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    s.comm_init(bytes(uid.cpu().tolist()))
    calls = [0]
    def model(X):
        calls[0] += 1
        f = -np.sum(X * X, axis=1)
        if calls[0] == 3 and rank == world - 1:
            f[0] = np.nan
        return f
    s.set_host_objective(model)
    s.run_generation(); s.run_generation()
    try:
        s.run_generation()
        raise AssertionError("rank %d: the non-finite value of rank %d went unnoticed" % (rank, world - 1))
    except KcmaError as exc:
        assert "Non finite" in str(exc), str(exc)
    s.close()
    if rank == 0:
        print("ok collective-error world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:   # a failed rank must not leave its peers blocked inside a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)
