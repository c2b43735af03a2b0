"""Parity at the sizes BASELINE.json quotes (VERDICT r01 "next round" item 1): one full generation of config 3
(N=1000, lambda=65536: the persistent 64-tile-row GEMM walk, the 4-way split-K SYRK with a device-side row count, the radix
sort) and of ONE RANK'S SHARD of config 4 (N=4096, 2^17 mirrored samples = 65536 z rows), checked piece by piece:

  Z     device Philox (stand-alone launch of the same kernel)
  Y     'BDZ Matrix' of the generation loop  vs  numpy f64  Z (B diag D)^T               <= 2e-14   (sampleSingle :494-513)
  F     'Value Vector'                        vs  oracle objective on x = m + sigma y      bit-exact (Optimization::evaluate)
  idx   'Sorting Index'                       vs  oracle sort_index                        bit-exact (:940-950)
  mean / ps / pc / sigma / C after tell       vs  oracle tell() (config 3; adaptC :690-718 loop for loop, ~1 min of host time)
                                              vs  a numpy restatement of :547-761 (config 4 shard: the oracle's scalar adaptC would
                                                  take half an hour at N=4096, mu=65536)    <= 1e-11

The hsig = 0 branch (:658, :662, :701) is asserted to FIRE here as well (the reference fixture never reaches it)."""
import numpy as np
import pytest
import torch
from conftest import relerr
from korali_b200 import _lib
from korali_b200._abi import INJ_BD, INJ_BDZ, INJ_F
from oracle import oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-11

if not torch.cuda.is_available():
    pytest.skip("needs a CUDA device", allow_module_level=True)

STATE_ARR = ["Covariance Matrix", "Current Mean", "Previous Mean", "Evolution Path", "Conjugate Evolution Path", "Best Ever Variables"]
STATE_SCA = ["Sigma", "Best Ever Value", "Current Best Value", "Previous Best Value", "Previous Best Ever Value"]


def copy_state(src, dst):
    for k in STATE_ARR:
        dst.set(k, src.get(k))
    for k in STATE_SCA:
        dst.set_scalar(k, src.scalar(k))
    dst.set_scalar("Current Generation", src.scalar("Current Generation"))
    dst.set_scalar("Model Evaluation Count", src.scalar("Model Evaluation Count"))


def test_config3_full_size_generation_against_oracle():
    case = dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
    n, lam = case["n"], case["population_size"]
    s = _lib.Solver(**case)
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    s.run_generation(); s.run_generation()          # C has moved away from I: B is a dense rotation in generation 3
    o = O.Oracle(**case); o.set_scalar("Oracle/RNG Kind", 1)
    copy_state(s, o)
    mean, sigma = s.get("Current Mean"), s.scalar("Sigma")
    gen = int(s.scalar("Current Generation")) + 1
    s.ask()
    b = s.get("Covariance Eigenvector Matrix").reshape(n, n); d = s.get("Axis Lengths")
    c = s.get("Covariance Matrix").reshape(n, n)
    assert np.abs((b * d**2) @ b.T - c).max() < 1e-12 * np.abs(c).max() and np.abs(b.T @ b - np.eye(n)).max() < 1e-12
    y = s.get("BDZ Matrix").reshape(lam, n)
    z = _lib.k_philox_normal(case["seed"], gen, 0, lam, n)
    assert relerr(y, z @ (b * d).T) < 2e-14
    del z
    x = mean + sigma * y
    s.eval()
    f = s.get("Value Vector")
    assert np.array_equal(f, O.objective("NegEllipsoid", x, s.get("Objective Coefficients")))
    del x
    o.inject(INJ_BD, np.concatenate([b.ravel(), d])); o.inject(INJ_BDZ, y.ravel()); o.inject(INJ_F, f)
    del y
    o.ask(); o.eval(); o.tell()
    s.tell()
    assert np.array_equal(s.get_index("Sorting Index"), O.sort_index(f))
    assert np.array_equal(s.get_index("Sorting Index"), o.get_index("Sorting Index"))
    for k in ["Current Mean", "Mean Update", "Evolution Path", "Conjugate Evolution Path", "Covariance Matrix", "Best Ever Variables",
              "Current Best Variables"]:
        e = relerr(s.get(k), o.get(k))
        assert e < TOL, (k, e)
    for k in ["Sigma", "Conjugate Evolution Path L2 Norm", "Best Ever Value", "Current Best Value", "Current Min Standard Deviation",
              "Current Max Standard Deviation"]:
        a, r = s.scalar(k), o.scalar(k)
        assert abs(a - r) <= TOL * abs(r), (k, a, r)
    cn = s.get("Covariance Matrix").reshape(n, n)
    assert np.array_equal(cn, cn.T)
    s.close()


def numpy_tell(state, x_sel, w, n, gen):
    """updateDistribution + adaptC + updateSigma (:547-761) for the unconstrained, non-diagonal case, vectorised."""
    m, sigma, cmat, ps, pc, b, d = (state[k] for k in ("mean", "sigma", "C", "ps", "pc", "B", "D"))
    cs, cc, mueff, damp, chi = (state[k] for k in ("cs", "cc", "mueff", "damp", "chi"))
    mean_new = w @ x_sel
    yv = (mean_new - m) / sigma
    t = (b.T @ yv) / d
    ps_new = (1. - cs) * ps + np.sqrt(cs * (2. - cs) * mueff) * (b @ t)
    psn = np.sqrt(np.sum(ps_new**2))
    hsig = 1.0 if (1.4 + 2.0 / (n + 1) > psn / np.sqrt(1. - (1. - cs)**(2.0 * (1.0 + gen))) / chi) else 0.0
    pc_new = (1. - cc) * pc + hsig * np.sqrt(cc * (2. - cc) * mueff) * yv
    c1 = 2.0 / ((n + 1.3)**2 + mueff)
    cmu = min(1.0 - c1, 2.0 * (mueff - 2. + 1. / mueff) / ((n + 2.0)**2 + mueff))
    tt = (x_sel - m) * np.sqrt(w)[:, None]
    p = tt.T @ tt / sigma**2
    c_new = (1 - c1 - cmu) * cmat + c1 * (np.outer(pc_new, pc_new) + (1 - hsig) * cc * (2. - cc) * cmat) + cmu * p
    sigma_new = sigma * np.exp(cs / damp * (psn / chi - 1.))
    return mean_new, ps_new, pc_new, c_new, sigma_new, psn


def test_config4_one_rank_shard_full_size():
    """The shapes ONE of 8 ranks sees in config 4: N=4096, 2^17 mirrored samples (65536 z rows), 65536 selected rows."""
    case = dict(n=4096, population_size=1 << 17, mirrored_sampling=1, objective="NegSphere", initial_value=1.0, initial_stddev=1.0, seed=1337)
    n, lam = case["n"], case["population_size"]
    mu = lam // 2
    s = _lib.Solver(**case)
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    s.run_generation()
    st = dict(mean=s.get("Current Mean"), sigma=s.scalar("Sigma"), C=s.get("Covariance Matrix").reshape(n, n),
              ps=s.get("Conjugate Evolution Path"), pc=s.get("Evolution Path"),
              cs=s.scalar("Sigma Cumulation Factor"), cc=s.scalar("Cumulative Covariance"), mueff=s.scalar("Effective Mu"),
              damp=s.scalar("Damp Factor"), chi=s.scalar("Chi Square Number"))
    gen = int(s.scalar("Current Generation")) + 1
    s.ask()
    b = s.get("Covariance Eigenvector Matrix").reshape(n, n); d = s.get("Axis Lengths")
    assert np.abs((b * d**2) @ b.T - st["C"]).max() < 1e-12 * np.abs(st["C"]).max() and np.abs(b.T @ b - np.eye(n)).max() < 1e-12
    st["B"], st["D"] = b, d
    y = s.get("BDZ Matrix").reshape(lam, n)                 # mirrored pairs expanded: rows 2i, 2i+1 = +y_i, -y_i
    assert np.array_equal(y[1::2], -y[0::2])
    z = _lib.k_philox_normal(case["seed"], gen, 0, lam // 2, n)
    assert relerr(y[0::2], z @ (b * d).T) < 2e-14
    del z
    x = st["mean"] + st["sigma"] * y
    del y
    s.eval()
    f = s.get("Value Vector")
    assert np.array_equal(f, O.objective("NegSphere", x))
    s.tell()
    idx = s.get_index("Sorting Index")
    assert np.array_equal(idx, O.sort_index(f))
    w = s.get("Mu Weights")
    assert w.size == mu
    mean_new, ps_new, pc_new, c_new, sigma_new, psn = numpy_tell(st, x[idx[:mu].astype(np.int64)], w, n, gen)
    del x
    assert relerr(s.get("Current Mean"), mean_new) < TOL
    assert relerr(s.get("Conjugate Evolution Path"), ps_new) < TOL
    assert relerr(s.get("Evolution Path"), pc_new) < TOL
    assert relerr(s.get("Covariance Matrix").reshape(n, n), c_new) < TOL
    assert abs(s.scalar("Sigma") - sigma_new) <= TOL * sigma_new
    assert abs(s.scalar("Conjugate Evolution Path L2 Norm") - psn) <= TOL * psn
    s.close()


def test_hsig_zero_branch_fires_and_matches_oracle():
    """Start far from the optimum with a tiny step size: consecutive mean shifts point the same way, |ps| outgrows the threshold and
    hsig = 0 (:658). Then pc <- (1-cc) pc exactly (:662) and adaptC adds the (1-hsig) cc (2-cc) C term (:701). Free-running lockstep of
    device and oracle on the same Philox stream; the branch must be seen, on both."""
    case = dict(n=10, population_size=32, objective="NegSphere", initial_value=10.0, initial_stddev=0.01, seed=11)
    n = case["n"]
    s = _lib.Solver(**case); o = O.Oracle(**case); o.set_scalar("Oracle/RNG Kind", 1)
    cs, cc, chi = (o.scalar(k) for k in ("Sigma Cumulation Factor", "Cumulative Covariance", "Chi Square Number"))
    saw_hsig0 = saw_hsig1 = 0
    for g in range(30):
        pc_s, pc_o = s.get("Evolution Path"), o.get("Evolution Path")
        s.run_generation(); o.run_generation()
        gen = o.scalar("Current Generation")
        hs = []
        for h in (s, o):
            psn = h.scalar("Conjugate Evolution Path L2 Norm")
            hs.append(1.4 + 2.0 / (n + 1) > psn / np.sqrt(1. - (1. - cs)**(2.0 * (1.0 + gen))) / chi)
        assert hs[0] == hs[1], g
        if not hs[0]:
            saw_hsig0 += 1
            assert np.array_equal(s.get("Evolution Path"), (1. - cc) * pc_s), g
            assert np.array_equal(o.get("Evolution Path"), (1. - cc) * pc_o), g
        else:
            saw_hsig1 += 1
        assert np.array_equal(s.get_index("Sorting Index"), o.get_index("Sorting Index")), g
        for k in ["Current Mean", "Evolution Path", "Conjugate Evolution Path", "Covariance Matrix"]:
            assert relerr(s.get(k), o.get(k)) < 1e-9, (g, k)
        assert abs(s.scalar("Sigma") - o.scalar("Sigma")) < 1e-9 * o.scalar("Sigma"), g
    assert saw_hsig0 >= 5 and saw_hsig1 >= 5, (saw_hsig0, saw_hsig1)
    s.close()


def test_config3_full_size_converges_to_the_optimum_within_1e8():
    """BASELINE.json: convergence to the same optimum within 1e-8 on every config — config 3 at its stated size (N = 1000,
    lambda = 65536, ill-conditioned ellipsoid, optimum 0 at x = 0), through kcma_run with the termination chain."""
    case = dict(n=1000, population_size=65536, objective="NegEllipsoid", initial_value=3.0, initial_stddev=1.0, seed=1337)
    s = _lib.Solver(**case)
    s.set_scalar("Termination Criteria/Max Value", -1e-9)
    s.set_scalar("Termination Criteria/Max Generations", 12000)   # ~6000 are needed (cond(C) has to grow to 1e6): about 70 s
    s.set_scalar("Termination Criteria/Max Model Evaluations", 1e18)
    done = s.run(12001)
    best = s.scalar("Best Ever Value")
    fin, reason = s.check_termination()
    print("config 3 at full size: best %.3e after %d generations (%s), max eigenvalue ratio %.3g" %
          (best, done, reason, s.scalar("Maximum Covariance Eigenvalue") / s.scalar("Minimum Covariance Eigenvalue")))
    assert fin and "Max Value" in reason and abs(best) < 1e-8, (best, done, reason)
    assert np.abs(s.get("Best Ever Variables")).max() < 1e-3
    s.close()
