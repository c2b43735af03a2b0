"""Extracts the reference's saved CMA-ES trajectory into a compact fixture.

Source (read-only, only available in the build container):
  /root/reference/tests/python/plot/cmaes/gen00000000.json .. gen00000100.json
These files are full getConfiguration() dumps written by the reference itself (N=10, lambda=32,
mu=16, objective -sum x^2, seed 0xC0FEE -> Normal Generator seed 790510); see SURVEY.md 4.3.
Output: tests/golden/cmaes_plot_trajectory.npz (arrays stacked over the 101 generations).
Run:  python tests/golden/make_golden.py
"""
import json
import os
import numpy as np

SRC = "/root/reference/tests/python/plot/cmaes"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cmaes_plot_trajectory.npz")

ARR = ["Sample Population", "BDZ Matrix", "Value Vector", "Sorting Index", "Mu Weights", "Current Mean",
       "Previous Mean", "Mean Update", "Evolution Path", "Conjugate Evolution Path", "Covariance Matrix",
       "Covariance Eigenvector Matrix", "Axis Lengths", "Best Ever Variables", "Current Best Variables"]
SCA = ["Sigma", "Effective Mu", "Sigma Cumulation Factor", "Damp Factor", "Cumulative Covariance",
       "Chi Square Number", "Trace", "Conjugate Evolution Path L2 Norm", "Best Ever Value", "Current Best Value",
       "Previous Best Value", "Previous Best Ever Value", "Maximum Covariance Eigenvalue",
       "Minimum Covariance Eigenvalue", "Maximum Diagonal Covariance Matrix Element",
       "Minimum Diagonal Covariance Matrix Element", "Current Min Standard Deviation",
       "Current Max Standard Deviation", "Infeasible Sample Count", "Model Evaluation Count"]


def main():
    out = {}
    gens = []
    for g in range(101):
        with open(os.path.join(SRC, "gen%08d.json" % g)) as fh:
            gens.append(json.load(fh))
    for k in ARR:
        rows = [np.asarray(j["Solver"][k], dtype=np.float64).reshape(-1) for j in gens]
        width = rows[1].size  # generation 0 is written before setInitialConfiguration: empty arrays -> NaN rows
        out[k] = np.stack([r if r.size == width else np.full(width, np.nan) for r in rows])
    for k in SCA:
        out[k] = np.asarray([float(j["Solver"][k]) for j in gens])
    out["Current Generation"] = np.asarray([j["Current Generation"] for j in gens])
    out["Is Finished"] = np.asarray([j["Is Finished"] for j in gens])
    v = gens[1]["Variables"]  # generation 0 is saved before defaults are inferred
    out["Lower Bound"] = np.asarray([x["Lower Bound"] for x in v])
    out["Upper Bound"] = np.asarray([x["Upper Bound"] for x in v])
    out["Initial Value"] = np.asarray([x["Initial Value"] for x in v])
    out["Initial Standard Deviation"] = np.asarray([x["Initial Standard Deviation"] for x in v])
    out["Normal Generator Seed"] = np.asarray([gens[1]["Solver"]["Normal Generator"]["Random Seed"]])
    out["Population Size"] = np.asarray([gens[1]["Solver"]["Population Size"]])
    out["Mu Value"] = np.asarray([gens[1]["Solver"]["Mu Value"]])
    np.savez_compressed(OUT, **{k.replace(" ", "_"): a for k, a in out.items()})
    print("wrote", OUT, os.path.getsize(OUT), "bytes")
    # two reference-written result files as they are (checkpoint interchange: e.loadState() of a file the reference wrote)
    import shutil
    for g in (50, 100):
        dst = os.path.join(os.path.dirname(OUT), "reference_gen%08d.json" % g)
        shutil.copyfile(os.path.join(SRC, "gen%08d.json" % g), dst)
        os.chmod(dst, 0o644)
        print("copied", dst)


if __name__ == "__main__":
    main()
