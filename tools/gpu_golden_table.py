"""All 14 golden ITE summaries of the reference (test/test_results/*.csv) against the CUDA path, as a discriminator for the
conventions that cannot be pinned without running Gen: elliptical-slice acceptance rule (ess_rule 0 = Gen's joint weight as
recollected, 1 = textbook), lengthscale convention (l^2 = current source, l = the older one the TODO at src/kernel.jl:17-18 hints
at) and inference budget (default 24/10/5 burn-in 10, and 4x that in outer iterations). C chains per configuration.

Per configuration and golden file:
  gate      fraction of chains that pass the reference's own gate (>= 50 % of the individuals' mean ITE inside the golden 90 %
            interval, test/driver.jl:46-52, test/test_utils.jl:3-12)
  inside    median over chains of that fraction;  pooled = the same fraction for the chain-pooled posterior mean
  corr      median correlation of a chain's mean ITEs with the golden Mean column
  mean/sd   mean and standard deviation over individuals of the pooled mean ITE   (BASELINE.md §2 targets)
  width     mean 90 % interval width (averaged over individuals and chains)
Writes JSON (argv[1]) and prints a markdown table."""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g  # noqa: E402

GOLD = os.path.join(root, "tests", "golden")
CASES = [("NEEC_sampled", ["0", "0.6", "1", "1.0"]), ("additive_linear", ["0", "1"]), ("additive_nonlinear", ["0", "1"]),
         ("multiplicative_linear", ["0", "1"]), ("multiplicative_nonlinear", ["0", "1"]), ("IHDP_sampled", ["true", "false"])]


def dot_value(tag):
    return {"true": True, "false": False}.get(tag, None) if tag in ("true", "false") else float(tag)


def evaluate(ite, exp):
    """ite [C, n, m] draws; exp = golden DataFrame."""
    lo, hi, gm = exp["LowerBound"].values, exp["UpperBound"].values, exp["Mean"].values
    means = ite.mean(axis=2)
    inside = ((lo[None] <= means) & (means <= hi[None])).mean(axis=1)
    with np.errstate(invalid="ignore"):
        corr = np.array([np.corrcoef(m, gm)[0, 1] for m in means])
    q = np.quantile(ite, [0.05, 0.95], axis=2)
    pooled = means.mean(axis=0)
    return {"gate": float((inside >= 0.5).mean()), "inside_median": float(np.median(inside)), "inside_min": float(inside.min()),
            "inside_max": float(inside.max()), "pooled_inside": float(((lo <= pooled) & (pooled <= hi)).mean()),
            "corr_median": float(np.nanmedian(corr)), "mean_of_mean": float(pooled.mean()), "sd_of_mean": float(pooled.std(ddof=1)),
            "ci_width": float((q[1] - q[0]).mean()),
            "golden": {"mean_of_mean": float(gm.mean()), "sd_of_mean": float(gm.std(ddof=1)), "ci_width": float((hi - lo).mean())}}


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_table.json"
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    budgets = {"default": (24, 10), "x4": (96, 40)}
    if len(sys.argv) > 3:
        budgets = {k: budgets[k] for k in sys.argv[3].split(",")}
    ctx = g.Context(0)
    rows = []
    t_all = time.time()
    for ls_unsq in (0, 1):
        os.environ["GPSLC_LS_UNSQUARED"] = str(ls_unsq)
        for rule in (0, 1):
            for bname, (nOuter, nBurn) in budgets.items():
                for name, tags in CASES:
                    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = nOuter, nBurn
                    t0 = time.time()
                    gobj = g.gpslc(os.path.join(GOLD, "data", name + ".csv"), hyperparams=h, seed=100, n_chains=C, ctx=ctx, ess_rule=rule)
                    t_fit = time.time() - t0
                    for tag in tags:
                        ite = g.sampleITE(gobj, dot_value(tag), all_chains=True, ctx=ctx)
                        exp = pd.read_csv(os.path.join(GOLD, "results", f"{name}_{tag}.csv"))
                        r = evaluate(ite, exp)
                        r.update({"file": f"{name}_{tag}.csv", "ess_rule": rule, "ls_unsquared": ls_unsq, "budget": bname, "chains": C,
                                  "fit_seconds": t_fit})
                        rows.append(r)
                        print(f"{r['file']:34s} rule={rule} ls_unsq={ls_unsq} {bname:7s} gate {r['gate']:.2f} inside {r['inside_median']:.2f} "
                              f"pooled {r['pooled_inside']:.2f} corr {r['corr_median']:+.2f} mean {r['mean_of_mean']:+.3f} ({r['golden']['mean_of_mean']:+.3f}) "
                              f"sd {r['sd_of_mean']:.3f} ({r['golden']['sd_of_mean']:.3f}) width {r['ci_width']:.3f} ({r['golden']['ci_width']:.3f})", flush=True)
    os.environ["GPSLC_LS_UNSQUARED"] = "0"
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    json.dump({"chains": C, "seconds": time.time() - t_all, "rows": rows}, open(out_path, "w"), indent=1)
    print("total", time.time() - t_all, "s")


if __name__ == "__main__":
    main()
