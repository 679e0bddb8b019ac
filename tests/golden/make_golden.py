"""Regenerates tests/golden/ from the read-only reference checkout (run in the build container only:
/root/reference does not exist on the GPU box). Two kinds of fixtures:

1. reference_kats.json — the known-answer values the reference's OWN tests assert for the hot path, transcribed with
   their file:line (the reference is Julia and cannot be executed here, so these are the pins the oracle is checked
   against).
2. data/, results/ — the reference's datasets and its golden ITE summaries (test/test_results/*.csv), copied verbatim;
   they are data fixtures, not source code.
"""
import json
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

kats = {
    "rbfKernelLog_magic": {  # test/kernel.jl:56-67 (Matrix input and Vector-of-Vectors input)
        "cite": "test/kernel.jl:56-67",
        "X": [[1, 2], [3, 4], [5, 6]], "LS": 1, "expected": [[0, -8, -32], [-8, 0, -8], [-32, -8, 0]]},
    "rbfKernelLog_ones": {  # test/kernel.jl:50-55
        "cite": "test/kernel.jl:50-55", "X_shape": [10, 5], "X_fill": 1.0, "LS": 0.1, "expected_fill": 0.0},
    "rbfKernelLogScalar_same_point": {  # test/kernel.jl:5-39: -(v - v)^2 / 0.1 for v in 11, 11.1, true, false
        "cite": "test/kernel.jl:5-39", "values": [11, 11.1, 1, 0], "LS": 0.1, "expected": 0.0},
    "processCov_scale": {"cite": "test/kernel.jl:70-75", "logCov": [[0.0]], "scale": 2.0, "expected": [[2.0]]},
    "processCov_noise": {"cite": "test/kernel.jl:76-82", "logCov": [[0.0]], "scale": 0.0, "noise": 1e-5, "expected": [[1e-5]]},
    "processCov_scale_noise": {"cite": "test/kernel.jl:83-89", "logCov": [[0.0]], "scale": 2.0, "noise": 1e-5,
                               "expected": [[2.0 + 1e-5]]},
    "logit_half": {"cite": "test/kernel.jl:91-93", "p": 0.5, "expected": 0.0},
    "expit_zero": {"cite": "test/kernel.jl:94-96", "x": 0.0, "expected": 0.5},
    "generateSigmaU": {  # test/utils.jl:2-16
        "cite": "test/utils.jl:2-16", "counts": [2, 3], "eps": 0.1, "cov": 2.0,
        "expected": [[1.1, 2.0, 0, 0, 0], [2.0, 1.1, 0, 0, 0], [0, 0, 1.1, 2.0, 2.0], [0, 0, 2.0, 1.1, 2.0],
                     [0, 0, 2.0, 2.0, 1.1]]},
    "removeAdjacent": {"cite": "test/utils.jl:17-22", "input": [1, 2, 2, 3, 4, 4, 5, 3, 4], "expected": [1, 2, 3, 4, 5, 3, 4]},
    "toMatrix_shape": {"cite": "test/utils.jl:23-34", "n_vectors": 10, "len": 5, "n": 10, "m": 5},
    "conditionalITE_zero_effect": {  # test/estimation.jl:6-67 with test/test_data.jl:35-52 parameters (n = 1)
        "cite": "test/estimation.jl:6-67; test/test_data.jl:35-52",
        "uyLS": [1.0], "xyLS": [1.0], "tyLS": 1.0, "yScale": 1.0, "yNoise": 1.0, "U": [[1.0]], "X": [[1.0]],
        "realT": [1.0], "binaryT": [True], "doT_real": 1.0, "doT_binary": True,
        "expected_mean": 0.0, "expected_cov": 0.0},
    "conditionalSATE_zero_effect": {"cite": "test/estimation.jl:69-137", "expected_mean": 0.0, "expected_var": 0.0},
    "ITEDistributions_jitter": {  # test/estimation.jl:139-247: mean 0, cov == predictionCovarianceNoise
        "cite": "test/estimation.jl:139-247", "predictionCovarianceNoise": 1e-10},
    "summarizeEstimates_quantiles": {  # test/driver.jl:54-71
        "cite": "test/driver.jl:54-71", "samples": list(range(0, 101)),
        "intervals": {"0.9": [5.0, 95.0], "0.8": [10.0, 90.0]}},
    "numPosteriorSamples": {"cite": "test/utils.jl:50-55; src/hyperparameters.jl:86-92", "nOuter": 24, "nBurnIn": 10,
                            "stepSize": 1, "expected": 15},
    "NEEC_gate": {"cite": "test/driver.jl:46-52; test/test_utils.jl:3-12", "data": "data/NEEC_sampled.csv",
                  "golden": "results/NEEC_sampled_0.6.csv", "doT": 0.6, "min_fraction_inside": 0.5},
    "default_hyperparameters": {"cite": "src/hyperparameters.jl:86-92", "nU": 1, "nOuter": 24, "nMHInner": 10, "nESInner": 5,
                                "nBurnIn": 10, "stepSize": 1, "predictionCovarianceNoise": 1e-10},
}

if __name__ == "__main__":
    json.dump(kats, open(os.path.join(HERE, "reference_kats.json"), "w"), indent=1)
    for f in ["NEEC_sampled.csv", "IHDP_sampled.csv", "minimal.csv", "no_cov.csv", "no_objects.csv", "no_objects_no_cov.csv",
              "additive_linear.csv", "additive_nonlinear.csv", "multiplicative_linear.csv", "multiplicative_nonlinear.csv"]:
        shutil.copy(os.path.join(REF, "test", "test_data", f), os.path.join(HERE, "data", f))
    for f in sorted(os.listdir(os.path.join(REF, "test", "test_results"))):
        if f.endswith(".csv"):
            shutil.copy(os.path.join(REF, "test", "test_results", f), os.path.join(HERE, "results", f))
    print("wrote", len(kats), "KATs and", len(os.listdir(os.path.join(HERE, "data"))), "datasets,",
          len(os.listdir(os.path.join(HERE, "results"))), "golden result files")
