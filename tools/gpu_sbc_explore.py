import sys, os
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import gpslc_b200 as g
import test_gpu_sbc as T
ctx = g.Context(0)
for thin, burn in ((6, 30), (20, 60)):
    T.THIN, T.BURN = thin, burn
    pv = T.run_sbc(ctx, 1, 2024)
    print("thin", thin, "burn", burn, {k[0] + str(k[1]): round(v, 4) for k, v in sorted(pv.items(), key=lambda kv: kv[1])[:6]})
