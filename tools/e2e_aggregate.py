"""Aggregate per-seed outputs of tools/e2e_slam.py into the system-level criterion of north_star: mean ATE / depth-L1
ratios (this build vs the reference's PyTorch path on the same GPU) with a confidence interval.

    python tools/e2e_aggregate.py gpurun_out/e2e500_s*.json > profiles/r02_e2e_slam.json
"""
import json
import math
import sys

import numpy as np

T95 = {2: 12.71, 3: 4.30, 4: 3.18, 5: 2.78, 6: 2.57, 7: 2.45, 8: 2.36, 9: 2.31, 10: 2.26, 11: 2.23, 12: 2.20, 13: 2.18,
       14: 2.16, 15: 2.14, 16: 2.13, 20: 2.09, 24: 2.07}


def ci(x):
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if n < 2:
        return float(x.mean()), None
    t = T95.get(n, 2.0 if n > 24 else T95[min(k for k in T95 if k >= n)])
    return float(x.mean()), float(t * x.std(ddof=1) / math.sqrt(n))


def main():
    runs = [json.load(open(p)) for p in sys.argv[1:]]
    a_ate = [r["b200"]["ate_rmse_m"] for r in runs]
    b_ate = [r["reference_torch_gpu"]["ate_rmse_m"] for r in runs]
    a_l1 = [r["b200"]["depth_l1_m"] for r in runs]
    b_l1 = [r["reference_torch_gpu"]["depth_l1_m"] for r in runs]
    out = {"config": runs[0]["config"], "seeds": [r["config"]["seed"] for r in runs], "n": len(runs)}
    for name, a, b in (("ate_rmse_m", a_ate, b_ate), ("depth_l1_m", a_l1, b_l1)):
        ma, ha = ci(a)
        mb, hb = ci(b)
        # the criterion compares the MEANS of the two arms; its uncertainty from the paired per-seed differences
        d = np.asarray(a) - np.asarray(b)
        md, hd = ci(d)
        out[name] = {"b200_mean": ma, "b200_ci95": ha, "reference_mean": mb, "reference_ci95": hb,
                     "ratio_of_means": ma / mb, "ratio_ci95_halfwidth": (hd / mb) if hd is not None else None,
                     "per_seed_b200": a, "per_seed_reference": b}
    out["speedup_wall"] = float(np.mean([r["speedup_wall"] for r in runs]))
    out["wall_s"] = {"b200": float(np.mean([r["b200"]["wall_s"] for r in runs])),
                     "reference_torch_gpu": float(np.mean([r["reference_torch_gpu"]["wall_s"] for r in runs]))}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
