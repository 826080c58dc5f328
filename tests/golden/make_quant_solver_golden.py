"""Generates tests/golden/quant_solver_golden.json by IMPORTING the reference's own solver
(/root/reference/training/quantization.py: adjust_quant) -- run in the build container only; the JSON
travels.  Doubles are stored as hex strings so the comparison is bit-exact."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference/training")
import quantization as Q  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
BLU = {22: [0.1111, 0.05, 0.05, 0.022, 0.022, 0], 27: [0.294, 0.172, 0.172, 0.101, 0.101, 0],
       32: [0.316, 0.198, 0.198, 0.125, 0.125, 0], 37: [0.349, 0.243, 0.243, 0.169, 0.169, 0]}   # quantization.py:69-76


def main():
    cases = []
    import pickle
    for qp in (22, 27, 32, 37):
        rows = pickle.load(open("/root/reference/training/quant_params%d.data" % qp, "rb"))
        cases.append(([float(r[0]) for r in rows], BLU[qp]))
    rng = np.random.default_rng(2024)
    for _ in range(40):
        stepw = (10 ** rng.uniform(-3.2, -1.6, 6)).tolist()
        blu = rng.uniform(0.02, 0.5, 6).tolist()
        cases.append((stepw, blu))
    out = []
    for stepw, blu in cases:
        rows = Q.adjust_quant(list(stepw), list(blu))
        out.append({"stepw": [float(v).hex() for v in stepw], "blu": [float(v).hex() for v in blu],
                    "rows": [[float(v).hex() for v in r] for r in rows]})
    json.dump(out, open(os.path.join(HERE, "quant_solver_golden.json"), "w"), indent=0)
    print("wrote %d cases" % len(out))


if __name__ == "__main__":
    main()
