#!/usr/bin/env python
"""Generates tests/golden/apd_golden_lm_v1.npz from the CPU oracle: LM traces of the cases in
tests/lm_cases.py (rejected trials, rejected-but-converged, "lm not converged!!", non-default initial lambda).

These vectors pin the ORACLE's walk through lsq_registration_impl.hpp:127-173; tests/test_reference_apdgicp.py::test_lm_branches
runs the same cases through the reference's own compiled sources (oracle/ref_apdgicp.cpp) and checks these vectors against them,
and tests/golden/apd_ref_golden_v1.npz holds the reference-made twins (lm_* keys).

    python tests/golden/make_golden_lm.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    import lm_cases as L
    src, tgt, _ = L.make_pair()
    out = {}
    for name in L.CASES:
        r = L.run_oracle(name, src, tgt)
        out[f"{name}_T"] = r["T"]
        out[f"{name}_state"] = np.array([int(r["converged"]), r["iterations"], int(r["lm_failed"])])
        out[f"{name}_trace"] = r["trace"]
        out[f"{name}_final_hessian"] = r["final_hessian"]
        out[f"{name}_fitness"] = np.array(r["fitness"])
        tr = r["trace"]
        print(f"{name}: rows {len(tr)} rejected {int((tr[:, 7] == 0).sum())} converged {r['converged']} iterations {r['iterations']} lm_failed {r['lm_failed']}")
    path = os.path.join(HERE, "apd_golden_lm_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
