#!/usr/bin/env python
"""Three C2 launches (14 CRN scenarios x 1M paths x 252 steps through MonteCarloPricer.greeks) - the target of a single-kernel
`ncu --set full --import-source on -k regex:european_kernel -s 2 -c 1` capture (tools/gpu/r02_c2_source.sh)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import optionslab_b200 as ob  # noqa: E402

pr = ob.MonteCarloPricer(1_000_000, 252, seed=42)
for _ in range(3):
    g = pr.greeks(100.0, 100.0, 1.0, 0.05, 0.2, "call")
print({k: round(v, 6) for k, v in g.items()})
