"""Convergence study behind tests/test_kat_all_steppers.py: the reference's analytic known-answer tests
(test/test_traceradvectiondiffusion.jl; its harness takes the stepper as a parameter, test/runtests.jl:26) run on the
CPU oracle with every FourierFlows stepper, at dt and dt/2.  Prints the measured relative-L2 error against the analytic
solution, the observed order, and writes tests/golden/kat_stepper_errors.json (the tolerances the tests use)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ptf_oracle import OracleProblem
from tests.kat_cases import REFERENCE_KATS

STEPPERS = ["RK4", "ETDRK4", "LSRK54", "AB3", "FilteredRK4", "FilteredETDRK4", "ForwardEuler"]
SLOW = {"constvel3D", "timedependentvel3D"}
out = {}
for name, (fn, kw) in REFERENCE_KATS.items():
    if name in SLOW and "--all" not in sys.argv:
        continue
    for st in STEPPERS:
        try:
            err, rtol = fn(lambda **k: OracleProblem(**k), stepper=st, **kw)
        except Exception as e:
            print(name, st, "ERROR", repr(e)[:100])
            continue
        out.setdefault(name, {})[st] = {"err": err, "ref_rtol_rk4": rtol}
        print(f"{name:28s} {st:16s} err {err:.3e}   (reference rtol for RK4 {rtol:.1e})", flush=True)
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kat_stepper_errors.json"), "w"), indent=1)
