"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package, and only as the checker or the reported CPU baseline.  stf_b200/ never
imports it: the product path has no CPU fallback and fails loudly without its CUDA library.

Parity pins (see DESIGN.md "Oracle"): the reference ships no tests and no golden vectors
(SURVEY.md section 4), so the pins are (1) the live reference imported from /root/reference in the
build container (oracle/ref_import.py, tests/test_oracle_pins.py::*_vs_live_reference), (2) the
golden fixtures under tests/golden/ recorded from it by oracle/gen_golden.py, and (3) the
reference's own C++ coder compiled into oracle/_ref by oracle/Makefile.
"""
