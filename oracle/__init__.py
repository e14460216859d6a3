"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference (BICLab/Spike2Former) hot path plus a loader
that executes the reference's own files when /root/reference is present.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import anything from here.  The product package
(spike2former_b200/) never imports oracle/: it fails loudly when its CUDA
library is missing instead of falling back to this code.

Parity status: the reference ships no golden vectors for this path
(SURVEY.md section 8c), so the port (oracle/port.py) is pinned against outputs of
the reference itself, produced in the build container by oracle/ref_loader.py
and committed under tests/golden/ together with tests/golden/make_golden.py.
"""
