"""TEST INFRASTRUCTURE.  CPU oracle for the hot path of rlqja1107/NL-VSGG.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from this package; the product (nlvsgg_b200/) never does and fails loudly when
its CUDA library is missing.
"""
