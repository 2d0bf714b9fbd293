"""TEST INFRASTRUCTURE ONLY — checkers for the CUDA hot path.

oracle/_ref/  the unmodified reference, compiled where it lies (oracle/Makefile `ref`), driven
              through oracle/ref_harness.cpp                      -> oracle.ref
oracle/lm_oracle.c  a plain-C FP64 restatement of the same path   -> oracle.port

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (lightmetrica-v3_b200/) never does.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(os.path.dirname(_HERE), "lightmetrica-v3_b200")
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)
