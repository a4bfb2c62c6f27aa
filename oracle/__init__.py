"""TEST INFRASTRUCTURE.  CPU oracle for the continuous-HMM hot path (see hmm_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package never does.
"""
