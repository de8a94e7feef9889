"""TEST INFRASTRUCTURE: CPU oracle for the yagre-mcmc hot path (see yagre_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (yagre_mcmc_b200) never does.
"""
