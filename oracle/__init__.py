"""TEST INFRASTRUCTURE ONLY — CPU oracle for toppra_b200 (see oracle/toppra_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product (toppra_b200/) never does."""
