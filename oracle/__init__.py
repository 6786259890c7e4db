"""CPU oracle for the F8Net int_op_only path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package; the product (f8net_b200/) never does.  See oracle/f8_oracle.c.
"""
