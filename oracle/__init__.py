"""CPU oracle for the BPR / MF / score+rank hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under cornac_b200/ may import this package (tests/test_no_oracle_in_product.py
enforces it).  See oracle/cornac_oracle.c for the parity-pinning statement.
"""
