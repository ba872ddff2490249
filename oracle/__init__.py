"""CPU oracle for pgtg_b200 -- TEST INFRASTRUCTURE ONLY (see oracle/pgtg_oracle.c)."""
