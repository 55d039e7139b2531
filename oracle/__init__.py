"""CPU oracle for the nufhe bootstrap hot path -- TEST INFRASTRUCTURE ONLY (see nufhe_oracle.c)."""
