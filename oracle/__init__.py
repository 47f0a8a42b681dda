"""CPU oracle for the RADEX/pyradex hot path -- TEST INFRASTRUCTURE ONLY (see radex_oracle.c)."""
