"""CPU oracle for the ppca_rs hot path — TEST INFRASTRUCTURE ONLY (see ppca_oracle.c)."""
