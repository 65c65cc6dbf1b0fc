"""CPU oracle for the SGRACE hot path -- TEST INFRASTRUCTURE ONLY (see oracle/sgrace_oracle.h)."""
