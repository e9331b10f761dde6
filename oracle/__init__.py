"""CPU oracle package (test infrastructure only -- see oracle/rr_oracle.c header)."""
