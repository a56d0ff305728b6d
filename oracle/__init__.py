"""CPU oracle (test infrastructure only; see ce_oracle.py header)."""
