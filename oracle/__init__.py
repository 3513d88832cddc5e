"""CPU oracle -- test infrastructure only (see oracle/chaos_oracle.c header)."""
