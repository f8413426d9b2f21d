"""CPU oracle package (TEST INFRASTRUCTURE ONLY -- see bn254_oracle.c header; parity unpinned)."""
