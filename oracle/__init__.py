"""CPU oracle (test infrastructure only) -- see sif_oracle.py / mmb_oracle.py headers."""
