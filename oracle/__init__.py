"""CPU oracle (test infrastructure only -- see oracle/wesup_ref.py header)."""
