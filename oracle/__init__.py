"""CPU oracle for the OQRL hot path (test infrastructure only, see pqc_oracle.py)."""
