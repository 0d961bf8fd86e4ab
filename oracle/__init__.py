"""Test infrastructure only: CPU oracle for the UGS sampling path (see ugs_oracle.c header)."""
