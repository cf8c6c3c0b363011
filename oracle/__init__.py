"""CPU oracle for the DynEdge hot path -- test infrastructure only (see dynedge_oracle.py)."""
