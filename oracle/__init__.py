"""CPU oracle for the SMPL forward path -- test infrastructure only (see smpl_ref.py header)."""
