"""Stand-in for pytorch3d==0.7.1 `transforms` (test infrastructure only).
The arithmetic lives in oracle/eslam_oracle.py so there is exactly one restatement."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
from eslam_oracle import quaternion_to_matrix, matrix_to_quaternion  # noqa: E402,F401
