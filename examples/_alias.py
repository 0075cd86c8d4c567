"""Every example runs the reference's user-facing calls unchanged; this makes `ppca_rs` resolve to the B200 engine."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ppca_rs_b200.compat as compat  # noqa: E402

compat.install_as_ppca_rs()
