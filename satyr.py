#!/usr/bin/env python3
"""`python satyr.py <model_config.yaml> <test_path> <iterations> [options]` -- the reference's predict command line
(reference src/satyr.py) on the B200 path; see pdp_solver_b200/satyr.py."""
import sys

from pdp_solver_b200.satyr import main

if __name__ == "__main__":
    sys.exit(main())
