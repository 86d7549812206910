#!/usr/bin/env python
"""Usage: make_strat.py <distribution> <strategy> [<seed>]  -- scripts/make_strat.cpp on the GPU (see
deepgroebner_b200/strat.py): reads data/stats/<dist>/<dist>.csv, writes data/stats/<dist>/<dist>_<strategy>.csv."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepgroebner_b200.strat import make_strat  # noqa: E402

if __name__ == "__main__":
    if len(sys.argv) < 3:
        print("Usage: make_strat <distribution> <strategy> <seed>")
        sys.exit(1)
    code, msg = make_strat(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else None)
    if code:
        print(msg)
    sys.exit(code)
