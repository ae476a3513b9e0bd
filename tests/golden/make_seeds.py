"""Regenerates tests/golden/seeds_*.npz (run once, in this container; results are committed)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tests.golden.cases import CASES, build_case  # noqa: E402

if __name__ == "__main__":
    for name in CASES:
        c = build_case(name, regenerate=True)
        print(name, c["points"].shape, c["states"].shape, c["cube"])
