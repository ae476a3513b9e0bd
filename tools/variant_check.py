"""Digest of golden cases under the environment the caller sets (a kernel variant), printed as JSON: the caller
compares it with the digest of the default build.  Run in a subprocess so that a trapping kernel cannot take the
test process down:   AM_B200_CLIP_MINB=4 python tools/variant_check.py chair mlp3x256s_cube"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from analyticmesh_b200 import cuam
from tests.golden.cases import build_case
from tests import parity

out = {}
for name in sys.argv[1:]:
    case = build_case(name)
    eng = parity.run_engine(case)
    d = cuam.digest()
    out[name] = {"raw": d["raw"], "faces": eng["stats"]["n_faces"], "seconds": eng["stats"]["seconds_march"],
                 "clip_s": eng["stats"]["seconds_clip"]}
cuam.Destroy()
print(json.dumps(out))
