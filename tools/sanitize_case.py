"""Dev tool: a small pass through every kernel family, to be run under compute-sanitizer (tools/sanitize.sh)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import helpers
import sibelia_b200 as sb
from oracle import restate

ctx = sb.Context(0)
st = helpers.strain_case(3, 40_000, seed=55)
for k in (25, 31, 40):
    got = ctx.enumerate(st, k)
    helpers.assert_tables_equal(got, restate.enumerate_bifurcations(st, k), "sanitize k=%d" % k)
chrs = [c.tobytes() for c in st]
op = [np.arange(len(c), dtype=np.uint32) for c in chrs]
out = ctx.simplify(chrs, op, 25, 150, 4)
print("sanitize_case: enumerate k=25/31/40 match the oracle; simplify stage collapsed %d bulges" % out[2])
ctx.close()
