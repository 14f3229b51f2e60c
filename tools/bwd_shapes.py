"""GPU: fp32 (and optionally bf16) backward timing on the shapes of bench.py's roofline_per_shape (same method: bench.time_shape).
usage: python tools/bwd_shapes.py [fp32|bf16] [fwd|bwd] [pdl: 0|1]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__  # noqa: F401  (sys.path)
import torch
import bench

dev = torch.device("cuda:0")
dname = sys.argv[1] if len(sys.argv) > 1 else "fp32"
which = sys.argv[2] if len(sys.argv) > 2 else "bwd"
if len(sys.argv) > 3:      # programmatic dependent launch of the Activation1d kernels on / off (afa_set_tuning(9, ...))
    from afa_b200 import _lib
    _lib.set_tuning(9, int(sys.argv[3]))
shapes = [(2, 512, 8192)]
for clips in (1, 8):
    shapes += [(b, c, t) for (b, c, t, _) in bench.stage_shapes(clips, bench.T_MEL_10S)]
shapes += [(32, c, mult * 32) for (c, mult, _) in bench.AMP_STAGES]
for (b, c, t) in shapes:
    row = bench.time_shape(dev, b, c, t, dname, which)
    print(json.dumps(row), flush=True)
