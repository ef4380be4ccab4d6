#!/usr/bin/env python
"""probe_denoise.py — DenoiseCompositor timing alone (bench.py's measure_denoise), for A/B runs with RT_CORE_LIB=build/rt_*.so."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from dxrexperiments_b200 import rtcore as rt  # noqa: E402

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx = rt.Context(0, stream=stream.cuda_stream)
d = bench.measure_denoise(None, ctx, rt, torch, stream)
print(json.dumps({"probe": "denoise", "ms": d["ms"], "frac": d["roofline"]["frac"], "lib": os.environ.get("RT_CORE_LIB", "default")}))
