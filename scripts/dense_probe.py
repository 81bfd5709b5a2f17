#!/usr/bin/env python
"""Config 1 (dense composite 4096 x 128 x 40) alone: `python scripts/dense_probe.py` prints bench.time_config1's
numbers; UCSA_DENSE_TMA=0 selects the load/compute forward kernel.  Meant to be run under ncu as well."""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, _ = bench.measured_peaks()
    out = bench.time_config1(dev, peak)
    print(json.dumps({k: out[k] for k in ("fwd", "bwd")}))
