#!/usr/bin/env python
"""Timing + roofline blocks of BASELINE.json configs 1-4 at full size on one B200 (the driver's headline bench
is bench.py = config 5, which embeds the same blocks).  Prints one JSON line per configuration.

    python tools/bench_configs.py [1] [2] [3] [4]
"""

from __future__ import annotations

import json
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
logging.disable(logging.INFO)

import torch  # noqa: E402

import cases  # noqa: E402

if __name__ == "__main__":
    dev = torch.device("cuda:0")
    peak = cases.fp64_peak_tflops(dev)
    for w in sys.argv[1:] or ["1", "2", "3", "4"]:
        fn = {"1": cases.config1, "2": cases.config2, "3": cases.config3, "4": cases.config4}[w]
        res = fn(dev) if w == "1" else fn(dev, peak=peak)
        print(json.dumps({"config": int(w), "fp64_peak_tflops_measured": peak, **res}), flush=True)
