#!/usr/bin/env python
"""cuBLAS FP64 yardsticks for the gLISA Hessian contraction (unit U2): the 1,536 x 65,536
panel product G^T G as DGEMM (torch.mm) -- library numbers to compare the hand-written SYRK against, not
product code.  Prints one JSON line."""
import json

import torch

dev = torch.device("cuda:0")
M, P = 1536, 65536
G = torch.randn(P, M, dtype=torch.float64, device=dev)
out = {}
for name, fn in (("dgemm_GtG", lambda: torch.mm(G.t(), G)),):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[name] = {"ms": ms, "tflops_full_2MMP": 2.0 * M * M * P / (ms * 1e-3) / 1e12,
                 "tflops_as_syrk_M(M+1)P": float(M) * (M + 1) * P / (ms * 1e-3) / 1e12}
print(json.dumps({"run": "cuBLAS FP64 yardstick", "M": M, "P": P, **out}))
