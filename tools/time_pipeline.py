#!/usr/bin/env python3
"""Host-buffer pipeline tuning: comparisons/s of LocalPipeline.compare for several (contexts, chunks) settings (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from protocols.secure_comparison_b200 import SchemeEngine
from protocols.secure_comparison_b200.keyfile import schemes_from_json
from protocols.secure_comparison_b200.pipeline import LocalPipeline

B, l = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 16
paillier, dgk = schemes_from_json(bench.key_file(2048, 2048, 16))
eng = SchemeEngine(0); eng.load_paillier(paillier); eng.load_dgk(dgk, l)
w, truth = bench.make_workload(B, l, 7, eng)
h = {k: torch.from_numpy(v.view(np.int32) if v.dtype == np.uint32 else v).pin_memory().numpy() for k, v in w.items()}
h = {k: (v.view(np.uint32) if v.dtype == np.int32 else v) for k, v in h.items()}
eng.close()
for nctx, chunks in ((1, 1), (2, 2), (2, 4), (2, 8), (3, 6), (3, 9), (4, 8)):
    pipe = LocalPipeline(paillier, dgk, l, n_contexts=nctx, chunks=chunks)
    res = pipe.arena.empty((B, pipe.engines[0].kc))
    pipe.compare(h, res)
    t0 = time.perf_counter()
    for _ in range(2):
        pipe.compare(h, res)
    dt = (time.perf_counter() - t0) / 2
    print(f"contexts={nctx} chunks={chunks}: {B / dt:9.1f} comparisons/s  ({dt * 1e3:.1f} ms)", flush=True)
    pipe.close()
