"""Standalone corpus top-k timing (one GPU's shard of BASELINE C5): python profiles/retrieval_bench.py [nq nd k reps]"""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from dssm_b200 import retrieval as rt

nq, nd, k, reps = (int(x) for x in (sys.argv[1:5] + ["4096", "1250000", "100", "3"][len(sys.argv) - 1:]))
g = torch.Generator(device="cuda").manual_seed(0)
Q = torch.relu(torch.randn((nq, 128), generator=g, device="cuda"))
D = torch.relu(torch.randn((nd, 128), generator=g, device="cuda"))
rt.corpus_topk(Q, D, k, method="tc")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    rt.corpus_topk(Q, D, k, method="tc")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"tf32 filter over the fp32 rows: nq={nq} nd={nd} k={k}: {ms:.3f} ms  {nd / ms * 1e3 / 1e6:.1f} M docs/s  fallback={rt.LAST_CALL['fallback']}")
e0.record()
index = rt.CorpusIndex(D)
e1.record()
torch.cuda.synchronize()
print(f"index build: {e0.elapsed_time(e1):.3f} ms, {index.nbytes / 1e6:.0f} MB")
rt.corpus_topk(Q, index, k)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    rt.corpus_topk(Q, index, k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"bf16 filter over the corpus index: nq={nq} nd={nd} k={k}: {ms:.3f} ms  {nd / ms * 1e3 / 1e6:.1f} M docs/s  fallback={rt.LAST_CALL['fallback']}")
