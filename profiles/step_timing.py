"""Where does the single-GPU step time go?  (a) graph replays alone, (b) stage + replay as bench.py's `value` leg issues
them, (c) the CPU time Python needs to issue one iteration of (b).  Run: python profiles/step_timing.py [C2]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from dssm_b200 import DSSMTower, baseline_config
from dssm_b200.synthetic import init_params, make_batch

name = sys.argv[1] if len(sys.argv) > 1 else "C2"
conf = baseline_config(name)
batches = [make_batch(conf, seed=s) for s in range(4)]
t = DSSMTower(conf, max_nnz=max(b.nnz for b in batches), params=init_params(conf, 0))
dev = [t.to_device(b) for b in batches]
t.capture_graph()
N = 200
def timed(fn):
    for i in range(10): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0 = time.perf_counter(); e0.record()
    for i in range(N): fn(i)
    c_issue = time.perf_counter() - c0
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N * 1e3, c_issue / N * 1e6
t.stage(dev[0])
print(name, "replay only      : gpu %.1f us/step, cpu issue %.1f us/step" % timed(lambda i: t.train_step_staged()))
print(name, "stage + replay   : gpu %.1f us/step, cpu issue %.1f us/step" % timed(lambda i: (t.stage(dev[i % 4]), t.train_step_staged())))
print(name, "stage only       : gpu %.1f us/step, cpu issue %.1f us/step" % timed(lambda i: t.stage(dev[i % 4])))
print("launches/step", t.launch_count)
tl = t.profile_timeline()
print("timeline of one graphed step (globaltimer stamps between the calls of the main stream), us:")
print("  " + "  ".join(f"{k}={v * 1e3:.1f}" for k, v in tl), " | sum %.1f" % (sum(v for _, v in tl) * 1e3))
