"""Probe: how much of a step is the torch.unique sync + prologue? (monkeypatches torch.unique with a cached result)"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from frameino_b200 import synth
dev = torch.device("cuda", 0)
cfg = dict(synth.WAN22_5B)
cfg["num_layers"] = int(os.environ.get("LAYERS", "4"))
model = synth.build_wan_on_device(cfg, seed=0, device=dev)
hidden, ts, text = synth.make_wan_inputs(cfg, 31, 44, 80, n_id=1, text_len=512, text_true_len=120, dtype=torch.bfloat16)
inp = dict(hidden_states=hidden.to(dev), timestep=ts.to(dev), encoder_hidden_states=text.to(dev), return_dict=False)
def run(steps=10):
    for _ in range(3): model(**inp)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for _ in range(steps): model(**inp)
    e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return s.elapsed_time(e) / steps, (t1 - t0) * 1e3 / steps
print("with torch.unique: gpu %.3f ms/step, cpu-issue %.3f ms/step" % run())
real = torch.unique
cached = real(ts.to(dev).reshape(-1).float(), return_inverse=True)
torch.unique = lambda *a, **k: cached
print("cached unique:     gpu %.3f ms/step, cpu-issue %.3f ms/step" % run())
torch.unique = real
