"""Spread of loss / gradient norm per step over repeated EAGER runs and repeated GRAPH runs of the tiny trainer (same seeds, same
batches): is an eager-vs-graph difference of 2.5e-4 in a gradient norm inside the run-to-run spread of the eager path itself?"""
import os, sys
os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_unet_gpu as T

def run(graph, n=6):
    data = T._batches(n)
    tr = T._tiny_trainer()
    tr.setup_fit(gradient_clip_val=1.0, seed=1215, cuda_graph=graph, graph_warmup_steps=2)
    losses, norms = [], []
    for i, b in enumerate(data):
        out = tr.fit_step(b, i)
        losses.append(out["loss"].item())
        norms.append(float(tr._fit["opt"].last_norm[0]))
    return losses, norms

def spread(rows):
    out = []
    for col in zip(*rows):
        m = sum(col) / len(col)
        out.append(max(abs(v - m) for v in col) / abs(m))
    return out

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
E = [run(False) for _ in range(N)]
G = [run(True) for _ in range(N)]
print("eager  loss spread per step:", ["%.1e" % v for v in spread([e[0] for e in E])])
print("eager  norm spread per step:", ["%.1e" % v for v in spread([e[1] for e in E])])
print("graph  loss spread per step:", ["%.1e" % v for v in spread([g[0] for g in G])])
print("graph  norm spread per step:", ["%.1e" % v for v in spread([g[1] for g in G])])
print("all    norm spread per step:", ["%.1e" % v for v in spread([r[1] for r in E + G])])
print("all    loss spread per step:", ["%.1e" % v for v in spread([r[0] for r in E + G])])
rows = E + G
ref = rows[1]  # second eager run
for idx, r in enumerate(rows):
    dl = max(abs(a - b) / abs(b) for a, b in zip(r[0], ref[0]))
    dn = max(abs(a - b) / abs(b) for a, b in zip(r[1], ref[1]))
    if dl > 0 or dn > 1e-7:
        print(f"DEVIATES run {idx} ({'eager' if idx < N else 'graph'} #{idx % N}): loss {dl:.2e} norm {dn:.2e}")
