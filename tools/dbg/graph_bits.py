"""Debug: where does the CUDA-graph step stop being bit-identical to the eager step (tiny UNet + LyCORIS)?"""
import os, sys
os.environ.setdefault("UWU_SYNTHETIC_CONDITIONING", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_unet_gpu as T

def run(graph, manual_from=2, n=5):
    data = T._batches(n)
    tr = T._tiny_trainer()
    tr.setup_fit(gradient_clip_val=1.0, seed=1215, cuda_graph=graph, graph_warmup_steps=2)
    f = tr._fit
    rec = []
    for i, b in enumerate(data):
        if i < manual_from:
            out = tr.fit_step(b, i)
            rec.append(dict(loss=out["loss"].item(), params=tr.lycoris_model.flat_params.clone()))
            continue
        if not graph:
            out = tr.training_step(b, i)
            out["loss"].backward()
            grads = tr.lycoris_model.flat_grads.clone()
            f["opt"].step()
            norm = f["opt"].last_norm.clone()
            if f["sched"] is not None: f["sched"].step()
            tr.lycoris_model.zero_grad()
            tr.global_step += 1
        else:
            g = f["graph"]
            prepared = tr.get_latent_and_conditioning(b)
            if g["state"] == "capture":
                tr._graph_capture(prepared)
            x, ctx, _m, added, _c = prepared
            g["x"].copy_(x)
            if ctx is not None: g["ctx"].copy_(ctx)
            for kk, v in added.items():
                if torch.is_tensor(v): g["added"][kk].copy_(v)
            tr._write_hyper()
            g["fwdbwd"].replay()
            grads = tr.lycoris_model.flat_grads.clone()
            g["optstep"].replay()
            norm = f["opt"].last_norm.clone()
            f["opt"]._step += 1
            if f["sched"] is not None: f["sched"].step()
            tr.global_step += 1
            out = g["out"]
        rec.append(dict(loss=out["loss"].item(), grads=grads, norm=norm, params=tr.lycoris_model.flat_params.clone(),
                        hyper=None if not graph else tr._hyper.clone()))
    return tr, rec

def cmp(tag, ra, rb, tr):
    for i, (a, b) in enumerate(zip(ra, rb)):
        line = f"[{tag}] step {i}: loss {a['loss']!r} {b['loss']!r} eq={a['loss']==b['loss']}"
        if "grads" in a and "grads" in b:
            d = (a["grads"] != b["grads"])
            line += f" grads_diff={int(d.sum())}/{d.numel()} maxabs={(a['grads']-b['grads']).abs().max().item():.3e} norm {a['norm'].tolist()} {b['norm'].tolist()}"
        d = (a["params"] != b["params"])
        line += f" params_diff={int(d.sum())} maxabs={(a['params']-b['params']).abs().max().item():.3e}"
        print(line)
        if "grads" in a and "grads" in b and int((a["grads"] != b["grads"]).sum()):
            # which adapters
            lm = tr.lycoris_model
            off = 0
            bad = []
            for name, p in lm.named_parameters() if hasattr(lm, "named_parameters") else []:
                pass
    sys.stdout.flush()

trA, ra = run(False)
trA2, ra2 = run(False)
trB, rb = run(True)
trB2, rb2 = run(True)
cmp("eager-vs-eager", ra, ra2, trA)
cmp("eager-vs-graph", ra, rb, trA)
cmp("graph-vs-graph", rb, rb2, trA)
print("hyper graph", [r["hyper"].tolist() for r in rb if r.get("hyper") is not None][:2])
o = trA._fit["opt"]
print("eager hyper_values now", o.hyper_values(0), "step", o._step)
