"""World-size-2 `gloo` tests of the data-parallel path on CPU: two processes run the product's
`DMTrainer.fit_step` with the bucketed gradient exchange (uwudiff_b200/parallel.py); kernels are the test-only torch
emulation (tests/fake_ops.py), so what is under test is the host logic that also drives NCCL on the GPUs:

  * per-rank seeding (seed + rank -> different timesteps / noise per rank),
  * every gradient element is exchanged exactly once, in buckets released by the hand-scheduled backward,
  * the exchanged gradient is the MEAN of the ranks' local gradients,
  * replicas stay bit-identical after the optimizer step.
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_world(tmp_path, mode, port):
    outs = [str(tmp_path / f"rank{r}.pt") for r in range(2)]
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "ddp_worker.py"), str(r), "2", str(port), outs[r], mode],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    logs = []
    for p in procs:
        try:
            log, _ = p.communicate(timeout=280)
        except subprocess.TimeoutExpired:
            for q in procs:
                q.kill()
            raise
        logs.append(log)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)[-4000:]
    return [torch.load(o) for o in outs]


@pytest.mark.parametrize("mode,port", [("lycoris", 29631), ("full", 29632)])
def test_two_rank_gloo_step(tmp_path, mode, port):
    r0, r1 = run_world(tmp_path, mode, port)
    # different draws per rank (seed + rank)
    assert not torch.equal(r0["t"], r1["t"]) and r0["loss"] != r1["loss"]
    # every element exchanged exactly once, through more than one bucket
    assert r0["reduced_elems"] == r0["n"] and r0["n_calls"] > 1
    # local gradients differ, exchanged gradient = mean of the two, identical on both ranks
    assert not torch.equal(r0["local"], r1["local"]) and float(r0["local"].abs().sum()) > 0
    mean = (r0["local"] + r1["local"]) / 2
    assert torch.allclose(r0["reduced"], mean, rtol=1e-6, atol=1e-9)
    assert torch.equal(r0["reduced"], r1["reduced"])
    # replicas start identical although every rank seeded its model differently (rank 0's state is broadcast: trainable,
    # frozen and adapter tensors), move, and stay identical over several steps
    assert torch.equal(r0["before"], r1["before"]) and torch.equal(r0["frozen"], r1["frozen"])
    assert torch.equal(r0["after3"], r1["after3"]) and not torch.equal(r0["after3"], r0["after"])
    assert torch.equal(r0["after"], r1["after"]) and not torch.equal(r0["after"], r0["before"])


def test_two_rank_gloo_gradient_accumulation(tmp_path):
    """Strong-scaling mode (global batch fixed, `accumulate_grad_batches` micro-batches per rank): one exchange per window."""
    r0, r1 = run_world(tmp_path, "accum", 29633)
    for r in (r0, r1):
        assert r["calls_mb1"] == 0 and r["n_calls"] > 1               # nothing exchanged before the last micro-batch
        assert torch.equal(r["mid"], r["before"])                     # parameters untouched inside the window
        assert float(r["g_mb1"].abs().sum()) > 0                      # ... while the local gradient accumulates
        assert r["reduced_elems"] == r["n"] and r["global_step"] == 1
    assert not torch.equal(r0["local"], r1["local"])
    mean = (r0["local"] + r1["local"]) / 2
    assert torch.allclose(r0["reduced"], mean, rtol=1e-6, atol=1e-9) and torch.equal(r0["reduced"], r1["reduced"])
    assert torch.equal(r0["after"], r1["after"]) and not torch.equal(r0["after"], r0["before"])
