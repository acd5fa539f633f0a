"""Generates tests/golden/adamw.npz by running torch.optim.AdamW ITSELF (CPU, single-tensor implementation: the code
the reference's `torch.optim.AdamW(...)` + `optim.step()` executes, train_vit.py:82,105) on seeded tensors.

    python tests/golden/make_golden_adamw.py

Three tensors of awkward sizes (a scalar-ish bias, a matrix that is not a multiple of the kernel's 4096-element chunk,
one spanning two chunks plus a ragged tail), 4 steps, the learning rate changed between steps the way an LR scheduler does, fresh
gradients every step.  Stored: initial parameters, every step's gradients and learning rate, parameters after every step,
both moments after the last."""
import os

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
SHAPES = [(37,), (64, 48), (8200,)]
STEPS = 4


def main():
    rng = np.random.default_rng(20240607)
    params = [torch.nn.Parameter(torch.from_numpy((rng.standard_normal(s) * 0.05).astype(np.float32))) for s in SHAPES]
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, foreach=False, fused=False)
    out = {"n_tensors": np.int64(len(SHAPES)), "steps": np.int64(STEPS), "betas": np.array([0.9, 0.999]),
           "eps": np.float64(1e-8), "weight_decay": np.float64(1e-2)}
    for i, p in enumerate(params):
        out[f"p0_{i}"] = p.detach().numpy().copy()
    lrs = []
    for s in range(STEPS):
        lr = 1e-3 * (0.5 + 0.5 * np.cos(0.7 * s))   # scheduler-like
        lrs.append(lr)
        for grp in opt.param_groups:
            grp["lr"] = float(lr)
        for i, p in enumerate(params):
            scale = 10.0 ** rng.uniform(-4, 0)      # gradients of very different magnitudes
            g = (rng.standard_normal(p.shape) * scale).astype(np.float32)
            p.grad = torch.from_numpy(g.copy())
            out[f"g{s}_{i}"] = g
        opt.step()
        for i, p in enumerate(params):
            out[f"p{s + 1}_{i}"] = p.detach().numpy().copy()
            if s == STEPS - 1:
                out[f"m{s + 1}_{i}"] = opt.state[p]["exp_avg"].numpy().copy()
                out[f"v{s + 1}_{i}"] = opt.state[p]["exp_avg_sq"].numpy().copy()
    out["lrs"] = np.array(lrs, dtype=np.float64)
    out["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, "adamw.npz"), **out)
    print("wrote adamw.npz", {k: v.shape for k, v in out.items() if k.startswith("p0_")})


if __name__ == "__main__":
    main()
