"""Generates tests/golden/cross_entropy.npz with torch.nn.functional.cross_entropy itself (CPU fp32) -- the call behind
nn.CrossEntropyLoss() at train_vit.py:81,102 and F.cross_entropy at train_videogpt.py:54: loss and d loss / d logits for
logits [R, C] with a few ignored rows (ignore_index = -100).

    python tests/golden/make_golden_ce.py
"""
import os

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    rng = np.random.default_rng(77)
    out = {}
    for tag, (R, C, n_ignored) in {"a": (37, 10, 0), "b": (12, 1000, 3), "c": (20, 1024, 4)}.items():
        x = (rng.standard_normal((R, C)) * 3.0).astype(np.float32)
        y = rng.integers(0, C, size=R).astype(np.int64)
        if n_ignored:
            y[rng.choice(R, n_ignored, replace=False)] = -100
        xt = torch.from_numpy(x).requires_grad_(True)
        loss = torch.nn.functional.cross_entropy(xt, torch.from_numpy(y))
        (loss * 1.5).backward()   # upstream gradient 1.5 (a GradScaler-like factor)
        out[f"{tag}_x"], out[f"{tag}_y"] = x, y
        out[f"{tag}_loss"] = loss.detach().numpy()
        out[f"{tag}_dx"] = xt.grad.numpy() / 1.5
    np.savez_compressed(os.path.join(OUT, "cross_entropy.npz"), **out)
    print("wrote cross_entropy.npz")


if __name__ == "__main__":
    main()
