"""ORACLE — test infrastructure only (never imported by the product path).

numpy fp32 restatement of the optimiser step the reference scripts run right after the hot path:
`torch.optim.AdamW(model.parameters(), lr=..., weight_decay=...)` (train_vit.py:82, train_titok.py:134,
train_vit_vqgan.py:131, train_videogpt.py:107) stepped through `scaler.step(optim)` (train_vit.py:105).

The algorithm lives in a third-party dependency that is not under /root/reference: torch 2.11.0 (the installed
version; the reference pins none), torch/optim/adam.py::_single_tensor_adam with decoupled_weight_decay=True
(what torch.optim.AdamW dispatches to).  Restated in its operation order, every tensor op rounded to float32 and the
Python-scalar arithmetic (bias corrections, step size) in double precision exactly as there:

    param.mul_(1 - lr * weight_decay)
    exp_avg.lerp_(grad, 1 - beta1)
    exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bias_correction1 = 1 - beta1 ** step ; bias_correction2 = 1 - beta2 ** step
    step_size = lr / bias_correction1 ; bias_correction2_sqrt = bias_correction2 ** 0.5
    denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    param.addcdiv_(exp_avg, denom, value=-step_size)

GradScaler protocol (torch/amp/grad_scaler.py, optimisers with _step_supports_amp_scaling): grad /= grad_scale;
if found_inf != 0 the step is skipped entirely (the step counter does not advance).

Pinning: tests/golden/adamw.npz holds trajectories produced by torch.optim.AdamW itself (CPU, single-tensor
implementation) by tests/golden/make_golden_adamw.py; tests/test_oracle_golden.py::test_adamw_oracle checks this file
against them.  Only tests/ may import this module.
"""
import numpy as np

F = np.float32


def adamw_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2, grad_scale=None,
               found_inf=0.0):
    """One step on fp32 arrays (updated copies are returned); `step` is the 1-based number of THIS step.
    Returns (p, m, v, step_taken) with step_taken False when found_inf skipped it."""
    if found_inf:
        return p.copy(), m.copy(), v.copy(), False
    p, g, m, v = (np.asarray(a, dtype=F) for a in (p, g, m, v))
    if grad_scale is not None:
        g = (g * (F(1.0) / F(grad_scale))).astype(F)
    p = (p * F(1.0 - lr * weight_decay)).astype(F)
    m = (m + (g - m).astype(F) * F(1.0 - beta1)).astype(F)
    v = ((v * F(beta2)).astype(F) + ((F(1.0 - beta2) * g).astype(F) * g).astype(F)).astype(F)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = F(lr / bc1)
    denom = (np.sqrt(v).astype(F) / F(bc2 ** 0.5)).astype(F) + F(eps)
    p = (p - step_size * (m / denom).astype(F)).astype(F)
    return p, m, v, True
