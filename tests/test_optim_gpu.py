"""Fused AdamW (csrc/optim.cu, b200vit/optim.py) on a real B200: against trajectories of torch.optim.AdamW itself
(tests/golden/adamw.npz), against the numpy oracle (GradScaler protocol, device step counter), and the freshness of the
bf16 GEMM operands after a step of ANY optimiser (torch's fused optimisers do not bump Tensor._version).

Tolerance: the update is fp32 op-for-op the arithmetic of torch/optim/adam.py::_single_tensor_adam; differences are
fused-multiply-add contractions, i.e. a few ulp of the tensor's scale per step: parameters rtol 2e-6 / atol 1e-7*max|p|."""
import os

import numpy as np
import pytest
import torch

from oracle import adamw_oracle as A

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(got, ref, what, rtol=2e-6, scale_atol=2e-7):
    got = got.detach().float().cpu().numpy()
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=scale_atol * float(np.abs(ref).max()) + 1e-30, err_msg=what)


def _golden_run(g, capturable):
    from b200vit.optim import AdamW
    n, steps = int(g["n_tensors"]), int(g["steps"])
    params = [torch.nn.Parameter(torch.from_numpy(g[f"p0_{i}"]).to(DEV)) for i in range(n)]
    b1, b2 = (float(v) for v in g["betas"])
    opt = AdamW(params, lr=1e-3, betas=(b1, b2), eps=float(g["eps"]), weight_decay=float(g["weight_decay"]),
                capturable=capturable)
    for s in range(steps):
        for grp in opt.param_groups:
            grp["lr"] = float(g["lrs"][s])
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(g[f"g{s}_{i}"]).to(DEV)
        opt.step()
        for i, p in enumerate(params):
            _close(p, g[f"p{s + 1}_{i}"], f"param {i} after step {s + 1}")
    for i, p in enumerate(params):
        st = opt.state[p]
        assert float(st["step"]) == steps
        _close(st["exp_avg"], g[f"m{steps}_{i}"], f"exp_avg {i}", scale_atol=5e-7)
        _close(st["exp_avg_sq"], g[f"v{steps}_{i}"], f"exp_avg_sq {i}", scale_atol=5e-7)
    return opt, params


@pytest.mark.parametrize("capturable", [False, True])
def test_adamw_matches_torch_trajectory(golden_dir, capturable):
    g = np.load(os.path.join(golden_dir, "adamw.npz"))
    opt, params = _golden_run(g, capturable)
    # state_dict has torch.optim.AdamW's layout and round-trips
    sd = opt.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}
    assert {"lr", "betas", "eps", "weight_decay"} <= set(sd["param_groups"][0].keys())
    opt.load_state_dict(sd)
    ref = torch.optim.AdamW(params, lr=1e-3)
    ref.load_state_dict(sd)    # torch's own optimiser accepts it


def test_adamw_grad_scaler_protocol():
    from b200vit.optim import AdamW
    rng = np.random.default_rng(5)
    p0 = (rng.standard_normal(5000) * 0.1).astype(np.float32)
    g0 = rng.standard_normal(5000).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(p0).to(DEV))
    opt = AdamW([p], lr=2e-3, weight_decay=0.05)
    scale = 1024.0
    # step 1: an inf was found -> nothing changes, the counter stays at 0
    p.grad = torch.from_numpy(g0 * scale).to(DEV)
    opt.grad_scale = torch.tensor(scale, device=DEV)
    opt.found_inf = torch.tensor(1.0, device=DEV)
    opt.step()
    assert torch.equal(p.detach().cpu(), torch.from_numpy(p0))
    assert float(opt.state[p]["step"]) == 0.0
    # step 2: clean -> one oracle step with the unscaled gradient, counter = 1
    opt.found_inf = torch.tensor(0.0, device=DEV)
    opt.step()
    del opt.grad_scale, opt.found_inf
    z = np.zeros_like(p0)
    pr, mr, vr, took = A.adamw_step(p0, g0 * np.float32(scale), z, z, 1, 2e-3, weight_decay=0.05, grad_scale=scale)
    assert took and float(opt.state[p]["step"]) == 1.0
    _close(p, pr, "param after the clean step")
    _close(opt.state[p]["exp_avg"], mr, "exp_avg", scale_atol=5e-7)
    _close(opt.state[p]["exp_avg_sq"], vr, "exp_avg_sq", scale_atol=5e-7)
    # through torch.amp.GradScaler itself (the reference's scaler.step(optim), train_vit.py:105)
    scaler = torch.amp.GradScaler("cuda", init_scale=scale)
    q = torch.nn.Parameter(torch.from_numpy(p0).to(DEV))
    opt2 = AdamW([q], lr=2e-3, weight_decay=0.05)
    loss = (q * torch.from_numpy(g0).to(DEV)).sum()
    scaler.scale(loss).backward()
    scaler.step(opt2)
    scaler.update()
    pr2, _, _, _ = A.adamw_step(p0, g0, z, z, 1, 2e-3, weight_decay=0.05)
    _close(q, pr2, "param after GradScaler.step")


def test_adamw_rejects_cpu_parameters():
    from b200vit._cabi import B200VitError
    from b200vit.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(8))
    p.grad = torch.ones(8)
    with pytest.raises(B200VitError):
        AdamW([p]).step()


@pytest.mark.parametrize("which", ["torch-fused", "torch-foreach", "b200vit"])
def test_bf16_operands_follow_the_optimizer(which):
    """After an optimiser step the GEMMs must see the updated weights.  torch.optim.AdamW(fused=True) does not bump
    Tensor._version, which the operand cache used to key on alone (stale weights for ever); b200vit.optim.AdamW writes the
    refreshed operands itself."""
    from b200vit import functional as Fn
    from b200vit import modules as M
    from b200vit.optim import AdamW
    torch.manual_seed(0)
    cfg = M.TransformerConfig(n_layers=2, n_heads=2, n_embd=128, block_size=16)
    net = M.Transformer(cfg).to(DEV)
    if which == "b200vit":
        opt = AdamW(net.parameters(), lr=1e-2)
    else:
        opt = torch.optim.AdamW(net.parameters(), lr=1e-2, fused=(which == "torch-fused"), foreach=(which == "torch-foreach") or None)
    x = torch.randn(4, 16, 128, device=DEV)
    outs = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = net(x)
        outs.append(y.detach().clone())
        y.square().mean().backward()
        opt.step()
        for name, p in net.named_parameters():
            if p.dim() == 2:
                w16 = Fn.bf16_of(p)
                assert torch.equal(w16, p.detach().to(torch.bfloat16)), f"{which}: stale bf16 operand for {name}"
    assert not torch.equal(outs[0], outs[1]) and not torch.equal(outs[1], outs[2]), f"{which}: the forward did not see the updates"


def test_adamw_full_model_matches_torch():
    """ViT-Ti classifier, 3 steps: b200vit.optim.AdamW against torch.optim.AdamW (foreach) from the same initial state and
    the SAME gradients (fed from one backward), so that only the optimiser arithmetic is compared."""
    from b200vit import modules as M
    from b200vit.optim import AdamW
    M.transformer_configs.setdefault("Ti", lambda **kw: M.TransformerConfig(12, 3, 192, **kw))
    torch.manual_seed(0)
    net = M.ViTClassifier(M.ViTConfig(32, 3, 4, "Ti", 1, 0.0), num_classes=10).to(DEV)
    import copy
    twin = copy.deepcopy(net)
    ours = AdamW(net.parameters(), lr=1e-3, weight_decay=1e-2)
    ref = torch.optim.AdamW(twin.parameters(), lr=1e-3, weight_decay=1e-2, foreach=True)
    x = torch.randn(8, 3, 32, 32, device=DEV)
    yl = torch.randint(0, 10, (8,), device=DEV)
    for step in range(3):
        ours.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = torch.nn.functional.cross_entropy(net(x).float(), yl)
        loss.backward()
        for p, q in zip(net.parameters(), twin.parameters()):
            q.grad = p.grad.clone()
        ours.step()
        ref.step()
        for (name, p), q in zip(net.named_parameters(), twin.parameters()):
            _close(p, q.detach().cpu().numpy(), f"{name} after step {step + 1}", rtol=5e-6, scale_atol=5e-7)


def test_adamw_skips_parameters_without_gradients_and_handles_changing_sets():
    """torch semantics: parameters whose .grad is None are left alone; the set may change from step to step."""
    from b200vit.optim import AdamW
    rng = np.random.default_rng(8)
    a0, b0 = rng.standard_normal(300).astype(np.float32), rng.standard_normal((7, 9)).astype(np.float32)
    a, b = torch.nn.Parameter(torch.from_numpy(a0).to(DEV)), torch.nn.Parameter(torch.from_numpy(b0).to(DEV))
    opt = AdamW([a, b], lr=1e-2)
    ga = rng.standard_normal(300).astype(np.float32)
    a.grad = torch.from_numpy(ga).to(DEV)
    opt.step()                                            # b has no gradient
    assert torch.equal(b.detach().cpu(), torch.from_numpy(b0)) and len(opt.state[b]) == 0
    z = np.zeros_like(a0)
    a1, m1, v1, _ = A.adamw_step(a0, ga, z, z, 1, 1e-2)
    _close(a, a1, "a after step 1")
    gb = rng.standard_normal((7, 9)).astype(np.float32)
    b.grad = torch.from_numpy(gb).to(DEV)
    a.grad = None
    opt.step()                                            # now only b steps (its first step)
    _close(a, a1, "a untouched in step 2")
    b1, _, _, _ = A.adamw_step(b0, gb, np.zeros_like(b0), np.zeros_like(b0), 1, 1e-2)
    _close(b, b1, "b after its first step")
    assert float(opt.state[a]["step"]) == 1.0 and float(opt.state[b]["step"]) == 1.0


def test_adamw_refuses_graph_capture_without_capturable_and_keeps_host_steps_on_the_host():
    """(1) capturing step() with a host-side step counter would freeze the bias corrections into the graph: refused, like
    torch's _cuda_graph_capture_health_check.  (2) load_state_dict must not leave `step` on the device for non-capturable
    groups (one .item() sync per parameter and step otherwise) and the trajectory continues exactly."""
    from b200vit.graph import GraphedTrainStep
    from b200vit.optim import AdamW
    rng = np.random.default_rng(5)
    w0 = rng.standard_normal(1000).astype(np.float32)
    gs = [rng.standard_normal(1000).astype(np.float32) for _ in range(4)]

    def run(reload_after):
        p = torch.nn.Parameter(torch.from_numpy(w0).to(DEV))
        opt = AdamW([p], lr=1e-2)
        for i, g in enumerate(gs):
            p.grad = torch.from_numpy(g).to(DEV)
            opt.step()
            if i == reload_after:
                sd = opt.state_dict()
                opt = AdamW([p], lr=1e-2)
                opt.load_state_dict(sd)
                assert opt.state[p]["step"].device.type == "cpu" and float(opt.state[p]["step"]) == i + 1
        return p.detach().cpu().numpy()

    assert np.array_equal(run(None), run(1))
    p = torch.nn.Parameter(torch.from_numpy(w0).to(DEV))
    opt = AdamW([p], lr=1e-2)
    p.grad = torch.from_numpy(gs[0]).to(DEV)
    opt.step()
    graph = torch.cuda.CUDAGraph()
    with pytest.raises(RuntimeError, match="capturable=True"):
        with torch.cuda.graph(graph):
            opt.step()
    torch.cuda.synchronize()
    with pytest.raises(ValueError, match="capturable=True"):
        GraphedTrainStep(torch.nn.Linear(4, 4).to(DEV), opt, None, torch.zeros(1, 4, device=DEV), torch.zeros(1, device=DEV))
