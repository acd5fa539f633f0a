"""One tiny forward+backward of the drop-in Transformer on the GPU, checked against the numpy oracle
(called by __graft_entry__.smoke(); the oracle is only the checker)."""
import numpy as np
import torch


def run(dev="cuda:0"):
    from oracle import vit_oracle as O

    from . import modules as M
    torch.manual_seed(0)
    cfg = M.TransformerConfig(n_layers=2, n_heads=2, n_embd=128, block_size=40, causal=False, dropout=0.0)
    model = M.Transformer(cfg).to(dev)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 40, 128)).astype(np.float32)
    dy = rng.standard_normal((3, 40, 128)).astype(np.float32)
    xt = torch.from_numpy(x).to(dev).requires_grad_(True)
    y = model(xt)
    y.backward(torch.from_numpy(dy).to(dev))
    layers = []
    for layer in model.layers:
        sd = {k: v.detach().cpu().double().numpy() for k, v in layer.state_dict().items()}
        layers.append({"qkv_w": sd["multi_attn.qkv.weight"], "qkv_b": sd["multi_attn.qkv.bias"],
                       "fc1_w": sd["mlp.0.weight"], "fc1_b": sd["mlp.0.bias"],
                       "fc2_w": sd["mlp.2.weight"], "fc2_b": sd["mlp.2.bias"]})
    y_ref, caches = O.transformer_fwd(x.astype(np.float64), layers, 2, False)
    dx_ref, grads = O.transformer_bwd(dy.astype(np.float64), caches)

    def rel(a, b):
        return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))

    e_y = rel(y.detach().cpu().numpy(), y_ref)
    e_dx = rel(xt.grad.cpu().numpy(), dx_ref)
    e_w = rel(model.layers[0].mlp[0].weight.grad.cpu().numpy(), grads[0]["fc1_w"])
    assert e_y < 1e-2 and e_dx < 2e-2 and e_w < 2e-2, (e_y, e_dx, e_w)
    return e_y, e_dx, e_w
