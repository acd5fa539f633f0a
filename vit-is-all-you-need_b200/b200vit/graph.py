"""CUDA-graph capture of a whole training step (forward + loss + backward + optimizer) over the drop-in modules.

Small configurations (BASELINE.json configs[0]: ViT-Ti, 32 px, batch 32) are launch-bound: ~350 kernel launches of a few
microseconds each per step, issued from Python.  Every kernel of this library takes raw pointers and the current stream,
allocates nothing and never synchronises, so a step can be captured once and replayed:

    step = GraphedTrainStep(model, optimizer, loss_fn, example_inputs, example_targets)
    loss = step(images, labels)        # copies into the static buffers, replays the graph, returns the loss tensor

Constraints (the usual CUDA-graph ones): fixed shapes; the optimizer must be capturable (b200vit.optim.AdamW(...,
capturable=True) or torch.optim.AdamW(..., capturable=True)); dropout masks are functions of host-generated seeds that get baked
into the graph, so with dropout > 0 a replay repeats the captured masks -- capture is therefore refused for dropout > 0;
data-parallel wrappers are not captured (their bucket bookkeeping lives on the host).
"""
import torch


class GraphedTrainStep:
    def __init__(self, model, optimizer, loss_fn, example_inputs, example_targets, autocast_dtype=torch.bfloat16,
                 warmup_steps=3):
        for m in model.modules():
            if float(getattr(m, "dropout", 0.0) or 0.0) > 0.0 and hasattr(m, "n_embd"):
                raise ValueError("GraphedTrainStep: dropout > 0 would replay the captured masks; use eager steps")
        for group in optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("GraphedTrainStep: the optimizer must be built with capturable=True (a host-side step counter "
                                 "would be frozen into the graph)")
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.autocast_dtype = autocast_dtype
        self.static_x = example_inputs.clone()
        self.static_y = example_targets.clone()
        self.static_loss = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):       # warm-up on a side stream (allocator pools, lazy initialisation, optimizer state)
            for _ in range(warmup_steps):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._invalidate_weight_cache()      # the bf16 weight casts must be part of the captured work
        self.graph = torch.cuda.CUDAGraph()
        self.optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.static_loss = self._eager_step(zero=False)
        # captured gradients live in the graph's private pool; later zero_grad(set_to_none=True) calls would drop them
        self._captured_grads = [p.grad for p in model.parameters()]

    def _invalidate_weight_cache(self):
        from . import functional as Fn
        Fn.invalidate_weight_cache(self.model)   # bf16 operands, LayerNorm-folded operands, de-patchify operands

    def _eager_step(self, zero=True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=self.autocast_dtype):
            loss = self.loss_fn(self.model(self.static_x).float(), self.static_y)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, inputs, targets):
        self.static_x.copy_(inputs, non_blocking=True)
        self.static_y.copy_(targets, non_blocking=True)
        self.graph.replay()
        return self.static_loss


__all__ = ["GraphedTrainStep"]
