"""Data-parallel plumbing for the ray batch.

Rays are independent, so each rank renders a contiguous shard and the only exchange per training step is
ONE all-reduce (sum) of the flat fp32 gradient buffer (~2.6 MB for the core config) over NCCL.  The
reference does the same thing one level up with DistributedDataParallel buckets over every registered
parameter (nnutils/train_utils.py:48-62, 98-106, 958)."""
import torch
import torch.distributed as dist


class FlatParams:
    """Re-homes a list of leaf tensors into one flat parameter buffer and one flat gradient buffer.

    After construction ``p.data`` and ``p.grad`` of every tensor are views into ``self.flat`` /
    ``self.grad``; autograd accumulates into the views in place, so ``allreduce()`` is a single collective
    and an optimizer can be built over the single tensor ``self.flat`` (its .grad is ``self.grad``)."""

    ALIGN = 64  # fp32 elements

    def __init__(self, tensors):
        self.tensors = [t for t in tensors]
        # every tensor starts on a 256-byte boundary: the kernels read weights, biases and head vectors with
        # 128-bit loads and TMA, exactly as they may for separately allocated nn.Parameters (the padding stays
        # zero in both buffers, so it is inert under the all-reduce and the optimizer)
        offs, n = [], 0
        for t in self.tensors:
            offs.append(n)
            n += (t.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        dev = self.tensors[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(n, device=dev, dtype=torch.float32)
        for t, off in zip(self.tensors, offs):
            k = t.numel()
            self.flat[off:off + k].copy_(t.detach().reshape(-1))
            t.data = self.flat[off:off + k].view(t.shape)
            t.grad = self.grad[off:off + k].view(t.shape)
            # the fused MLP backward (chain_tc._grad_targets) accumulates straight into this view instead of
            # handing autograd a temporary; consequence: torch.autograd.grad() on a re-homed tensor sees None
            t._moda_grad_inplace = True
        self.offsets = offs
        self.flat.requires_grad_(True)
        self.flat.grad = self.grad
        self.numel = n

    def zero_grad(self):
        """Zeroes the flat gradient buffer in place.  Do NOT use ``optimizer.zero_grad()`` (set_to_none=True drops the
        views the kernels accumulate into): ``check()`` below catches that instead of silently training on stale
        gradients."""
        self.grad.zero_()

    optimizer_zero_grad = zero_grad

    def check(self):
        """The fused MLP backward adds straight into ``t.grad`` (a view of ``self.grad``) and hands autograd None, which
        is only correct while those views are intact.  Raises if something (optimizer.zero_grad(set_to_none=True),
        module.zero_grad(), a fresh .grad assignment) replaced them."""
        if self.flat.grad is not self.grad:
            raise RuntimeError("FlatParams: flat.grad was replaced (optimizer.zero_grad(set_to_none=True)?); use "
                               "FlatParams.zero_grad() so that the parameter .grad views stay attached")
        lo, hi = self.grad.data_ptr(), self.grad.data_ptr() + 4 * self.numel
        for t, off in zip(self.tensors, self.offsets):
            g = t.grad
            if g is None or g.data_ptr() != lo + 4 * off or not (lo <= g.data_ptr() < hi):
                raise RuntimeError("FlatParams: a parameter's .grad no longer aliases the flat gradient buffer "
                                   "(module.zero_grad() / set_to_none?); use FlatParams.zero_grad()")

    def adamw_step(self, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        """One AdamW step on the flat buffers (``moda_adamw_flat``: the arithmetic of torch.optim.AdamW, defaults included),
        graph-capturable: moments and the step counter live on the device.  Replaces an optimizer built over ``self.flat``
        (whose multi-tensor kernel covers one 650k-element tensor with ~10 CTAs)."""
        from ._lib import call, ptr, stream
        self.check()
        if not hasattr(self, "exp_avg"):
            self.exp_avg = torch.zeros_like(self.grad)
            self.exp_avg_sq = torch.zeros_like(self.grad)
            self.opt_state = torch.zeros(3, device=self.grad.device, dtype=torch.float32)
        with torch.no_grad():
            call("moda_adamw_flat", ptr(self.flat), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq), self.numel,
                 ptr(self.opt_state), float(lr), float(betas[0]), float(betas[1]), float(eps), float(weight_decay), stream())

    def allreduce(self, scale=None):
        """Sum of the per-rank gradients (each rank scales its loss by its share of the global batch)."""
        self.check()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
        if scale is not None:
            self.grad.mul_(scale)


def shard_rays(rays, rank, world):
    """Contiguous N/world slice of every per-ray tensor (the rays dict of geom_utils.py:785-794)."""
    out = {}
    for k, v in rays.items():
        if torch.is_tensor(v) and v.dim() >= 1:
            n = v.shape[0]
            per = (n + world - 1) // world
            out[k] = v[rank * per:min(n, (rank + 1) * per)]
        else:
            out[k] = v
    return out
