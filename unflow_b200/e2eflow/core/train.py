"""Data-parallel training step: forward + loss + backward + gradient all-reduce + Adam.

Counterpart of the reference Trainer's hot loop (src/e2eflow/core/train.py:147-185 graph
construction, :222-251 ``sess.run([train_op, loss_])``, :388-422 ``average_gradients``) -- only
the step itself; the reference's session / checkpoint / summary / eval scaffolding is TF specific
and outside the hot path.

Design (one process per GPU):
  * every trainable variable is a view into ONE flat fp32 buffer, every gradient a view into a
    second one -> the tower-gradient mean of the reference (concat + reduce_mean on the CPU,
    train.py:388-422) is a single ``all_reduce`` over NCCL / NVLink on the flat gradient buffer,
    and the optimiser is a single fused Adam kernel (csrc/adam.cu) that also clears the gradients;
  * unlike the reference (whose towers all consume the same dequeued batch, train.py:169,191)
    each rank gets its own shard of the global batch; the averaged gradient equals the 1-GPU
    gradient of the concatenated batch;
  * learning-rate schedule: ``learning_rate`` halved every ``decay_interval`` iterations after
    ``decay_after`` (train.py:225-244).
"""
import torch
import torch.distributed as dist

from ... import _native
from ..._native import check
from .flownet import FlowNetVariables
from .unsupervised import unsupervised_loss


def learning_rate_at(decay_iters, params):
    """Host-side LR schedule of the reference training loop (train.py:225-244).

    ``decay_iters`` = number of iterations already done (the reference's
    ``local_i + iter_offset``, 0 for the first step)."""
    if 'manual_decay_lrs' in params and 'manual_decay_iters' in params:
        decay_index = 0
        iter_counter = 0
        for decay_i, manual_decay_iter in enumerate(params['manual_decay_iters']):
            iter_counter += manual_decay_iter
            if decay_iters <= iter_counter:
                decay_index = decay_i
                break
        return params['manual_decay_lrs'][decay_index]
    lr = params.get('learning_rate', 1.0e-4)
    decay_interval = params.get('decay_interval')
    if not decay_interval:
        return lr
    decay_after = params.get('decay_after', 0)
    if decay_iters >= decay_after:
        decay_minimum = decay_after / decay_interval
        decay = (decay_iters // decay_interval) - decay_minimum
        return lr / (2 ** decay)
    return lr


class Trainer:
    def __init__(self, params, normalization, device, variables=None, seed=1234,
                 loss_fn=unsupervised_loss, process_group=None):
        self.params = dict(params)
        self.normalization = normalization
        self.device = torch.device(device)
        self.loss_fn = loss_fn
        self.pg = process_group
        self.world_size = dist.get_world_size(process_group) if dist.is_initialized() else 1
        spec = self.params.get('flownet', 'S')
        if variables is None:
            variables = FlowNetVariables(spec, self.params.get('full_res'), seed=seed)
        self.variables = variables.to(self.device)
        # Only the final network of a stack is trained unless train_all (config.ini:55-58).
        n_nets = len(spec)
        if self.params.get('train_all') or n_nets == 1:
            scopes = list(self.variables.kinds)
        else:
            scopes = self.variables.scopes_of_net(n_nets - 1)
        trainable = []
        for sc in scopes:
            w, b = self.variables.weights(sc)
            trainable += [w, b]
        train_ids = {id(p) for p in trainable}
        for p in self.variables.parameters():
            p.requires_grad_(id(p) in train_ids)
        self.trainable = trainable
        self._flatten()
        self.iteration = 0

    def _flatten(self):
        n = sum(p.numel() for p in self.trainable)
        npad = (n + 3) // 4 * 4
        self.flat_param = torch.zeros(npad, device=self.device, dtype=torch.float32)
        self.flat_grad = torch.zeros(npad, device=self.device, dtype=torch.float32)
        self.adam_m = torch.zeros(npad, device=self.device, dtype=torch.float32)
        self.adam_v = torch.zeros(npad, device=self.device, dtype=torch.float32)
        off = 0
        with torch.no_grad():
            for p in self.trainable:
                k = p.numel()
                self.flat_param[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_param[off:off + k].view_as(p)
                p.grad = self.flat_grad[off:off + k].view_as(p)
                off += k
        self.num_params = n

    def broadcast_variables(self, src=0):
        if self.world_size > 1:
            dist.broadcast(self.flat_param, src, group=self.pg)

    def loss(self, im1, im2):
        return self.loss_fn((im1, im2), self.params, self.normalization, augment=False,
                            variables=self.variables)

    def reduce_gradients(self):
        """Mean of the gradients over ranks: ONE all-reduce on the flat buffer (NCCL over NVLink on
        GPUs).  Returns the factor still to be applied (folded into the Adam kernel)."""
        if self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.pg)
            return 1.0 / self.world_size
        return 1.0

    def apply_update(self, lr, grad_scale=1.0):
        if self.flat_param.device.type != "cuda":
            raise RuntimeError("the Adam update is a CUDA kernel (csrc/adam.cu); no CPU fallback")
        from ..ops import kernel_timer
        with torch.cuda.device(self.device), kernel_timer.span("adam", 32 * self.flat_param.numel()):
            check(_native.lib().unflow_adam_step(
                self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.adam_m.data_ptr(),
                self.adam_v.data_ptr(), self.flat_param.numel(), float(lr), 0.9, 0.999, 1e-8,
                self.iteration, float(grad_scale), 1, torch.cuda.current_stream().cuda_stream),
                "adam_step")

    def step(self, im1, im2, lr=None):
        """One optimisation step on this rank's shard; returns the (local) loss tensor."""
        self.iteration += 1
        loss = self.loss(im1, im2)
        loss.backward()   # accumulates into the flat gradient views
        scale = self.reduce_gradients()
        if lr is None:
            lr = learning_rate_at(self.iteration - 1, self.params)
        self.apply_update(lr, scale)
        return loss.detach()
