"""The training step around the hot path.

`TrainStep` composes one optimisation step the way `InterfacePhysics.run_train_interface` does
(interface/interface_physics.py:436-515 with the values of configs/DeepPhysiNet_NCEP_cfg.py:136-165):

    step <  pde_start_step (2000):  loss = margin data loss                         (:436-441, :471-474)
    step >= pde_start_step:         loss = margin + interior PDE + margin PDE       (:476-503)
    zero_grad, backward, clip_grad_norm_(max_norm=2.5e7), Adam(lr=1e-4, weight_decay=1e-4).step()   (:505-515)
    CosineAnnealingLR(T_max=5, eta_min=5e-6) stepped once per epoch, checkpoint per epoch           (:831-845)

Differences that do not change the result: the encoder and hyper-network run ONCE per step instead of three times
(`InterfacePhysics.training_losses`), and with several ranks the gradients are averaged by one flat NCCL all-reduce
(`parallel.FlatGradAllReduce`) before clipping - the DDP semantics of run_train_interface_dist (:899-907).
Optimiser, scheduler and checkpoint container are stock PyTorch, as in the reference (utils/optims/builder.py:5-26).
"""
import torch

from .config import DEFAULT_LOSS_FACTOR
from .parallel import FlatGradAllReduce


class TrainStep:
    def __init__(self, model, loss_factor=None, lr=1e-4, weight_decay=1e-4, pde_start_step=2000, max_grad_norm=2.5e7,
                 beta=0.1, t_max=5, eta_min=5e-6, global_step=0, last_epoch=-1):
        self.model = model
        self.loss_factor = dict(DEFAULT_LOSS_FACTOR, margin_factor=1.0e6) if loss_factor is None else loss_factor
        params = [{"params": model.physics_net.parameters(), "initial_lr": lr}]          # :394-395
        self.optimizer = torch.optim.Adam(params, lr=lr, weight_decay=weight_decay)
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.optimizer, T_max=t_max, eta_min=eta_min,
                                                                    last_epoch=last_epoch)
        self.reducer = FlatGradAllReduce(model.physics_net.parameters())
        self.pde_start_step, self.max_grad_norm, self.beta = pde_start_step, max_grad_norm, beta
        self.global_step = global_step

    def __call__(self, batch):
        """batch: device tensors as documented at InterfacePhysics.training_losses.  Returns the parts of the loss
        (tensors, no host synchronisation) plus the pre-clip gradient norm."""
        m = self.model
        with_pde = self.global_step >= self.pde_start_step                               # :436-441 (decided before the increment)
        m.with_clip = True
        self.global_step += 1
        total, parts = m.training_losses(batch, self.loss_factor, with_pde=with_pde, beta=self.beta)
        self.optimizer.zero_grad(set_to_none=True)
        total.backward()
        self.reducer()                                                                   # no-op for a single rank
        parts["grad_norm"] = torch.nn.utils.clip_grad_norm_(m.physics_net.parameters(), max_norm=self.max_grad_norm)
        self.optimizer.step()
        parts["train_loss"] = total.detach()
        return parts

    def end_epoch(self, epoch, checkpoint_path=None, **meta):
        """Once per epoch (:831-845): scheduler step, then the reference's checkpoint (save_model :53-62)."""
        self.scheduler.step()
        if checkpoint_path is not None:
            self.model.save_model(checkpoint_path, epoch, self.global_step, prefix="physics", dx=self.model.dx, dy=self.model.dy,
                                  dt=self.model.dt, pred_t_span=self.model.pred_t_span, obs_norm_cfg=self.model.obs_norm_cfg, **meta)
        return self.optimizer.param_groups[0]["lr"]

    @classmethod
    def resume(cls, model, checkpoint_path, **kw):
        """load_model (:64-88, :389-397): restores the weights and continues the epoch / step counters; the optimiser state is
        not part of a reference checkpoint, the scheduler restarts from last_epoch = epoch - 1."""
        state, epoch, step = model.load_model(checkpoint_path, prefix="physics")
        if state is not None:
            model.physics_net.load_state_dict(state["model"], strict=True)
        return cls(model, global_step=step, last_epoch=epoch - 1, **kw), epoch
