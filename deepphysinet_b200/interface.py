"""Drop-in `InterfacePhysics` for the hot path (reference: DeepPhysiNet/interface/interface_physics.py).

Keeps the constructor (:33-50), `place_one_batch` (:271-320, canonical argument order = the DDP call
site :1033-1035), `encoding_coord` (:322-332), `inverse_norm` (:232-262) and the checkpoint format of
`save_model` / `load_model` (:53-88).  The trainer loops, dataset, logging and visualisation of the
reference are out of scope.
"""
import os
import shutil

import torch
import torch.nn as nn

from . import functional as Fn
from .config import FACTOR_KEYS, TERM_NAMES, VAR_ORDER, PhysicsConsts
from .pe import SineCosPE
from .physics_net import PhysicsNet


class InterfacePhysics(nn.Module):
    def __init__(self, meta_cfg: dict, net_cfg: dict, obs_norm_cfg: dict, variable_cfg: dict = None,
                 train_cfg: dict = None, test_cfg=None, inference_cfg: dict = None, **kwargs):
        super().__init__()
        self.net_cfg, self.obs_norm_cfg, self.variable_cfg = net_cfg, obs_norm_cfg, variable_cfg
        self.train_cfg, self.test_cfg, self.inference_cfg = train_cfg or {}, test_cfg, inference_cfg
        self.physics_net = PhysicsNet(meta_cfg, net_cfg)
        self.pe = SineCosPE(3, include_input=False)
        img_size = self.train_cfg.get("img_size", (145, 257))
        if isinstance(img_size, (int, float)):
            self.lat_size, self.lon_size = img_size, img_size
        elif isinstance(img_size, (list, tuple)) and len(img_size) == 2:
            self.lat_size, self.lon_size = img_size
        else:
            raise NotImplementedError
        # set by run_train_interface in the reference (:339-342, :418, :438); given defaults from the config here
        self.dx = float(self.train_cfg.get("dx", 27000))
        self.dy = float(self.train_cfg.get("dy", 27000))
        self.dt = 3600.0
        self.pred_t_span = 86400
        self.with_clip = True
        self.mode = None            # None -> functional.default_mode()
        self.last_terms = None      # per-sample loss terms [B,6] (float64) of the last place_one_batch call

    # ---- checkpoints (:53-88) -------------------------------------------------------------------
    def save_model(self, checkpoint_path, epoch, global_step, prefix="physics", **kwargs):
        os.makedirs(checkpoint_path, exist_ok=True)
        checkpoint_file = os.path.join(checkpoint_path, "%s_%d.pth" % (prefix, epoch))
        state = {"model": self.physics_net.state_dict(), "epoch": epoch, "gobal_step": global_step}
        state.update(kwargs)
        torch.save(state, checkpoint_file)
        shutil.copy(checkpoint_file, os.path.join(checkpoint_path, "%s_latest.pth" % prefix))

    def load_model(self, checkpoint_path, current_epoch=None, prefix="downscale", map_location="cpu"):
        if os.path.isfile(checkpoint_path):
            model_file = checkpoint_path
        elif current_epoch is None:
            model_file = os.path.join(checkpoint_path, "%s_latest.pth" % prefix)
        else:
            model_file = os.path.join(checkpoint_path, "%s_%d.pth" % (prefix, current_epoch))
        if not os.path.exists(model_file):
            print("warning:%s does not exist!" % model_file)
            return None, 0, 0
        state = torch.load(model_file, map_location=map_location)
        step = state.pop("gobal_step", 0)
        epoch = state.pop("epoch", 0)
        return state, epoch + 1, step

    # ---- small PyTorch helpers kept for surface compatibility -------------------------------------
    def encoding_coord(self, x, y, t, pred_t_span):
        x = x / self.dx / (self.lon_size - 1)
        y = y / self.dy / (self.lat_size - 1)
        t = t / pred_t_span
        z = torch.stack([x, y, t], dim=1) if x.dim() == 1 else torch.cat([x, y, t], dim=1)
        return self.pe(z)

    def inverse_norm(self, u, v, P, T, q, rio, obs_norm_cfg, with_clip=False):
        outs = []
        for i, (data, key) in enumerate(zip((u, v, P, T, q, rio), VAR_ORDER)):
            cfg = obs_norm_cfg[key]
            if cfg.get("use_norm", True):
                if cfg.get("norm_type", "mean_norm").lower() != "mean_norm":
                    raise NotImplementedError("only mean_norm is supported on this path")
                data = data * cfg["norm_factor"][1] + cfg["norm_factor"][0]
                if i >= 2 and self.with_clip:      # :256-261 - the argument is shadowed by self.with_clip
                    data = torch.clip(data, cfg["bound"][0], cfg["bound"][1])
            outs.append(data)
        return tuple(outs)

    def consts(self, loss_factor) -> PhysicsConsts:
        return PhysicsConsts.from_cfg(self.obs_norm_cfg, loss_factor, dx=self.dx, dy=self.dy,
                                      lat_size=int(self.lat_size), lon_size=int(self.lon_size),
                                      pred_t_span=float(self.pred_t_span), with_clip=bool(self.with_clip))

    # ---- the hot path ---------------------------------------------------------------------------
    def place_one_batch(self, x, y, t, f, field_data, input_data, forecast_h, criterion, loss_factor,
                        global_step, local_rank, device, summary=None, prefix="inter", log_step=100):
        """Same contract as interface_physics.py:271-320: returns the scalar PDE loss (sum of the six weighted
        residual MSEs) whose .backward() reaches every physics_net parameter.  Host tensors are moved to
        `device` exactly as the reference does (:273-276); batches of samples ([B,N,...]) are allowed and
        give the mean over samples."""
        if not isinstance(criterion, nn.MSELoss) or getattr(criterion, "reduction", "mean") != "mean":
            raise NotImplementedError("the fused residual kernel implements nn.MSELoss() (losses/builder.py:10); got %r" % (criterion,))
        dev = torch.device(device)
        f, x, y, t = (a.to(dev, non_blocking=True) for a in (f, x, y, t))
        field_data = field_data.to(dev, non_blocking=True)
        input_data = input_data.to(dev, non_blocking=True)
        forecast_h = forecast_h.to(dev, non_blocking=True)
        B = field_data.shape[0]
        W = self.physics_net.decoder_weights(field_data, forecast_h)
        shp = (B, -1)
        total, terms = Fn.pde_residual(x.reshape(shp), y.reshape(shp), t.reshape(shp), f.reshape(shp),
                                       input_data.reshape(B, -1, 6), W, consts=self.consts(loss_factor), mode=self.mode)
        self.last_terms = terms
        if global_step % log_step == 1 and local_rank == 0:
            vals = terms.mean(dim=0).tolist()            # .item()-style sync, as in the reference's logging branch
            if summary is not None:
                summary.add_scalar("%s/total_loss" % prefix, float(sum(vals)), global_step)
                for name, v in zip(TERM_NAMES, vals):
                    summary.add_scalar("%s/%s" % (prefix, name), v, global_step)
            print("%s:" % prefix + ",".join("%s:%f" % (n, v) for n, v in zip(TERM_NAMES, vals)))
        return total.float()
