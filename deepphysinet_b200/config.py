"""Constants of the hot path, with the reference config as the single source of defaults
(configs/DeepPhysiNet_NCEP_cfg.py:64-83 obs_norm_cfg, :93-95 geometry, :139-148 loss factors;
physics constants from interface/interface_physics.py:126,146,177)."""
from dataclasses import dataclass, field
from typing import Dict, Tuple

VAR_ORDER = ("u10", "v10", "pres", "t2", "q2", "rio")          # interface_physics.py:256-261
NET_ATTRS = ("U_net", "V_net", "P_net", "T_net", "q_net", "rio_net")  # physics_net.py:49-54 (call order)
TERM_NAMES = ("montion_u_loss", "montion_v_loss", "continous_loss", "energy_loss", "vapor_loss", "gas_loss")
FACTOR_KEYS = ("motion_u_factor", "motion_v_factor", "continuous_factor", "energy_factor", "vapor_factor", "gas_factor")

DEFAULT_OBS_NORM = dict(
    u10=dict(norm_factor=[0.14507186950562942, 3.0050219075895894], bound=[-500, 500]),
    v10=dict(norm_factor=[-0.17325370241478535, 3.006602165591562], bound=[-500, 500]),
    pres=dict(norm_factor=[89741.36105771353, 13296.749084125422], bound=[10000, 500000]),
    t2=dict(norm_factor=[283.58054561520305, 15.583177935722373], bound=[50, 500]),
    q2=dict(norm_factor=[0.007909478276582905, 0.006304067969976075], bound=[1e-6, 10]),
    rio=dict(norm_factor=[1.0966503643401704, 0.15166081218127583], bound=[1e-6, 10]),
)
DEFAULT_LOSS_FACTOR = dict(motion_u_factor=1.0e3, motion_v_factor=1.0e3, continuous_factor=1.0e10,
                           energy_factor=1e1, vapor_factor=1.0e14, gas_factor=1.0e-7)


@dataclass
class PhysicsConsts:
    """Everything the residual kernel needs besides tensors; mirrored 1:1 by `DpnConsts` in include/dpn_b200.h."""
    dx: float = 27000.0
    dy: float = 27000.0
    lat_size: int = 145
    lon_size: int = 257
    pred_t_span: float = 86400.0
    with_clip: bool = True
    mean: Tuple[float, ...] = tuple(DEFAULT_OBS_NORM[k]["norm_factor"][0] for k in VAR_ORDER)
    std: Tuple[float, ...] = tuple(DEFAULT_OBS_NORM[k]["norm_factor"][1] for k in VAR_ORDER)
    lo: Tuple[float, ...] = tuple(float(DEFAULT_OBS_NORM[k]["bound"][0]) for k in VAR_ORDER)
    hi: Tuple[float, ...] = tuple(float(DEFAULT_OBS_NORM[k]["bound"][1]) for k in VAR_ORDER)
    factor: Tuple[float, ...] = tuple(DEFAULT_LOSS_FACTOR[k] for k in FACTOR_KEYS)
    c_p: float = 1005.0
    L: float = 2.5e6
    R_v: float = 461.5
    R_d: float = 287.0

    @staticmethod
    def from_cfg(obs_norm_cfg: Dict, loss_factor: Dict, **geom) -> "PhysicsConsts":
        for k in VAR_ORDER:
            c = obs_norm_cfg[k]
            if not c.get("use_norm", True) or c.get("norm_type", "mean_norm").lower() != "mean_norm":
                raise NotImplementedError("only mean_norm de-normalisation is on the hot path "
                                          "(interface_physics.py:250-251); got %r" % (c,))
        return PhysicsConsts(
            mean=tuple(float(obs_norm_cfg[k]["norm_factor"][0]) for k in VAR_ORDER),
            std=tuple(float(obs_norm_cfg[k]["norm_factor"][1]) for k in VAR_ORDER),
            lo=tuple(float(obs_norm_cfg[k]["bound"][0]) for k in VAR_ORDER),
            hi=tuple(float(obs_norm_cfg[k]["bound"][1]) for k in VAR_ORDER),
            factor=tuple(float(loss_factor[k]) for k in FACTOR_KEYS), **geom)
