"""Configuration the hot path reads — same keys, defaults and per-model overlays as the reference's global ``CFG``
(utility/utils.py:18-62 argparse defaults, utility/config.py:1-81 overlays).

The reference parses ``sys.argv`` at import time into a module-level dict that every class reads in ``_config``
(e.g. model/lightgcn.py:25-35).  Here ``CFG`` is a plain dict the caller fills (``get_config``/``set_config``) or
binds to the reference's own dict (``bind(CFG_of_reference)``) so both code bases see the same object.
"""
import torch

_DEFAULTS = dict(
    model="lightgcn", data_root="data", dataset="synthetic",
    train_batch=512, test_batch=512, has_val=False, use_tag=True, patient_epoch=10, test_interval=5,
    early_stop_key="ndcg", topks=[10, 20], lr=0.01, reg=0.0, cor_reg=0.0, epochs=1000, dim_latent=64,
    dim_layer_list=[64, 32, 16], message_drop_list=[0.0, 0.0, 0.0], node_drop=0.0, seed=2020, cpu_core=4,
    split_adj_k=1,
    # new keys (no reference equivalent)
    sampler="device",        # "device": Philox sampler kernel | "mt19937": bit-exact numpy-legacy stream (cpu_core=1)
    eval_auc=True,           # compute the reference's per-user AUC (training/utils.py:37-45) on device
    eval_chunk=16384,        # users per K3 launch (the reference's test_batch bounds a B x n_item matrix we never build)
    eval_shard=True,         # torch.distributed initialised: shard evaluation users over the ranks
)

# utility/config.py:1-81
MODEL_OVERLAY = {
    "ngcf": {"norm_type": "ngcf", "agg_type": "bi_agg", "mul_loss_func": "logsigmoid"},
    "lightgcn": {"mul_loss_func": "softplus", "norm_type": "bi_norm", "cor_batch": 100},
    "dgcf": {"mul_loss_func": "softplus", "norm_type": "plain", "factor_k": 4, "iterate_k": 2, "cor_batch": 100},
    "disengcn": {"mul_loss_func": "softplus", "norm_type": "plain", "factor_k": 4, "iterate_k": 2, "cor_batch": 100},
    "kgat": {"dim_relation": 64, "transe_reg": 0.0001, "transe_batch": 1024, "agg_type": "bi_agg",
             "mul_loss_func": "softplus"},
    "tgcn": {"dim_weight": 10, "dim_atten": 32, "num_bit_conv": 32, "num_vec_conv": 8, "margin": 1,
             "transtag_batch": 512, "neighbor_k": 25, "transtag_reg": 0.0001, "mul_loss_func": "logsigmoid"},
}

CFG = {}


def get_config(model="lightgcn", **overrides):
    """utility/utils.py:50-62 get_config: defaults, device pick, then the per-model overlay, then overrides."""
    cfg = dict(_DEFAULTS)
    cfg["model"] = model
    cfg["device"] = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    cfg.update(MODEL_OVERLAY.get(model, {}))
    cfg.update(overrides)
    return cfg


def set_config(model="lightgcn", **overrides):
    """Fill the module-level CFG in place (objects hold a reference to this dict, like the reference's classes)."""
    CFG.clear()
    CFG.update(get_config(model, **overrides))
    return CFG


def bind(external_cfg):
    """Use the reference's own CFG dict (``from utility.word import CFG``) as this package's configuration."""
    global CFG
    for k, v in _DEFAULTS.items():
        external_cfg.setdefault(k, v)
    CFG = external_cfg
    import sys
    pkg = sys.modules.get("tagrec_b200")
    if pkg is not None:
        pkg.CFG = external_cfg
    return CFG


def current():
    import sys
    pkg = sys.modules.get("tagrec_b200")
    return getattr(pkg, "CFG", CFG) if pkg is not None else CFG


set_config("lightgcn")
