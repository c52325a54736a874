"""tagrec_b200 — B200-native (sm_100a) graph-embedding training + full-sort evaluation hot path of
chenzheng5555/tag-aware-recommendation, behind the reference's own Python interfaces.

Public surface (same names / call conventions as the reference):
    CFG / set_config / bind            utility/word.py CFG
    creat_adj, split_mm, CsrGraph      model/help/adj.py
    LightGCN, NGCF, TGCN, DGCF, DisenGCN, KGAT    model/*.py
    BPR_training_data                  train_data/bpr_training_data.py
    Basic_train, Basic_test, Early_stop   training/*.py
All device work goes through libtagrec_b200.so (include/tagrec_b200.h); there is no CPU fallback.
"""
from .config import CFG, bind, get_config, set_config          # noqa: F401
from . import _lib                                             # noqa: F401
from ._lib import TagrecError, launch_count                    # noqa: F401
from .adj import CsrGraph, build_csr, creat_adj, split_mm, spmm_raw, node_drop   # noqa: F401
from .lightgcn import LightGCN                                 # noqa: F401
from .ngcf import NGCF                                         # noqa: F401
from .dgcf import DGCF                                         # noqa: F401
from .disengcn import DisenGCN                                 # noqa: F401
from .tgcn import TGCN                                         # noqa: F401
from .kgat import KGAT, KGAT_training_data                     # noqa: F401
from . import routing                                          # noqa: F401
from .bpr_training_data import (Abstract_training_data, BPR_training_data, DGCF_training_data,   # noqa: F401
                                TransTag_training_data)
from .basic_train import Basic_train, epoch_training           # noqa: F401
from .basic_test import Basic_test                             # noqa: F401
from .early_stop import Early_stop                             # noqa: F401
from .optim import FusedAdam, ShardedFusedAdam, make_optimizer   # noqa: F401
from .graph_step import GraphedStep                            # noqa: F401
from . import data, distributed                                # noqa: F401
