"""stc_gnn_b200 -- B200-native (sm_100a) implementation of STC-GNN's graph-convolution GRU cell.

Public surface:
  STC_Cell            drop-in for the reference class (framework/STC_GNN.py:51-79)
  stc_cell_forward    functional form (differentiable)
  CsrSupport          constant sparse spatial support
  support_apply       the spatial mode product on its own
  install / run_main  rebind STC_GNN.STC_Cell so Model_Trainer.py / Main.py run unchanged
  dp                  batch data-parallel gradient bucket (one flat all-reduce per step)
  halo                row-partitioned CSR support with per-hop halo exchange
  GraphedStep         CUDA-graph capture/replay of one forward+backward step (launch-bound small batches)
"""
from .cell import STC_Cell, GraphConvParams, stc_cell_forward  # noqa: F401
from .support import CsrSupport, support_apply  # noqa: F401
from .install import install, run_main  # noqa: F401
from .stack import RecurrentStack  # noqa: F401
from .graph import GraphedStep  # noqa: F401
from . import dp, halo, mgp  # noqa: F401
