"""dagl_b200 — B200-native (sm_100a) implementation of DAGL's dynamic attentive
graph block (reference: jianzhangcs/DAGL, */model/dagl.py classes CE / CES).

Public surface:
    CE, CES            drop-in nn.Modules (reference constructor / state_dict)
    patch_reference    swap reference CE instances inside a built network
    install            rebind model.dagl.CE before make_model(args)
    RR                 the whole network (dagl.py:10-54) from this package's modules; loads the reference checkpoints
    ResBlock, resblocks_forward, patch_resblocks
                       the ResBlock chains either side of the graph blocks (common.py:59-79) on the tcgen05 convolution
"""
from .ce import CE, CES, install, patch_reference  # noqa: F401
from .resblock import ResBlock, patch_resblocks, resblocks_forward  # noqa: F401
from .network import RR  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["CE", "CES", "RR", "install", "patch_reference", "ResBlock", "patch_resblocks", "resblocks_forward"]
