"""Drop-in replacement for the reference's top-level ``STMGCN`` module (import name fixed by ``Main.py:5``).

Exports ``CG_LSTM`` (reference ``STMGCN.py:7-57``) and ``ST_MGCN`` (reference ``STMGCN.py:61-119``); class
names matter (``Model_Trainer.py:11,34`` dispatches on ``model.__class__.__name__ == 'ST_MGCN'``).
Implementation: ``st-mgcn_b200/stmgcn_b200/modules.py`` over ``libstmgcn_b200.so``.
"""
from GCN import GCN                                       # noqa: F401  (same import the reference makes, STMGCN.py:3)
from stmgcn_b200.modules import CG_LSTM, ST_MGCN         # noqa: F401

__all__ = ["CG_LSTM", "ST_MGCN", "GCN"]
