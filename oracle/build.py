"""Build step of the oracle's C restatement (called from __graft_entry__.build())."""
from .c_oracle import build  # noqa: F401
