"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for `torch_geometric` (reference src/gnn.py:2,7,49).
The reference only uses `torch_geometric.data.Data` as an attribute bag and `Dataset` as a base."""
from . import data  # noqa: F401
