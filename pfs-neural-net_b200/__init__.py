"""pfs-neural-net_b200: the B200 (sm_100a) message-passing layer of pfs-neural-net.

The directory name is not a Python identifier; import it as `pfs_neural_net_b200` (the sibling
alias package points its __path__ here).  Contents: `gnn` (drop-in for the reference's src/gnn.py),
`functional` (autograd Functions over the C ABI), `topology` (edge_index -> dense / CSR / CSC),
`dp` (data-parallel helpers), `_abi` (ctypes binding of csrc/libpfs_b200.so).
"""
__version__ = "0.1.0"
