"""Compatibility shim for running the reference's PLUMBING (train.py, helper.py, utils.py, preprocessing/) on this
container's library versions — TEST INFRASTRUCTURE.  The hot path (layers.py / models.py) needs none of it.

  (1) np.int / np.float were removed from numpy            — utils.py:161…, helper.py:121,150
  (2) networkx.to_scipy_sparse_matrix was removed in nx 3   — preprocessing/structure_generation.py:53
  (3) DataFrame.applymap was removed in pandas 3            — helper.py:216
  (6) np.random.normal(loc = 1×1 np.matrix, …) no longer broadcasts to a 1-D sample        — helper.py:131,143
  (5) train.get_gnn_model imports every baseline model unconditionally (train.py:93-100) and those import
      torch_geometric / torch_scatter, which are not installed → stub modules (none of the baselines is instantiated)
Reference files are never modified."""
import sys
import types

import numpy as np
import scipy.sparse as sp


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        import torch
        return type(name, (torch.nn.Module,), {})


def apply():
    import networkx as nx
    import pandas as pd
    np.int = int
    np.float = float
    if not hasattr(nx, "to_scipy_sparse_matrix"):
        nx.to_scipy_sparse_matrix = lambda g, nodelist=None, **kw: sp.csr_matrix(nx.to_scipy_sparse_array(g, nodelist=nodelist, **kw))
    if not hasattr(pd.DataFrame, "applymap"):
        pd.DataFrame.applymap = pd.DataFrame.map
    if not getattr(np.random.normal, "_ctgcn_compat", False):
        _normal = np.random.normal

        def normal(loc=0.0, scale=1.0, size=None):
            if isinstance(loc, np.matrix) and loc.size == 1:      # `for degree in degrees` over an [N,1] matrix (helper.py:130)
                loc = float(loc.item())
            return _normal(loc, scale, size)
        normal._ctgcn_compat = True
        np.random.normal = normal
    for mod in ("torch_geometric", "torch_geometric.nn", "torch_geometric.nn.conv", "torch_geometric.nn.inits",
                "torch_geometric.utils", "torch_geometric.nn.conv.gcn_conv", "torch_geometric.data", "torch_scatter",
                "torch_sparse", "torch_cluster"):
        sys.modules.setdefault(mod, _Stub(mod))
