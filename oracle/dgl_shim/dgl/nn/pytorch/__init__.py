"""TEST INFRASTRUCTURE ONLY -- ``dgl.nn.pytorch.HeteroGraphConv`` of the dgl shim.

Rule (iv) of dgl/__init__.py, restating dgl 0.5.2 ``HeteroGraphConv.forward``: iterate the
canonical etypes; skip relations with zero edges or whose src/dst inputs are missing; for blocks
the dst inputs are the first ``number_of_dst_nodes`` rows of the src inputs; per destination type
stack the per-relation results and reduce with sum / mean / max / min / stack; destination
types that received nothing are omitted. Sub-modules live in ``self.mods`` (an ``nn.ModuleDict``
keyed by etype name) so state_dict keys read ``layers.{i}.mods.{etype}.fc_self.weight``.
"""
import torch
import torch.nn as nn


class HeteroGraphConv(nn.Module):
    def __init__(self, mods, aggregate='sum'):
        super().__init__()
        self.mods = nn.ModuleDict(mods)
        if aggregate not in ('sum', 'mean', 'max', 'min', 'stack'):
            raise KeyError(aggregate)
        self.aggregate = aggregate

    def forward(self, g, inputs, mod_args=None, mod_kwargs=None):
        outputs = {nty: [] for nty in g.dsttypes}
        if isinstance(inputs, tuple) or g.is_block:
            if isinstance(inputs, tuple):
                src_inputs, dst_inputs = inputs
            else:
                src_inputs = inputs
                dst_inputs = {k: v[:g.number_of_dst_nodes(k)] for k, v in inputs.items()}
            for stype, etype, dtype in g.canonical_etypes:
                rel_graph = g[stype, etype, dtype]
                if rel_graph.number_of_edges() == 0:
                    continue
                if stype not in src_inputs or dtype not in dst_inputs:
                    continue
                outputs[dtype].append(self.mods[etype](rel_graph, (src_inputs[stype], dst_inputs[dtype])))
        else:
            for stype, etype, dtype in g.canonical_etypes:
                rel_graph = g[stype, etype, dtype]
                if rel_graph.number_of_edges() == 0:
                    continue
                if stype not in inputs:
                    continue
                outputs[dtype].append(self.mods[etype](rel_graph, inputs[stype]))
        rsts = {}
        for nty, alist in outputs.items():
            if len(alist) != 0:
                stacked = torch.stack(alist, dim=0)
                if self.aggregate == 'sum':
                    rsts[nty] = stacked.sum(0)
                elif self.aggregate == 'mean':
                    rsts[nty] = stacked.mean(0)
                elif self.aggregate == 'max':
                    rsts[nty] = stacked.max(0)[0]
                elif self.aggregate == 'min':
                    rsts[nty] = stacked.min(0)[0]
                else:
                    rsts[nty] = torch.stack(alist, dim=1)
        return rsts
