"""TEST INFRASTRUCTURE ONLY -- ``dgl.dataloading`` of the dgl shim (see dgl/__init__.py).

``MultiLayerFullNeighborSampler(n)`` + ``NodeDataLoader`` restate dgl 0.5.2: per batch of seed
nodes, for each layer from the output inwards: ``in_subgraph(g, seeds)`` -> ``to_block`` ->
seeds = the block's source nodes. ``negative_sampler.Uniform(k)`` draws, per positive edge (u, v) of
a canonical etype, k edges (u, randint(0, num_nodes(dsttype))) laid out k-consecutive (rule viii).
"""
import torch

from .. import in_subgraph, to_block, NID


class MultiLayerFullNeighborSampler:
    def __init__(self, n_layers, return_eids=False):
        self.num_layers = n_layers

    def sample_blocks(self, g, seed_nodes):
        blocks = []
        for _ in range(self.num_layers):
            frontier = in_subgraph(g, seed_nodes)
            block = to_block(frontier, seed_nodes)
            seed_nodes = {t: block.srcnodes[t].data[NID] for t in block.srctypes}
            blocks.insert(0, block)
        return blocks


class NodeDataLoader:
    def __init__(self, g, nids, block_sampler, batch_size=1, shuffle=False, drop_last=False, num_workers=0,
                 **kwargs):
        self.g, self.sampler = g, block_sampler
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.items = [(t, int(i)) for t in sorted(nids) for i in torch.as_tensor(nids[t]).reshape(-1).tolist()]

    def __len__(self):
        n = len(self.items)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = torch.randperm(len(self.items)).tolist() if self.shuffle else list(range(len(self.items)))
        for b in range(len(self)):
            batch = [self.items[j] for j in order[b * self.batch_size:(b + 1) * self.batch_size]]
            seeds = {}
            for t, i in batch:
                seeds.setdefault(t, []).append(i)
            seeds = {t: torch.tensor(v, dtype=torch.int64) for t, v in seeds.items()}
            blocks = self.sampler.sample_blocks(self.g, seeds)
            input_nodes = {t: blocks[0].srcnodes[t].data[NID] for t in blocks[0].srctypes}
            output_nodes = {t: blocks[-1].dstnodes[t].data[NID] for t in blocks[-1].dsttypes}
            yield input_nodes, output_nodes, blocks


class _Uniform:
    def __init__(self, k):
        self.k = k

    def __call__(self, g, eids_dict, generator=None):
        out = {}
        for c, eids in eids_dict.items():
            c = g.to_canonical_etype(c)
            s, _ = g.find_edges(eids, etype=c)
            src = s.repeat_interleave(self.k)
            dst = torch.randint(0, g.num_nodes(c[2]), (src.numel(),), generator=generator)
            out[c] = (src, dst)
        return out


class negative_sampler:  # noqa: N801 - mirrors the dgl module name
    Uniform = _Uniform
