"""Samplers of positive train pairs (src/sampler.py).

RandomSampler (src/sampler.py:110-131, the default) and EverythingSampler (:51-69) draw through the same torch
DataLoader mechanism as the reference (same global-RNG consumption), so batches are bit-identical to the
reference's for an identical RNG state.  NeighborSampler (:72-107) restates the reference's BFS sub-graph
sampling over the HBM-resident interaction graph's host CSR instead of a networkx graph: same use of Python's
global `random`, same set / dict iteration semantics (CPython), so the sampled drug sets, their order and the order
of the induced pairs equal the reference's run under the networkx installed next to it (tests/golden/
bignn_samplers.npz was recorded with networkx 3.x; the reference pins 2.2, whose sub-graph views iterate
differently for hub nodes -- pair ORDER may differ there, the pair SET cannot).
"""
import random
from collections import deque

import numpy as np
from torch.utils.data import DataLoader


def _host_adjacency(graph):
    """neighbour lists of the interaction graph in the reference graph's order: ascending (nx.from_scipy_sparse_matrix
    adds the edges of a sorted symmetric matrix row by row)."""
    ptr = np.asarray(graph.csr._row_ptr_host, np.int64)
    col = np.asarray(graph.col_host, np.int64)
    return [col[ptr[i]:ptr[i + 1]].tolist() for i in range(graph.number_of_nodes())]


def _induced_edges(adj, n_nodes, nodes, copied):
    """Edges of the sub-graph induced by `nodes`, in the order networkx 3.x reports them: the node view walks the
    filter set when it is less than half the size of the graph, else the graph's node order; each node's neighbours
    come in adjacency order; an undirected edge is reported from its first endpoint.  `copied`: the order of
    `G.subgraph(nodes).copy().edges` (the copy keeps first-insertion order of each adjacency), else of the view's
    own `.edges`."""
    keep = set(n for n in nodes)                     # nbunch_iter -> a fresh set with its own iteration order
    order = [n for n in keep] if 2 * len(keep) < n_nodes else [n for n in range(n_nodes) if n in keep]
    if copied:
        sub = {n: {} for n in order}
        for u in order:
            for v in adj[u]:
                if v in keep:
                    sub[u][v] = None
                    sub[v][u] = None
    else:
        sub = {u: [v for v in adj[u] if v in keep] for u in order}
    edges, seen = [], set()
    for n, nbrs in sub.items():
        for v in nbrs:
            if v not in seen:
                edges.append((n, v))
        seen.add(n)
    return order, edges


class RandomSampler(object):
    """Shuffled mini-batches of positive train pairs (src/sampler.py:110-131); with `sample_induced` every train pair
    between the drugs of the drawn batch (:133-142)."""

    def __init__(self, data, batch_size, sample_induced=False):
        self.batch_size = batch_size
        self.sample_induced = sample_induced
        self.data_loader = DataLoader(data, batch_size=batch_size, shuffle=True)
        self.data_iterable = iter(self.data_loader)
        if sample_induced:
            ds = data.dataset
            self.id_map, self.gs_map = ds.id_map, ds.gs_map
            self.num_nodes = ds.interaction_combo_nxgraph.number_of_nodes()
            self.nodes_visited_counter = np.zeros(self.num_nodes)
            self._adj = _host_adjacency(ds.interaction_combo_nxgraph)

    def _next_pairs(self):
        pairs = next(self.data_iterable, None)
        if pairs is None:                                   # epoch over: a fresh shuffle
            self.data_iterable = iter(self.data_loader)
            pairs = next(self.data_iterable)
        return pairs.cpu().detach().numpy()

    def sample_next_training_batch(self):
        batch_gids = self._next_pairs()
        if not self.sample_induced:
            return batch_gids, np.unique(batch_gids), None
        rows = [self.gs_map[int(g)] for g in np.unique(batch_gids)]
        _, edges = _induced_edges(self._adj, self.num_nodes, rows, copied=False)
        induced = np.asarray([(self.id_map[u], self.id_map[v]) for u, v in edges])
        for r in np.unique(np.asarray(edges)):
            self.nodes_visited_counter[r] += 1
        return induced, np.unique(induced), None


class EverythingSampler(RandomSampler):
    """All train pairs as one shuffled batch (src/sampler.py:51-69)."""

    def __init__(self, data):
        super().__init__(data, len(data.dataset.train_pairs))


class SampledSubgraph(object):
    """What the reference hands on as an nx sub-graph: the sampled drug rows and the induced train pairs."""

    def __init__(self, nodes, edges):
        self.nodes, self.edges = list(nodes), list(edges)

    def number_of_nodes(self):
        return len(self.nodes)


class NeighborSampler(object):
    def __init__(self, data, neighbor_size, batch_size):
        ds = data.dataset
        self.id_map, self.gs_map = ds.id_map, ds.gs_map
        g = ds.interaction_combo_nxgraph
        self.batch_size = batch_size
        self.neighbor_size = neighbor_size              # int, or a fraction of the neighbours
        self.num_nodes = g.number_of_nodes()
        self.nodes_visited_counter = np.zeros(self.num_nodes)
        self._adj = _host_adjacency(g)

    # -- src/sampler.py:80-86
    def _sampled_neighbours(self, node):
        nbrs = list(self._adj[node])
        random.shuffle(nbrs)
        k = self.neighbor_size if type(self.neighbor_size) == int else int(len(nbrs) * self.neighbor_size)
        return iter(set(nbrs[:k]))

    # -- src/sampler.py:10-34: breadth-first growth until `limit` drugs are visited
    def _bfs(self, source, limit):
        visited = {source}
        frontier = deque([(source, self.num_nodes, self._sampled_neighbours(source))])
        while frontier:
            parent, depth, children = frontier[0]
            child = next(children, None)
            if child is None:
                frontier.popleft()
                continue
            if child in visited:
                continue
            visited.add(child)
            if len(visited) == limit:
                break
            if depth > 1:
                frontier.append((child, depth - 1, self._sampled_neighbours(child)))
        return visited

    # -- src/sampler.py:94-107
    def _sample_nodes(self):
        nodes = set()
        while len(nodes) < self.batch_size:
            room = self.batch_size - len(nodes)
            cand = random.randint(0, self.num_nodes - 1)
            while cand in nodes or (np.any(self.nodes_visited_counter == 0) and self.nodes_visited_counter[cand] > 0):
                cand = random.randint(0, self.num_nodes - 1)
            self.nodes_visited_counter[cand] += 1
            nodes = nodes.union(self._bfs(cand, room))
        return nodes

    def sample_next_training_batch(self):
        nodes = self._sample_nodes()
        order, edges = _induced_edges(self._adj, self.num_nodes, nodes, copied=True)
        batch_gids = np.asarray([(self.id_map[u], self.id_map[v]) for u, v in edges])
        return batch_gids, [self.id_map[n] for n in nodes], SampledSubgraph(order, edges)
