"""`torch_geometric.utils` (1.1.2) subset."""
import torch
from torch_scatter import scatter_add, scatter_max
from torch_sparse import coalesce


def maybe_num_nodes(index, num_nodes=None):
    return int(index.max().item()) + 1 if num_nodes is None else num_nodes


def to_undirected(edge_index, num_nodes=None):
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    row, col = edge_index
    row, col = torch.cat([row, col], dim=0), torch.cat([col, row], dim=0)
    edge_index = torch.stack([row, col], dim=0)
    edge_index, _ = coalesce(edge_index, None, num_nodes, num_nodes)
    return edge_index


def is_undirected(edge_index, num_nodes=None):
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    edge_index, _ = coalesce(edge_index, None, num_nodes, num_nodes)
    undirected = to_undirected(edge_index, num_nodes=num_nodes)
    return edge_index.size(1) == undirected.size(1)


def contains_self_loops(edge_index):
    row, col = edge_index
    return bool((row == col).sum().item() > 0)


def remove_self_loops(edge_index, edge_attr=None):
    row, col = edge_index
    mask = row != col
    edge_attr = edge_attr if edge_attr is None else edge_attr[mask]
    mask = mask.unsqueeze(0).expand_as(edge_index)
    edge_index = edge_index[mask].view(2, -1)
    return edge_index, edge_attr


def add_self_loops(edge_index, num_nodes=None):
    num_nodes = maybe_num_nodes(edge_index, num_nodes)
    loop = torch.arange(0, num_nodes, dtype=torch.long, device=edge_index.device)
    loop = loop.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index, loop], dim=1)


def dense_to_sparse(tensor):
    index = tensor.nonzero().t().contiguous()
    value = tensor[index[0], index[1]]
    return index, value


def softmax(src, index, num_nodes=None):
    num_nodes = maybe_num_nodes(index, num_nodes)
    out = src - scatter_max(src, index, dim=0, dim_size=num_nodes)[0][index]
    out = out.exp()
    out = out / (scatter_add(out, index, dim=0, dim_size=num_nodes)[index] + 1e-16)
    return out
