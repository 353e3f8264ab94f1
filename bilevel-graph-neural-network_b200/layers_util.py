"""MLP / activation helpers with the reference's module layout
(model/layers_util.py:12-76) over the fused GEMM+bias+activation kernel."""
import torch
import torch.nn as nn

from . import ops


class Act(nn.Module):
    """Activation tag; the arithmetic is the epilogue of the producing kernel."""

    def __init__(self, name):
        super().__init__()
        self.name = name
        self.code = ops.act_code(name)

    def forward(self, x):       # standalone use (not on the fused path)
        raise RuntimeError('Act is applied as a kernel epilogue; call the owning layer')

    def extra_repr(self):
        return self.name


def create_act(act, num_parameters=None):
    if act in ('relu', 'sigmoid', 'tanh', 'identity'):
        return Act(act)
    if act == 'prelu':
        raise NotImplementedError('prelu is not on the B200 path yet')
    raise ValueError('Unknown activation function {}'.format(act))


class MLP(nn.Module):
    """model/layers_util.py:12-57: Linear stack, xavier_uniform(gain=relu) weights, the
    activation on every layer but the last.  `final_act` lets the caller fold a trailing
    sigmoid (LinkPred) into the last GEMM's epilogue."""

    def __init__(self, input_dim, output_dim, activation_type='relu', num_hidden_lyr=2,
                 hidden_channels=None, bn=False):
        super().__init__()
        self.out_dim = output_dim
        if not hidden_channels:
            hidden_channels = [input_dim for _ in range(num_hidden_lyr)]
        elif len(hidden_channels) != num_hidden_lyr:
            raise ValueError('number of hidden layers should be the same as the lengh of hidden_channels')
        if bn:
            raise NotImplementedError('MLP with BatchNorm is not on the Bi-GNN path')
        self.layer_channels = [input_dim] + list(hidden_channels) + [output_dim]
        self.activation = create_act(activation_type)
        self.layers = nn.ModuleList()
        for i in range(len(self.layer_channels) - 1):
            lin = nn.Linear(self.layer_channels[i], self.layer_channels[i + 1])
            nn.init.xavier_uniform_(lin.weight, gain=nn.init.calculate_gain('relu'))
            self.layers.append(lin)
        self.bn = bn

    def forward(self, x, final_act=0):
        n = len(self.layers)
        for i, lin in enumerate(self.layers):
            act = self.activation.code if i < n - 1 else final_act
            x = ops.linear_act(x, lin.weight, lin.bias, act, 'oi')
        return x
