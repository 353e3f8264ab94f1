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


class PReLUAct(nn.PReLU):
    """nn.PReLU parameters (state_dict key `act.weight`); applied by the prelu kernels."""
    name = 'prelu'
    code = None

    def forward(self, x):
        return ops.prelu(x, self.weight)


def create_act(act, num_parameters=None):
    if act in ('relu', 'sigmoid', 'tanh', 'identity'):
        return Act(act)
    if act == 'prelu':
        return PReLUAct(num_parameters if num_parameters is not None else 1)
    raise ValueError('Unknown activation function {}'.format(act))


def apply_linear_act(x, lin, act_module, layout='oi', act_bwd_by_consumer=False, input_act=0):
    """act(linear(x)) with the activation fused into the GEMM epilogue when it can be.
    act_bwd_by_consumer: the (BatchNorm) consumer folds act' into its backward (ops.seg_batch_norm input_act)."""
    if act_module is None:
        return ops.linear_act(x, lin.weight, lin.bias, 0, layout, False, input_act)
    if act_module.code is not None:
        return ops.linear_act(x, lin.weight, lin.bias, act_module.code, layout, act_bwd_by_consumer, input_act)
    return act_module(ops.linear_act(x, lin.weight, lin.bias, 0, layout, False, input_act))


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
        self.layer_channels = [input_dim] + list(hidden_channels) + [output_dim]
        self.activation = create_act(activation_type)
        self.layers = nn.ModuleList()
        for i in range(len(self.layer_channels) - 1):
            lin = nn.Linear(self.layer_channels[i], self.layer_channels[i + 1])
            nn.init.xavier_uniform_(lin.weight, gain=nn.init.calculate_gain('relu'))
            self.layers.append(lin)
        self.bn = bn
        if self.bn:
            self.bn = nn.ModuleList([nn.BatchNorm1d(dim) for dim in self.layer_channels[1:-1]])

    def forward(self, x, final_act=0, seg=None):
        """seg = (row_ptr int32 device tensor, S): independent BatchNorm batches (needed when bn)."""
        n = len(self.layers)
        for i, lin in enumerate(self.layers):
            if i == n - 1:
                x = ops.linear_act(x, lin.weight, lin.bias, final_act, 'oi')
            elif self.bn:
                x = ops.linear_act(x, lin.weight, lin.bias, 0, 'oi')
                b = self.bn[i]
                if self.training:
                    ptr, S = seg
                    x = ops.seg_batch_norm(x, b.weight, b.bias, ptr, S, b.running_mean, b.running_var,
                                           b.num_batches_tracked, b.eps, b.momentum)
                else:
                    x = ops.bn_eval(x, b.weight, b.bias, b.running_mean, b.running_var, b.eps)
                x = ops.activation(x, self.activation.code) if self.activation.code is not None else self.activation(x)
            else:
                x = apply_linear_act(x, lin, self.activation)
        return x
