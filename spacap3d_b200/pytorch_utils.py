"""Same public API and parameter tree as the reference's lib/pointnet2/pytorch_utils.py
(SharedMLP :11-36, _BNBase :39-64, _ConvBase :67-120, Conv1d/2d/3d :123-233, FC :236-268,
set_bn_momentum_default / BNMomentumScheduler :270-296), written from scratch.

The state-dict contract is what matters: a SharedMLP owns `layer{i}` children, each
a sequential block with `conv` (bias only when there is no BN), `bn` (itself a block with one
child `bn`) and `activation`, so keys read `...layer0.conv.weight`, `...layer0.bn.bn.running_mean`.
"""
import torch.nn as nn

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_NORM = {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}


class _BNBase(nn.Sequential):
    def __init__(self, in_size, batch_norm=None, name=""):
        super().__init__()
        norm = batch_norm(in_size)
        nn.init.constant_(norm.weight, 1.0)
        nn.init.constant_(norm.bias, 0)
        self.add_module(name + "bn", norm)


class BatchNorm1d(_BNBase):
    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, batch_norm=_NORM[1], name=name)


class BatchNorm2d(_BNBase):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, batch_norm=_NORM[2], name=name)


class BatchNorm3d(_BNBase):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, batch_norm=_NORM[3], name=name)


class _ConvBase(nn.Sequential):
    """[bn, act,] conv [, bn, act] depending on `preact`; conv bias only without BN."""

    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                 conv=None, batch_norm=None, bias=True, preact=False, name=""):
        super().__init__()
        use_bias = bias and not bn
        conv_unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride,
                         padding=padding, bias=use_bias)
        init(conv_unit.weight)
        if use_bias:
            nn.init.constant_(conv_unit.bias, 0)
        norm_unit = batch_norm(in_size if preact else out_size) if bn else None

        def _norm_act():
            if norm_unit is not None:
                self.add_module(name + "bn", norm_unit)
            if activation is not None:
                self.add_module(name + "activation", activation)

        if preact:
            _norm_act()
        self.add_module(name + "conv", conv_unit)
        if not preact:
            _norm_act()


def _make_conv(dim, wrapper):
    ones, zeros = (1,) * dim, (0,) * dim
    if dim == 1:
        ones, zeros = 1, 0

    class _Conv(_ConvBase):
        def __init__(self, in_size, out_size, *, kernel_size=ones, stride=ones, padding=zeros,
                     activation=nn.ReLU(inplace=True), bn=False, init=nn.init.kaiming_normal_,
                     bias=True, preact=False, name=""):
            super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init,
                             conv=_CONV[dim], batch_norm=wrapper, bias=bias, preact=preact,
                             name=name)

    _Conv.__name__ = _Conv.__qualname__ = "Conv%dd" % dim
    return _Conv


Conv1d = _make_conv(1, BatchNorm1d)
Conv2d = _make_conv(2, BatchNorm2d)
Conv3d = _make_conv(3, BatchNorm3d)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d (+BN +ReLU) blocks named layer0, layer1, ...  (reference :11-36)."""

    def __init__(self, args, *, bn=False, activation=nn.ReLU(inplace=True), preact=False,
                 first=False, name=""):
        super().__init__()
        for i, (cin, cout) in enumerate(zip(args[:-1], args[1:])):
            plain = first and preact and i == 0   # the very first pre-activation block is bare
            self.add_module(name + "layer{}".format(i),
                            Conv2d(cin, cout, bn=bn and not plain,
                                   activation=None if plain else activation, preact=preact))


class FC(nn.Sequential):
    def __init__(self, in_size, out_size, *, activation=nn.ReLU(inplace=True), bn=False, init=None,
                 preact=False, name=""):
        super().__init__()
        fc = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(fc.weight)
        if not bn:
            nn.init.constant_(fc.bias, 0)

        def _norm_act(width):
            if bn:
                self.add_module(name + "bn", BatchNorm1d(width))
            if activation is not None:
                self.add_module(name + "activation", activation)

        if preact:
            _norm_act(in_size)
        self.add_module(name + "fc", fc)
        if not preact:
            _norm_act(out_size)


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum
    return fn


class BNMomentumScheduler(object):
    """Sets every BN layer's momentum to bn_lambda(epoch) (used by lib/solver.py:19)."""

    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model = model
        self.setter = setter
        self.lmbd = bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
