"""Same public API and parameter tree as the reference's lib/pointnet2/pytorch_utils.py
(SharedMLP :11-36, _BNBase :39-64, _ConvBase :67-120, Conv1d/2d/3d :123-233, FC :236-268,
set_bn_momentum_default / BNMomentumScheduler :270-296), written from scratch.

The state-dict contract is what matters: a SharedMLP owns `layer{i}` children, each
a sequential block with `conv` (bias only when there is no BN), `bn` (itself a block with one
child `bn`) and `activation`, so keys read `...layer0.conv.weight`, `...layer0.bn.bn.running_mean`.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

# Training mode: run BatchNorm2d + ReLU of a conv block through the fused sm_100a kernels (csrc/bn_relu.cu)
# instead of cuDNN batch-norm + a separate ReLU pass.  Same parameters, buffers and results (fp32, within
# summation-order rounding); set to False to get the plain module sequence.
FUSED_BN_RELU_TRAINING = True


class _BnReluTrain(torch.autograd.Function):
    """z = relu(batch_norm(y, training=True)); saves y and the batch statistics, recomputes the ReLU mask."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps):
        from . import _ext
        z, mean, invstd = _ext.bn_relu_train_forward(y, gamma, beta, running_mean, running_var, momentum, eps)
        ctx.save_for_backward(y, gamma, beta, mean, invstd)
        return z

    @staticmethod
    def backward(ctx, dz):
        from . import _ext
        y, gamma, beta, mean, invstd = ctx.saved_tensors
        dy, dgamma, dbeta = _ext.bn_relu_train_backward(dz.contiguous(), y, gamma, beta, mean, invstd)
        return dy, dgamma, dbeta, None, None, None, None


def _max_pool_last(x):
    """The SA module's pooling (pointnet2_modules.py:256-259) for the shapes the fused kernel does not take."""
    return torch.nn.functional.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)


class _BnReluMaxPoolTrain(torch.autograd.Function):
    """pooled = max_k relu(batch_norm(y, training=True))[..., k] for y (B,C,npoint,nsample)."""

    @staticmethod
    def forward(ctx, y, gamma, beta, running_mean, running_var, momentum, eps):
        from . import _ext
        pooled, argmax, ymax, mean, invstd = _ext.bn_relu_maxpool_train_forward(
            y, gamma, beta, running_mean, running_var, momentum, eps)
        ctx.save_for_backward(y, gamma, beta, mean, invstd, argmax, ymax)
        return pooled

    @staticmethod
    def backward(ctx, dpool):
        from . import _ext
        y, gamma, beta, mean, invstd, argmax, ymax = ctx.saved_tensors
        dy, dgamma, dbeta = _ext.bn_relu_maxpool_train_backward(dpool.contiguous(), argmax, ymax, y, gamma, beta,
                                                                mean, invstd)
        return dy, dgamma, dbeta, None, None, None, None


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
        # conv -> BatchNorm2d -> ReLU is the only arrangement SpaCap3D instantiates; remember it for forward()
        self._fusable = (not preact and norm_unit is not None and isinstance(activation, nn.ReLU)
                         and isinstance(conv_unit, nn.Conv2d))
        self._conv_name, self._bn_name = name + "conv", name + "bn"   # names, not modules: no second registration

    def _fused_training_ok(self, x):
        if not (self._fusable and FUSED_BN_RELU_TRAINING and self.training and x.is_cuda
                and x.dtype == torch.float32 and torch.is_grad_enabled()):
            return False
        bn = getattr(self, self._bn_name)[0]
        return bn.training and bn.affine and bn.track_running_stats and bn.momentum is not None

    def forward(self, x, max_pool_last_dim=False):
        """max_pool_last_dim=True additionally takes the max over the last axis (the SA module's
        F.max_pool2d(kernel=[1, nsample]) + squeeze) and returns (B, C, npoint)."""
        conv = getattr(self, self._conv_name)
        if x.shape[1] > conv.in_channels:
            # the caller appended all-zero channels for 16-byte-aligned GEMM rows (QueryAndGroup pad_channels_to):
            # the same convolution with matching zero weight columns; autograd slices the weight gradient back
            extra = x.shape[1] - conv.in_channels
            w = F.pad(conv.weight, (0, 0) * (conv.weight.dim() - 2) + (0, extra))
            run_conv = lambda t: F.conv2d(t, w, conv.bias, conv.stride, conv.padding)   # noqa: E731
        else:
            run_conv = conv
        if self._fused_training_ok(x):
            bn = getattr(self, self._bn_name)[0]
            y = run_conv(x)
            if y.is_contiguous() and y.shape[0] * y.shape[1] <= 65535:
                args = (y, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.momentum), float(bn.eps))
                if max_pool_last_dim and y.dim() == 4 and y.shape[3] in (16, 32, 64):
                    out = _BnReluMaxPoolTrain.apply(*args)
                else:
                    out = _BnReluTrain.apply(*args)
                    if max_pool_last_dim:
                        out = _max_pool_last(out)
                if bn.num_batches_tracked is not None:
                    bn.num_batches_tracked.add_(1)
                return out
            out = torch.relu_(getattr(self, self._bn_name)(y))
        elif run_conv is not conv:
            out = x
            for name, mod in self.named_children():
                out = run_conv(out) if name == self._conv_name else mod(out)
        else:
            out = super().forward(x)
        return _max_pool_last(out) if max_pool_last_dim else out


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

    def forward_max_pooled(self, x):
        """self(x) followed by the max over the last axis, with the last block's BatchNorm + ReLU + max
        fused in training mode (csrc/bn_relu.cu); returns (B, C_out, npoint)."""
        blocks = list(self.children())
        for blk in blocks[:-1]:
            x = blk(x)
        return blocks[-1](x, max_pool_last_dim=True)


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
