"""Weight blob layout + Keras-like wrapper of the device network (seam S1).

The blob is the ``ck_net_set_weights`` format of include/ckb200.h: float32, Keras layouts
(conv ``[kh,kw,Cin,Cout]``, dense ``[in,out]``), in the layer order of ``create_nn``
(reference training_pipeline.py:44-120) grouped trunk / policy head / value head.
PyTorch is only used to hold the blob on the device (``TorchWeights``).
"""
import collections

import numpy as np

NET_PARAM_COUNT = 1321774
BN_EPS = 1e-3


def layout():
    """-> OrderedDict name -> (offset, shape) over the flat blob."""
    out = collections.OrderedDict()
    o = 0

    def add(name, shape):
        nonlocal o
        out[name] = (o, tuple(shape))
        o += int(np.prod(shape))

    def conv(name, k, cin, cout):
        add(name + "/kernel", (k, k, cin, cout))
        add(name + "/bias", (cout,))
        for s in ("gamma", "beta", "mean", "var"):
            add(name + "/bn_" + s, (cout,))

    conv("conv0", 3, 14, 128)
    for i in range(1, 7):
        conv("conv%d" % i, 3, 128, 128)
    conv("policy_conv1", 3, 128, 128)
    conv("policy_conv2", 1, 128, 8)
    add("policy_head/kernel", (512, 512))
    add("policy_head/bias", (512,))
    conv("value_conv1", 1, 128, 1)
    add("value_dense1/kernel", (64, 64))
    add("value_dense1/bias", (64,))
    for s in ("gamma", "beta", "mean", "var"):
        add("value_dense1/bn_" + s, (64,))
    add("value_head/kernel", (64, 1))
    add("value_head/bias", (1,))
    assert o == NET_PARAM_COUNT, o
    return out


def unpack(blob):
    blob = np.asarray(blob, dtype=np.float32).reshape(-1)
    return {k: blob[o:o + int(np.prod(s))].reshape(s) for k, (o, s) in layout().items()}


def random_init_blob(seed=0, bn_jitter=0.0):
    """Keras default initialisation: Glorot-uniform kernels, zero biases, BN
    gamma=1 beta=0 mean=0 var=1.  ``bn_jitter`` > 0 perturbs biases and BN statistics so that
    tests exercise every parameter (a trained net has non-trivial ones)."""
    rng = np.random.RandomState(seed)
    blob = np.zeros(NET_PARAM_COUNT, dtype=np.float32)
    for name, (o, shape) in layout().items():
        n = int(np.prod(shape))
        if name.endswith("/kernel"):
            if len(shape) == 4:
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
            else:
                fan_in, fan_out = shape
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            v = rng.uniform(-lim, lim, size=n)
        elif name.endswith("bn_gamma") or name.endswith("bn_var"):
            v = np.ones(n) + (rng.uniform(-bn_jitter, bn_jitter, size=n) if bn_jitter else 0.0)
            if name.endswith("bn_var"):
                v = np.abs(v)
        else:
            v = rng.uniform(-bn_jitter, bn_jitter, size=n) if bn_jitter else np.zeros(n)
        blob[o:o + n] = v.astype(np.float32)
    return blob


class TorchWeights(object):
    """Holds the blob as a torch CUDA tensor (the only thing PyTorch does on this path)."""

    def __init__(self, blob, device=0):
        import torch
        self.tensor = torch.from_numpy(np.ascontiguousarray(blob, dtype=np.float32)).to("cuda:%d" % device)

    def attach(self, net):
        net.set_weights_device(self.tensor.data_ptr(), self.tensor.numel())
        return net


class StubNet(object):
    """Deterministic stub evaluator that exists on the device (parity tests / demos): ``kind`` in
    'uniform_zero', 'uniform_material', 'hash', 'hash_salted'."""

    def __init__(self, kind):
        self.ck_evaluator = kind

    def predict(self, x):
        raise TypeError("StubNet evaluates on the device only (inside the search kernels)")


class KerasLikeNet(object):
    """Object with Keras' ``predict`` signature (Checkers.py:433) backed by libckb200."""
    ck_evaluator = "net"

    def __init__(self, blob, device=0, impl=None):
        from . import lib
        self.net = lib.Net(device, impl)
        self.net.set_weights(blob)

    def predict(self, x):
        return self.net.predict(np.asarray(x, dtype=np.float32))
