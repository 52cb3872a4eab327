"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference network in PyTorch.

PARITY UNPINNED: the reference evaluates its network with TensorFlow 2.2 / Keras 2.4.3
(requirements.txt:42-43,103-106), which is neither vendored under /root/reference nor
installable here, and the reference ships no golden activations.  This module restates
``create_nn`` (training_pipeline.py:44-120) from the layer list and Keras' documented layer
semantics (Conv2D 'same' cross-correlation, bias, ReLU, then BatchNormalization with moving
statistics and eps=1e-3; Flatten over (x, y, c); Dense kernels [in, out]); the CUDA kernels are
compared against it, not against TensorFlow.
"""
import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3


def forward(params, x, dtype=torch.float64, pre_activation=False, features=False):
    """params: dict from ckb200.net.unpack; x: [n,8,8,14] channels-last -> (policy [n,512], value [n]);
    with ``pre_activation`` also the policy logits [n,512] (before the softmax) and the value head's
    pre-tanh output [n] -- the quantities north_star's 1e-5 contract is stated on."""
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)

    def bn(h, name, dim):
        shape = [1] * h.dim()
        shape[dim] = -1
        g, b, m, v = (t(params[name + "/bn_" + s]).reshape(shape) for s in ("gamma", "beta", "mean", "var"))
        return (h - m) / torch.sqrt(v + BN_EPS) * g + b

    def conv(h, name):
        k = t(params[name + "/kernel"]).permute(3, 2, 0, 1)       # [kh,kw,Cin,Cout] -> [Cout,Cin,kh,kw]
        h = F.conv2d(h, k, t(params[name + "/bias"]), padding=k.shape[-1] // 2)
        return bn(F.relu(h), name, 1)

    h = t(x).permute(0, 3, 1, 2)                                   # NHWC -> NCHW
    for i in range(7):
        h = conv(h, "conv%d" % i)
    p = conv(conv(h, "policy_conv1"), "policy_conv2")
    p = p.permute(0, 2, 3, 1).reshape(len(x), 512)                 # Flatten over (x, y, c)
    pflat = p
    logits = p @ t(params["policy_head/kernel"]) + t(params["policy_head/bias"])
    p = torch.softmax(logits, dim=1)
    v = conv(h, "value_conv1").permute(0, 2, 3, 1).reshape(len(x), 64)
    vconv = v
    v = F.relu(v @ t(params["value_dense1/kernel"]) + t(params["value_dense1/bias"]))
    v = bn(v, "value_dense1", 1)
    vpre = v @ t(params["value_head/kernel"]) + t(params["value_head/bias"])
    v = torch.tanh(vpre)
    if features:        # also what the tower kernels hand to the heads: flattened policy features and the value conv output
        return p.numpy(), v.reshape(-1).numpy(), logits.numpy(), vpre.reshape(-1).numpy(), pflat.numpy(), vconv.numpy()
    if pre_activation:
        return p.numpy(), v.reshape(-1).numpy(), logits.numpy(), vpre.reshape(-1).numpy()
    return p.numpy(), v.reshape(-1).numpy()


class TorchKerasLike(object):
    """Keras-like ``predict`` on the CPU (the stand-in for ``load_model`` in the reference arm)."""

    def __init__(self, blob, dtype=torch.float32):
        from ckb200 import net as N
        self.params = N.unpack(blob)
        self.dtype = dtype
        self.calls = 0

    def predict(self, x):
        self.calls += 1
        with torch.no_grad():
            p, v = forward(self.params, np.asarray(x, dtype=np.float32), self.dtype)
        return [p.astype(np.float32), v.astype(np.float32).reshape(-1, 1)]
