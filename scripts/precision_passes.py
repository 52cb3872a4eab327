"""How many fp16 tensor-core passes does the 1e-5 contract need?  CPU emulation (PyTorch float64) of the tower's
split-operand arithmetic: every convolution / dense product is formed from fp16 pieces of the weights
(W = Whi + Wlo) and of the activations (A = Ahi + Alo) with exact accumulation, keeping

    3 passes   Whi*Ahi + Whi*Alo + Wlo*Ahi          (the product kernels; drops only Wlo*Alo ~ 2^-22)
    2 passes   Whi*Ahi + Wlo*Ahi                    (activations rounded to fp16)
    2 passes   Whi*Ahi + Whi*Alo                    (weights rounded to fp16)
    1 pass     Whi*Ahi
    tf32       both operands rounded to 10-bit mantissas (one kind::tf32 pass = the cost of two fp16 passes)

and compares policy (softmax outputs) and value (tanh) with the float64 network on positions from random play.
Usage: python scripts/precision_passes.py [weights.npy | weights.h5] [n_positions]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "checkers-mcts_b200")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from ckb200 import codec, net as N  # noqa: E402
from oracle import oracle as O  # noqa: E402

BN_EPS = 1e-3


def split16(t):
    hi = t.to(torch.float16).to(torch.float64)
    lo = (t - hi).to(torch.float16).to(torch.float64)
    return hi, lo


def tf32(t):
    x = t.to(torch.float32).contiguous()
    bits = x.view(torch.int32)
    bits = (bits + 0x1000) & ~0x1FFF                   # round to nearest, 10 explicit mantissa bits
    return bits.view(torch.float32).to(torch.float64)


def product(op, w, a, mode):
    """op(weights, activations) linear in both; w, a float64 tensors"""
    a = a.to(torch.float32).to(torch.float64)          # activations live in fp32 between layers
    if mode == "f64":
        return op(w, a)
    if mode == "tf32":
        return op(tf32(w), tf32(a))
    whi, wlo = split16(w)
    ahi, alo = split16(a)
    if mode == "3":
        return op(whi, ahi + alo) + op(wlo, ahi)
    if mode == "2a":
        return op(whi + wlo, ahi)
    if mode == "2w":
        return op(whi, ahi + alo)
    if mode == "1":
        return op(whi, ahi)
    raise ValueError(mode)


def forward(params, x, mode):
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(torch.float64)

    def bn(h, name, dim):
        shape = [1] * h.dim()
        shape[dim] = -1
        g, b, m, v = (t(params[name + "/bn_" + s]).reshape(shape) for s in ("gamma", "beta", "mean", "var"))
        return (h - m) / torch.sqrt(v + BN_EPS) * g + b

    def conv(h, name):
        k = t(params[name + "/kernel"]).permute(3, 2, 0, 1)
        y = product(lambda w, a: F.conv2d(a, w, None, padding=k.shape[-1] // 2), k, h, mode)
        return bn(F.relu(y + t(params[name + "/bias"]).reshape(1, -1, 1, 1)), name, 1)

    def dense(h, name):
        return product(lambda w, a: a @ w, t(params[name + "/kernel"]), h, mode) + t(params[name + "/bias"])

    h = t(x).permute(0, 3, 1, 2)
    for i in range(7):
        h = conv(h, "conv%d" % i)
    p = conv(conv(h, "policy_conv1"), "policy_conv2").permute(0, 2, 3, 1).reshape(len(x), 512)
    p = torch.softmax(dense(p, "policy_head"), dim=1)
    v = conv(h, "value_conv1").permute(0, 2, 3, 1).reshape(len(x), 64)
    v = bn(F.relu(dense(v, "value_dense1")), "value_dense1", 1)
    v = torch.tanh(dense(v, "value_head"))
    return p.numpy(), v.reshape(-1).numpy()


def positions(n, seed=0):
    rng = np.random.RandomState(seed)
    out = []
    while len(out) < n:
        pos = O.start_position()
        for _ in range(rng.randint(0, 80)):
            kids, mask, status, p5 = O.movegen(pos)
            if status != 0:
                break
            pos = kids[rng.randint(len(kids))]
        kids, mask, status, p5 = O.movegen(pos)
        if status == 0:
            out.append(codec.nn_input_planes(pos, mask, p5))
    return np.stack(out).astype(np.float32)


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else None
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    if src is None:
        blob, name = N.random_init_blob(0, bn_jitter=0.2), "random init (seed 0, BN statistics jittered)"
    else:
        from training_pipeline import load_blob
        blob, name = load_blob(src), os.path.basename(src)
    params = N.unpack(blob)
    x = positions(n)
    with torch.no_grad():
        ref_p, ref_v = forward(params, x, "f64")
        print("weights: %s; %d positions; max abs error against float64 (contract: 1e-5)" % (name, n))
        for mode, label in (("3", "3 fp16 passes (product)"), ("2a", "2 passes, activations fp16"), ("2w", "2 passes, weights fp16"),
                            ("1", "1 pass"), ("tf32", "tf32 operands")):
            p, v = forward(params, x, mode)
            print("  %-28s policy %.2e   value %.2e" % (label, np.abs(p - ref_p).max(), np.abs(v - ref_v).max()))


if __name__ == "__main__":
    main()
