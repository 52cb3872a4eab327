"""Training step of the AlphaZero loop in PyTorch: ``create_nn`` / ``train_nn`` of the
reference (training_pipeline.py:44-179) on the weight-blob format the CUDA evaluator consumes.

Off the hot path and plain PyTorch by design (the task statement keeps hand-written kernels for the
self-play path).  Semantics follow the reference's Keras code:
  * network: conv -> bias -> ReLU -> BatchNorm(eps 1e-3, momentum 0.99) x 7, policy / value heads (:57-112);
  * loss: POLICY_LOSS_WEIGHT * categorical cross-entropy (softmax output vs visit probabilities) +
    VALUE_LOSS_WEIGHT * MSE(value, (q + z) / 2) + L2 on every conv / dense kernel AND bias (:54-55,
    ``kernel_regularizer`` / ``bias_regularizer``);
  * optimiser: Adam (Keras defaults, eps 1e-7) with the triangular cyclical learning rate of
    CLR/clr_callback.py:105-111, stepped per batch, step size CLR_SS_COEFF * len(train) / BATCH_SIZE;
  * shuffle, hold out the last VAL_SPLIT of the data, early stopping on val_loss (PATIENCE, MIN_DELTA),
    keep the best epoch's weights and write them to data/model/Checkers_Model{it+1}_{timestamp}.npy.
Weights are exchanged as the flat float32 blob of ``ckb200.net.layout`` (Keras layouts).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import net as N


class CheckersNet(nn.Module):
    """create_nn (training_pipeline.py:44-120); parameters live in PyTorch layouts, blob() / load_blob()
    convert to / from the Keras-ordered blob."""

    def __init__(self, blob=None, seed=0):
        super().__init__()

        def conv(cin, cout, k):
            return nn.Conv2d(cin, cout, k, padding=k // 2, bias=True), nn.BatchNorm2d(cout, eps=N.BN_EPS, momentum=0.01)

        self.names = ["conv%d" % i for i in range(7)] + ["policy_conv1", "policy_conv2", "value_conv1"]
        shapes = [(14, 128, 3)] + [(128, 128, 3)] * 7 + [(128, 8, 1), (128, 1, 1)]
        self.convs, self.bns = nn.ModuleDict(), nn.ModuleDict()
        for name, (cin, cout, k) in zip(self.names, shapes):
            self.convs[name], self.bns[name] = conv(cin, cout, k)
        self.policy_head = nn.Linear(512, 512)
        self.value_dense1 = nn.Linear(64, 64)
        self.value_bn = nn.BatchNorm1d(64, eps=N.BN_EPS, momentum=0.01)
        self.value_head = nn.Linear(64, 1)
        self.load_blob(N.random_init_blob(seed) if blob is None else blob)

    # ---- blob <-> parameters ---------------------------------------------------------------------
    def load_blob(self, blob):
        p = N.unpack(np.asarray(blob, dtype=np.float32))
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        with torch.no_grad():
            for name in self.names:
                self.convs[name].weight.copy_(t(p[name + "/kernel"]).permute(3, 2, 0, 1))     # [kh,kw,Cin,Cout] -> [Cout,Cin,kh,kw]
                self.convs[name].bias.copy_(t(p[name + "/bias"]))
                bn = self.bns[name]
                bn.weight.copy_(t(p[name + "/bn_gamma"])); bn.bias.copy_(t(p[name + "/bn_beta"]))
                bn.running_mean.copy_(t(p[name + "/bn_mean"])); bn.running_var.copy_(t(p[name + "/bn_var"]))
            for lin, name in ((self.policy_head, "policy_head"), (self.value_dense1, "value_dense1"), (self.value_head, "value_head")):
                lin.weight.copy_(t(p[name + "/kernel"]).t())                                   # [in,out] -> [out,in]
                lin.bias.copy_(t(p[name + "/bias"]))
            bn = self.value_bn
            bn.weight.copy_(t(p["value_dense1/bn_gamma"])); bn.bias.copy_(t(p["value_dense1/bn_beta"]))
            bn.running_mean.copy_(t(p["value_dense1/bn_mean"])); bn.running_var.copy_(t(p["value_dense1/bn_var"]))
        return self

    def blob(self):
        out = np.zeros(N.NET_PARAM_COUNT, dtype=np.float32)
        lay = N.layout()

        def put(key, tensor):
            off, shape = lay[key]
            a = tensor.detach().cpu().numpy().astype(np.float32)
            assert a.shape == tuple(shape), (key, a.shape, shape)
            out[off:off + a.size] = a.reshape(-1)

        for name in self.names:
            put(name + "/kernel", self.convs[name].weight.permute(2, 3, 1, 0).contiguous())
            put(name + "/bias", self.convs[name].bias)
            bn = self.bns[name]
            put(name + "/bn_gamma", bn.weight); put(name + "/bn_beta", bn.bias)
            put(name + "/bn_mean", bn.running_mean); put(name + "/bn_var", bn.running_var)
        for lin, name in ((self.policy_head, "policy_head"), (self.value_dense1, "value_dense1"), (self.value_head, "value_head")):
            put(name + "/kernel", lin.weight.t().contiguous())
            put(name + "/bias", lin.bias)
        bn = self.value_bn
        put("value_dense1/bn_gamma", bn.weight); put("value_dense1/bn_beta", bn.bias)
        put("value_dense1/bn_mean", bn.running_mean); put("value_dense1/bn_var", bn.running_var)
        return out

    def save(self, filename):
        """stand-in for Keras ``model.save``: writes the blob as .npy (an '.h5' suffix is replaced)"""
        if filename.endswith(".h5"):
            filename = filename[:-3] + ".npy"
        np.save(filename, self.blob())
        return filename if filename.endswith(".npy") else filename + ".npy"

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, x):
        """x: [B,8,8,14] channels-last (the Keras_Generator layout) -> (policy probabilities [B,512], value [B])"""
        h = x.permute(0, 3, 1, 2)

        def block(h, name):
            return self.bns[name](F.relu(self.convs[name](h)))

        for i in range(7):
            h = block(h, "conv%d" % i)
        p = block(block(h, "policy_conv1"), "policy_conv2")
        p = p.permute(0, 2, 3, 1).reshape(x.shape[0], 512)                  # Flatten over (x, y, c)
        p = torch.softmax(self.policy_head(p), dim=1)
        v = block(h, "value_conv1").permute(0, 2, 3, 1).reshape(x.shape[0], 64)
        v = self.value_bn(F.relu(self.value_dense1(v)))
        v = torch.tanh(self.value_head(v)).reshape(-1)
        return p, v

    def predict(self, x):
        """Keras-like inference: [policy float32 [B,512], value float32 [B,1]]"""
        was = self.training
        self.eval()
        with torch.no_grad():
            dev = next(self.parameters()).device
            p, v = self(torch.as_tensor(np.asarray(x, dtype=np.float32), device=dev))
        self.train(was)
        return [p.cpu().numpy(), v.cpu().numpy().reshape(-1, 1)]

    def regularised(self):
        """tensors under the reference's l2 regulariser: every conv / dense kernel and bias (:54-55)"""
        for name in self.names:
            yield self.convs[name].weight
            yield self.convs[name].bias
        for lin in (self.policy_head, self.value_dense1, self.value_head):
            yield lin.weight
            yield lin.bias


def clr_triangular(iteration, base_lr, max_lr, step_size):
    """CLR/clr_callback.py:105-111, mode 'triangular' (scale_fn = 1)"""
    cycle = np.floor(1 + iteration / (2.0 * step_size))
    x = np.abs(iteration / float(step_size) - 2 * cycle + 1)
    return base_lr + (max_lr - base_lr) * max(0.0, 1 - x)


def loss_terms(model, x, probs, target_v, policy_w=1.0, value_w=1.0, conv_reg=1e-3, dense_reg=1e-3):
    p, v = model(x)
    ce = -(probs * torch.log(torch.clamp(p, min=1e-7))).sum(dim=1).mean()       # Keras clips to [eps, 1-eps]
    mse = F.mse_loss(v, target_v)
    reg = 0.0
    dense = {id(t) for lin in (model.policy_head, model.value_dense1, model.value_head) for t in (lin.weight, lin.bias)}
    for t in model.regularised():
        reg = reg + (dense_reg if id(t) in dense else conv_reg) * (t * t).sum()
    return policy_w * ce + value_w * mse + reg, ce, mse


def _batches(generator, device):
    for states, (probs, v) in generator:
        yield (torch.as_tensor(np.asarray(states, dtype=np.float32), device=device),
               torch.as_tensor(np.asarray(probs, dtype=np.float32), device=device),
               torch.as_tensor(np.asarray(v, dtype=np.float32), device=device))


def train_nn(training_data, neural_network, generator_cls, timestamp, device=None, verbose=True, **kwargs):
    """reference train_nn (:123-179).  neural_network: CheckersNet or a weight blob.  Returns
    (history dict with per-epoch 'loss' / 'val_loss' / 'lr', filename of the best epoch's blob)."""
    device = device or ("cuda" if torch.cuda.is_available() else "cpu")
    model = neural_network if isinstance(neural_network, CheckersNet) else CheckersNet(neural_network)
    model.to(device).train()
    batch = kwargs["BATCH_SIZE"]
    np.random.shuffle(training_data)
    val_data = None
    n_val = int(len(training_data) * kwargs["VAL_SPLIT"])
    if kwargs["VAL_SPLIT"] > 0 and n_val > 0:
        val_data = training_data[-n_val:]
        del training_data[-n_val:]
    step_size = max(1, int(kwargs["CLR_SS_COEFF"] * (len(training_data) / batch)))
    opt = torch.optim.Adam(model.parameters(), lr=kwargs["NN_BASE_LR"], betas=(0.9, 0.999), eps=1e-7)
    lw = dict(policy_w=kwargs.get("POLICY_LOSS_WEIGHT", 1.0), value_w=kwargs.get("VALUE_LOSS_WEIGHT", 1.0),
              conv_reg=kwargs.get("CONV_REG", 1e-3), dense_reg=kwargs.get("DENSE_REG", 1e-3))
    filepath = "data/model/Checkers_Model%d_%s.npy" % (kwargs["TRAINING_ITERATION"] + 1, timestamp)
    history = {"loss": [], "val_loss": [], "lr": []}
    ckpt_best, es_best, wait, it = np.inf, np.inf, 0, 0
    for epoch in range(kwargs["EPOCHS"]):
        order = np.random.permutation(len(training_data))                      # fit(shuffle=True)
        shuffled = [training_data[i] for i in order]
        tot, cnt = 0.0, 0
        for x, probs, tv in _batches(generator_cls(shuffled, batch), device):
            for g in opt.param_groups:
                g["lr"] = clr_triangular(it, kwargs["NN_BASE_LR"], kwargs["NN_MAX_LR"], step_size)
            loss, _ce, _mse = loss_terms(model, x, probs, tv, **lw)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            it += 1
            tot += loss.item() * len(x); cnt += len(x)
        history["loss"].append(tot / max(cnt, 1))
        history["lr"].append(opt.param_groups[0]["lr"])
        if val_data:
            model.eval()
            vt, vc = 0.0, 0
            with torch.no_grad():
                for x, probs, tv in _batches(generator_cls(val_data, batch), device):
                    loss, _ce, _mse = loss_terms(model, x, probs, tv, **lw)
                    vt += loss.item() * len(x); vc += len(x)
            model.train()
            monitor = vt / max(vc, 1)
            history["val_loss"].append(monitor)
        else:
            monitor = history["loss"][-1]
        if verbose:
            print("epoch %d: loss %.4f%s lr %.2e" % (epoch + 1, history["loss"][-1],
                                                      (" val_loss %.4f" % monitor) if val_data else "", history["lr"][-1]))
        if monitor < ckpt_best:                                                  # ModelCheckpoint(save_best_only=True)
            ckpt_best = monitor
            model.save(filepath)
        if monitor < es_best - kwargs["MIN_DELTA"]:                              # EarlyStopping(min_delta, patience)
            es_best, wait = monitor, 0
        else:
            wait += 1
            if wait >= kwargs["PATIENCE"]:
                break
    return history, filepath
