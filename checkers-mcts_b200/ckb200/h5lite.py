"""Keras ``.h5`` -> weight blob without h5py / TensorFlow (seam S1 input side).

The reference stores its networks with ``model.save(...h5)`` (training_pipeline.py:186-191) and reads
them back with ``load_model`` (:345, 515-516).  Those files are HDF5 with a version-0 superblock,
version-1 object headers and contiguous, uncompressed float32 datasets (TF 2.2 / h5py "earliest"
format), which is a small enough subset to read directly:

    superblock -> root symbol-table entry -> group B-tree (v1) + local heap -> symbol nodes (SNOD)
    object header (v1, with continuation blocks) -> dataspace / datatype / contiguous layout messages

``read_datasets`` returns every dataset below ``/model_weights`` by path; ``keras_h5_to_blob`` maps
them onto the flat blob of ``ckb200.net.layout`` using the layer graph in the file's ``model_config``
attribute (layer names such as ``conv2d_7`` depend on how many models the saving process had built, so
layers are identified by topology and shape, not by name).
"""
import json
import struct

import numpy as np

from . import net as _N

_SIG = b"\x89HDF\r\n\x1a\n"


class H5Error(ValueError):
    pass


class _File(object):
    def __init__(self, data):
        self.b = data
        if data[:8] != _SIG:
            raise H5Error("not an HDF5 file")
        if data[8] != 0:
            raise H5Error("only version-0 superblocks are supported (got %d)" % data[8])
        self.O, self.L = data[13], data[14]                 # size of offsets / lengths
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets/lengths are supported")
        # superblock v0: 8 sig, 8 version bytes, 2+2 group K, 4 flags, then base, free-space, EOF, driver addresses
        p = 24
        self.base = self.u64(p)
        p += 4 * self.O
        self.root = self.symbol_entry(p)

    def u16(self, p): return struct.unpack_from("<H", self.b, p)[0]
    def u32(self, p): return struct.unpack_from("<I", self.b, p)[0]
    def u64(self, p): return struct.unpack_from("<Q", self.b, p)[0]

    def symbol_entry(self, p):
        """-> dict(name_off, header, cache, btree, heap)"""
        e = {"name_off": self.u64(p), "header": self.u64(p + 8), "cache": self.u32(p + 16)}
        if e["cache"] == 1:
            e["btree"], e["heap"] = self.u64(p + 24), self.u64(p + 32)
        return e

    # ---- groups ---------------------------------------------------------------------------------
    def heap_data(self, addr):
        a = self.base + addr
        if self.b[a:a + 4] != b"HEAP":
            raise H5Error("bad local heap signature at %d" % a)
        return self.base + self.u64(a + 8 + 2 * self.L)

    def group_entries(self, btree, heap):
        """-> list of (name, symbol entry) below a v1 group B-tree"""
        hd = self.heap_data(heap)
        out = []

        def walk(addr):
            a = self.base + addr
            if self.b[a:a + 4] != b"TREE":
                raise H5Error("bad B-tree signature at %d" % a)
            ntype, level, used = self.b[a + 4], self.b[a + 5], self.u16(a + 6)
            if ntype != 0:
                raise H5Error("not a group B-tree")
            p = a + 8 + 2 * self.O
            for i in range(used):
                child = self.u64(p + self.L + i * (self.L + self.O))      # key, child, key, child, ..., key
                if level > 0:
                    walk(child)
                else:
                    s = self.base + child
                    if self.b[s:s + 4] != b"SNOD":
                        raise H5Error("bad symbol node signature at %d" % s)
                    n = self.u16(s + 6)
                    for j in range(n):
                        e = self.symbol_entry(s + 8 + j * (2 * self.O + 24))
                        q = hd + e["name_off"]
                        name = self.b[q:self.b.index(b"\x00", q)].decode()
                        out.append((name, e))

        walk(btree)
        return out

    # ---- object headers --------------------------------------------------------------------------
    def messages(self, addr):
        """-> list of (type, bytes) of a version-1 object header, continuation blocks included"""
        a = self.base + addr
        if self.b[a] != 1:
            raise H5Error("only version-1 object headers are supported (got %d at %d)" % (self.b[a], a))
        nmsg, size = self.u16(a + 2), self.u32(a + 8)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self.u16(p), self.u16(p + 2)
                body = self.b[p + 8:p + 8 + msize]
                if mtype == 0x10:                                        # continuation
                    blocks.append((self.base + struct.unpack_from("<Q", body, 0)[0], struct.unpack_from("<Q", body, 8)[0]))
                out.append((mtype, body))
                p += 8 + msize
        return out

    def children(self, entry):
        """group entry -> list of (name, entry); [] for datasets"""
        if entry.get("cache") == 1:
            return self.group_entries(entry["btree"], entry["heap"])
        for mtype, body in self.messages(entry["header"]):
            if mtype == 0x11:                                            # symbol table message
                return self.group_entries(struct.unpack_from("<Q", body, 0)[0], struct.unpack_from("<Q", body, 8)[0])
        return []

    def dataset_extent(self, entry):
        """-> (file offset, nbytes, shape, dtype) of a contiguous little-endian float dataset, else None"""
        shape = dtype = addr = size = None
        for mtype, body in self.messages(entry["header"]):
            if mtype == 0x1:
                ver, rank = body[0], body[1]
                off = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank))
            elif mtype == 0x3:
                cls, bits0 = body[0] & 0x0F, body[1]
                sz = struct.unpack_from("<I", body, 4)[0]
                if cls == 1 and not (bits0 & 1) and sz in (4, 8):
                    dtype = np.dtype("<f%d" % sz)
            elif mtype == 0x8 and body[0] == 3 and body[1] == 1:
                addr, size = struct.unpack_from("<Q", body, 2)[0], struct.unpack_from("<Q", body, 10)[0]
        if shape is None or dtype is None or addr is None or addr == 0xFFFFFFFFFFFFFFFF:
            return None
        return self.base + addr, size, shape, dtype

    def dataset(self, entry):
        """-> numpy array, or None when the object is not a contiguous little-endian float dataset"""
        shape = dtype = None
        addr = size = None
        for mtype, body in self.messages(entry["header"]):
            if mtype == 0x1:                                             # dataspace
                ver, rank = body[0], body[1]
                off = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", body, off + 8 * i)[0] for i in range(rank))
            elif mtype == 0x3:                                           # datatype
                cls, bits0 = body[0] & 0x0F, body[1]
                sz = struct.unpack_from("<I", body, 4)[0]
                if cls == 1 and not (bits0 & 1) and sz in (4, 8):        # IEEE float, little-endian
                    dtype = np.dtype("<f%d" % sz)
            elif mtype == 0x8:                                           # layout
                if body[0] == 3 and body[1] == 1:                        # version 3, contiguous
                    addr, size = struct.unpack_from("<Q", body, 2)[0], struct.unpack_from("<Q", body, 10)[0]
                elif body[0] == 3 and body[1] == 0:                      # compact: data inside the message
                    n = struct.unpack_from("<H", body, 2)[0]
                    addr, size = ("compact", body[4:4 + n]), n
        if shape is None or dtype is None or addr is None:
            return None
        if isinstance(addr, tuple):
            raw = addr[1]
        else:
            if addr == 0xFFFFFFFFFFFFFFFF:
                return np.zeros(shape, dtype=dtype)
            raw = self.b[self.base + addr:self.base + addr + size]
        n = int(np.prod(shape)) if shape else 1
        return np.frombuffer(raw, dtype=dtype, count=n).reshape(shape).copy()


def read_datasets(path, prefix="model_weights"):
    """-> (dict dataset path -> array for everything below /<prefix>, raw file bytes)"""
    with open(path, "rb") as f:
        data = f.read()
    h5 = _File(data)
    out = {}

    def walk(entry, name):
        kids = h5.children(entry)
        if kids:
            for k, e in kids:
                walk(e, name + "/" + k if name else k)
        else:
            arr = h5.dataset(entry)
            if arr is not None:
                out[name] = arr

    for k, e in h5.children(h5.root):
        if prefix is None or k == prefix:
            walk(e, k)
    return out, data


def model_config(data):
    """the ``model_config`` attribute (a variable-length string in the global heap): found by its JSON
    prefix rather than through the attribute machinery"""
    for key in (b'{"class_name": "Model"', b'{"class_name": "Functional"'):
        i = data.find(key)
        if i >= 0:
            cfg, _end = json.JSONDecoder().raw_decode(data[i:i + (1 << 22)].decode("utf-8", "replace"))
            return cfg
    raise H5Error("model_config not found")


def _inbound(layer):
    return [n[0] for node in layer.get("inbound_nodes", []) for n in node]


def layer_roles(cfg):
    """layer graph of create_nn (training_pipeline.py:44-120) -> dict role -> Keras layer name, roles as in
    ckb200.net.layout: conv0..conv6, policy_conv1/2, value_conv1, value_dense1, policy_head, value_head and
    '<role>/bn' for the BatchNormalization that follows a layer"""
    layers = cfg["config"]["layers"]
    consumers = {}
    for l in layers:
        for src in _inbound(l):
            consumers.setdefault(src, []).append(l)
    inp = [l for l in layers if l["class_name"] == "InputLayer"]
    if len(inp) != 1:
        raise H5Error("expected one InputLayer")
    roles = {}

    def only(src, cls, pred=lambda l: True):
        c = [l for l in consumers.get(src, []) if l["class_name"] == cls and pred(l)]
        if len(c) != 1:
            raise H5Error("layer graph does not match create_nn after %r (%s x%d)" % (src, cls, len(c)))
        return c[0]

    def ksize(l): return tuple(l["config"]["kernel_size"])

    cur = inp[0]["name"]
    for i in range(7):
        c = only(cur, "Conv2D", lambda l: ksize(l) == (3, 3) and l["config"]["filters"] == 128 and (i > 0 or True))
        b = only(c["name"], "BatchNormalization")
        roles["conv%d" % i], roles["conv%d/bn" % i] = c["name"], b["name"]
        cur = b["name"]
    p1 = only(cur, "Conv2D", lambda l: ksize(l) == (3, 3))
    roles["policy_conv1"], roles["policy_conv1/bn"] = p1["name"], only(p1["name"], "BatchNormalization")["name"]
    p2 = only(roles["policy_conv1/bn"], "Conv2D", lambda l: ksize(l) == (1, 1) and l["config"]["filters"] == 8)
    roles["policy_conv2"], roles["policy_conv2/bn"] = p2["name"], only(p2["name"], "BatchNormalization")["name"]
    fl = only(roles["policy_conv2/bn"], "Flatten")
    roles["policy_head"] = only(fl["name"], "Dense", lambda l: l["config"]["units"] == 512)["name"]
    v1 = only(cur, "Conv2D", lambda l: ksize(l) == (1, 1) and l["config"]["filters"] == 1)
    roles["value_conv1"], roles["value_conv1/bn"] = v1["name"], only(v1["name"], "BatchNormalization")["name"]
    fl = only(roles["value_conv1/bn"], "Flatten")
    d1 = only(fl["name"], "Dense", lambda l: l["config"]["units"] == 64)
    roles["value_dense1"], roles["value_dense1/bn"] = d1["name"], only(d1["name"], "BatchNormalization")["name"]
    roles["value_head"] = only(roles["value_dense1/bn"], "Dense", lambda l: l["config"]["units"] == 1)["name"]
    for l in layers:
        if l["class_name"] == "BatchNormalization":
            eps = l["config"].get("epsilon", 1e-3)
            if abs(eps - _N.BN_EPS) > 1e-9:
                raise H5Error("BatchNormalization epsilon %g is not the %g the kernels fold" % (eps, _N.BN_EPS))
    return roles


def keras_h5_to_blob(path):
    """reference model file -> float32 blob for ck_net_set_weights (ckb200.net.layout order)"""
    dsets, data = read_datasets(path)
    roles = layer_roles(model_config(data))
    by_layer = {}
    for name, arr in dsets.items():
        parts = name.split("/")
        if len(parts) >= 3:
            by_layer.setdefault(parts[1], {})[parts[-1].split(":")[0]] = arr
    blob = np.zeros(_N.NET_PARAM_COUNT, dtype=np.float32)
    bn_names = {"bn_gamma": "gamma", "bn_beta": "beta", "bn_mean": "moving_mean", "bn_var": "moving_variance"}
    for key, (off, shape) in _N.layout().items():
        role, param = key.split("/")
        if param in bn_names:
            layer, pname = roles[role + "/bn"], bn_names[param]
        else:
            layer, pname = roles[role], param
        try:
            arr = by_layer[layer][pname]
        except KeyError:
            raise H5Error("dataset %s/%s (for %s) is missing" % (layer, pname, key))
        if tuple(arr.shape) != tuple(shape):
            raise H5Error("%s/%s has shape %s, expected %s" % (layer, pname, arr.shape, shape))
        blob[off:off + arr.size] = arr.astype(np.float32).reshape(-1)
    return blob


def blob_to_keras_h5(blob, template_path, out_path):
    """Writes ``blob`` as a Keras ``.h5`` model file the reference's ``load_model`` reads (training_pipeline.py:186-191,
    345, 515-516) WITHOUT h5py / TensorFlow: ``template_path`` is any model file the reference itself saved for this
    architecture (``data/model/Checkers_Model*.h5``); its HDF5 structure, ``model_config`` and training configuration
    are kept byte for byte and only the payloads of the weight datasets (contiguous, uncompressed float32) are replaced.
    The optimizer state that Keras stored next to the weights is zeroed (a fresh Adam).  -> out_path"""
    blob = np.ascontiguousarray(blob, dtype=np.float32).reshape(-1)
    if blob.size != _N.NET_PARAM_COUNT:
        raise H5Error("weight blob has %d values, expected %d" % (blob.size, _N.NET_PARAM_COUNT))
    with open(template_path, "rb") as f:
        data = bytearray(f.read())
    h5 = _File(bytes(data))
    extents = {}

    def walk(entry, name):
        kids = h5.children(entry)
        if kids:
            for k, e in kids:
                walk(e, name + "/" + k if name else k)
        else:
            ext = h5.dataset_extent(entry)
            if ext is not None:
                extents[name] = ext

    for k, e in h5.children(h5.root):
        walk(e, k)
    roles = layer_roles(model_config(bytes(data)))
    by_layer = {}
    for name, ext in extents.items():
        parts = name.split("/")
        if parts[0] == "model_weights" and len(parts) >= 3:
            by_layer.setdefault(parts[1], {})[parts[-1].split(":")[0]] = ext
        elif parts[0] == "optimizer_weights":
            off, size, _shape, _dt = ext
            data[off:off + size] = bytes(size)                           # Adam moments of the template's training run
    bn_names = {"bn_gamma": "gamma", "bn_beta": "beta", "bn_mean": "moving_mean", "bn_var": "moving_variance"}
    for key, (boff, shape) in _N.layout().items():
        role, param = key.split("/")
        layer, pname = (roles[role + "/bn"], bn_names[param]) if param in bn_names else (roles[role], param)
        try:
            off, size, fshape, dt = by_layer[layer][pname]
        except KeyError:
            raise H5Error("template has no dataset %s/%s (for %s)" % (layer, pname, key))
        n = int(np.prod(shape))
        if tuple(fshape) != tuple(shape) or size != n * dt.itemsize:
            raise H5Error("template dataset %s/%s has shape %s, expected %s" % (layer, pname, fshape, shape))
        data[off:off + size] = blob[boff:boff + n].astype(dt).tobytes()
    with open(out_path, "wb") as f:
        f.write(data)
    return out_path
