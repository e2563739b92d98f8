"""Reader for the reference's BSON.jl checkpoints (utils.jl:179-196, 408-418; compute_D.jl:84).

Host-side only.  Julia structs are {tag:"struct", type:{tag:"datatype", name:[...]}, data:[fields]}; bits-type
arrays are {tag:"array", type, size:[d1,d2,...], data:<binary>} little-endian COLUMN-major; shared types are
{tag:"backref", ref:n} (1-based into the top-level `_backrefs`).  In a Flux.Chain the layers sit in
data[0].data; Conv fields are (sigma, weight, bias, stride, pad, dilation, groups), Dense (weight, bias, sigma).
"""
import struct

import numpy as np

_DT = {"Float32": "<f4", "Float64": "<f8", "Int64": "<i8", "Int32": "<i4", "UInt8": "u1", "Bool": "u1", "UInt64": "<u8"}


def _cstring(b, o):
    e = b.index(b"\x00", o)
    return b[o:e].decode("utf8"), e + 1


def parse_document(b, o=0, as_list=False):
    (n,) = struct.unpack_from("<i", b, o)
    end = o + n
    o += 4
    out = [] if as_list else {}
    while b[o] != 0:
        t = b[o]
        o += 1
        k, o = _cstring(b, o)
        if t == 0x01:
            (v,) = struct.unpack_from("<d", b, o); o += 8
        elif t == 0x02:
            (ln,) = struct.unpack_from("<i", b, o); o += 4
            v = b[o:o + ln - 1].decode("utf8"); o += ln
        elif t == 0x03:
            v, o = parse_document(b, o)
        elif t == 0x04:
            v, o = parse_document(b, o, as_list=True)
        elif t == 0x05:
            (ln,) = struct.unpack_from("<i", b, o); o += 5
            v = bytes(b[o:o + ln]); o += ln
        elif t == 0x08:
            v = bool(b[o]); o += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            (v,) = struct.unpack_from("<i", b, o); o += 4
        elif t == 0x12:
            (v,) = struct.unpack_from("<q", b, o); o += 8
        else:
            raise ValueError("unsupported BSON element type 0x%02x" % t)
        if as_list:
            out.append(v)
        else:
            out[k] = v
    if o + 1 != end:
        raise ValueError("corrupt BSON document")
    return out, end


class BsonFile:
    def __init__(self, path_or_bytes):
        raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
        try:
            self.doc, _ = parse_document(raw)
        except (IndexError, struct.error, UnicodeDecodeError, KeyError, OverflowError, MemoryError) as e:
            raise ValueError("corrupt or truncated BSON document (%s: %s)" % (type(e).__name__, e)) from e
        self.backrefs = self.doc.get("_backrefs", [])

    def deref(self, x):
        while isinstance(x, dict) and x.get("tag") == "backref":
            x = self.backrefs[x["ref"] - 1]
        return x

    def type_name(self, x):
        t = self.deref(x.get("type")) if isinstance(x, dict) and "type" in x else None
        if isinstance(t, dict) and t.get("tag") == "datatype":
            return ".".join(t["name"])
        return None

    def array(self, x):
        """Julia bits array -> numpy array with Julia's shape (column-major semantics preserved)."""
        x = self.deref(x)
        if not (isinstance(x, dict) and x.get("tag") == "array"):
            raise TypeError("not a Julia array")
        el = (self.type_name(x) or "").split(".")[-1]
        if el not in _DT:
            raise TypeError("unsupported element type %r" % el)
        shape = tuple(int(d) for d in x["size"])
        return np.frombuffer(x["data"], dtype=_DT[el]).reshape(shape, order="F").copy()

    def fields(self, x):
        x = self.deref(x)
        return [self.deref(f) for f in x["data"]]


def chain_layers(bf, chain):
    """[(kind, dict)] for a Flux.Chain object: conv {W (k1,k2,cin,cout), b, pad, stride}, dense {W (out,in), b}."""
    chain = bf.deref(chain)
    if not (bf.type_name(chain) or "").endswith("Chain"):
        raise TypeError("not a Flux.Chain: %r" % bf.type_name(chain))
    layers = []
    for layer in bf.fields(bf.fields(chain)[0]):
        tn = bf.type_name(layer) or ""
        if tn.endswith("Conv"):
            f = bf.fields(layer)
            pad = [int(v) for v in bf.deref(f[4])["data"]]
            layers.append(("conv", {"W": bf.array(f[1]), "b": bf.array(f[2]), "pad": pad,
                                    "stride": [int(v) for v in bf.deref(f[3])["data"]], "act": bf.type_name(f[0])}))
        elif tn.endswith("Dense"):
            f = bf.fields(layer)
            layers.append(("dense", {"W": bf.array(f[0]), "b": bf.array(f[1]), "act": bf.type_name(f[2])}))
        elif "flatten" in tn:
            layers.append(("flatten", {}))
        else:
            raise TypeError("unsupported layer %r" % tn)
    return layers


def load_trainer_nets(path):
    """(q_net layers, t_net layers) from ./trainers/<name>.bson (key `tr`, Trainer.model = DQNModel(q_net, t_net, opt))."""
    bf = BsonFile(path)
    tr = bf.deref(bf.doc["tr"])
    model = None
    for f in bf.fields(tr):
        if isinstance(f, dict) and (bf.type_name(f) or "").endswith("DQNModel"):
            model = f
    if model is None:
        raise ValueError("no DQNModel in trainer")
    mf = bf.fields(model)
    return chain_layers(bf, mf[0]), chain_layers(bf, mf[1])


def load_deviation_matrix(path):
    """The centred D of compute_D.jl:84 (`BSON.@save ... deviation_matrix = deviation_matrix`), read back the way
    plot_traj.jl:7 does: a (P, K) Float64 numpy array in Julia's shape.  `DeviationMatrix.from_numpy` takes it."""
    bf = BsonFile(path)
    if "deviation_matrix" not in bf.doc:
        raise ValueError("no `deviation_matrix` in %s (keys: %s)" % (path, sorted(k for k in bf.doc if k != "_backrefs")))
    D = bf.array(bf.doc["deviation_matrix"])
    if D.ndim != 2 or D.dtype != np.float64:
        raise ValueError("deviation_matrix must be a Float64 matrix, found %s %s" % (D.dtype, D.shape))
    return D


def destructure(layers):
    """Flux.destructure order: per layer weight then bias, each column-major (compute_D.jl:43,68)."""
    parts = []
    for kind, p in layers:
        if kind in ("conv", "dense"):
            parts += [p["W"].reshape(-1, order="F"), p["b"].reshape(-1, order="F")]
    return np.concatenate(parts)
