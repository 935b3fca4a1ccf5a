"""Readers (and a matching writer) for the reference's on-disk artefacts: BSON files written by BSON.jl.

The reference stores two things:

* ``data/data.bson`` -- ``@save data_path data`` with ``data = (latent_data, u0s, ps, high_dim_data)``
  (``examples/pendulum_friction-less/model_train.jl:84-99``, produced by ``create_data.jl:30-55``):
  ``latent_data::Vector{Matrix{Float32}}`` (z x T per trajectory), ``u0s::Vector{Vector{Float64}}``,
  ``ps::Vector{Matrix{Float64}}`` (p x 1) and ``high_dim_data::Vector{Vector{Matrix{Float32}}}`` (28 x 28 frames);
* ``output/best_model_weights.bson`` -- ``@save ... weights`` with ``weights = Flux.params(model)``
  (``model_train.jl:212-217``): a ``Zygote.Params`` whose ``order`` buffer lists the arrays in functor-traversal order.

Format: the BSON wire format (bsonspec.org: little-endian documents of typed, named elements) plus BSON.jl's lowering
of Julia values [3P BSON.jl 0.3, restated from its published conventions]: a dense array of a bits type is the document
``{tag: "array", type: {tag: "datatype", name: [module path..., type name], params: [...]}, size: [...], data: <binary>}``
with the bytes in Julia's column-major order; vectors of non-bits values are plain BSON arrays; tuples are
``{tag: "tuple", data: [...]}``; structs ``{tag: "struct", type: ..., data: [fields...]}``; symbols ``{tag: "symbol", name}``;
an object referenced more than once is stored once in the top-level ``_backrefs`` list and referenced as
``{tag: "backref", ref: i}`` (1-based).

No Julia-written file exists in this image (no Julia, no network), so the reader is exercised against files produced by
the writer below, which emits exactly the structures described above; ``julia/make_golden.jl`` is the script a maintainer
runs once under real Julia to produce ``tests/golden/julia_*.bson``, which ``tests/test_golden.py`` then consumes.
"""
from __future__ import annotations

import struct
from typing import Any

import numpy as np

_JL_DTYPES = {"Float32": np.float32, "Float64": np.float64, "Float16": np.float16, "Int64": np.int64, "Int32": np.int32,
              "Int16": np.int16, "Int8": np.int8, "UInt64": np.uint64, "UInt32": np.uint32, "UInt16": np.uint16,
              "UInt8": np.uint8, "Bool": np.bool_}
_NP_JL = {np.dtype(v): k for k, v in _JL_DTYPES.items()}


# ---- BSON wire format ---------------------------------------------------------------------------------------------
class Binary(bytes):
    """A BSON binary element (subtype 0)."""


def _cstring(buf: memoryview, pos: int):
    end = pos
    while buf[end] != 0:
        end += 1
    return bytes(buf[pos:end]).decode("utf-8"), end + 1


def _parse_document(buf: memoryview, pos: int, as_list: bool = False):
    (size,) = struct.unpack_from("<i", buf, pos)
    end = pos + size
    if size < 5 or end > len(buf) or buf[end - 1] != 0:
        raise ValueError("malformed BSON document")
    pos += 4
    out: Any = [] if as_list else {}
    while pos < end - 1:
        etype = buf[pos]
        name, pos = _cstring(buf, pos + 1)
        if etype == 0x01:
            (val,) = struct.unpack_from("<d", buf, pos); pos += 8
        elif etype == 0x02:
            (n,) = struct.unpack_from("<i", buf, pos)
            val = bytes(buf[pos + 4:pos + 4 + n - 1]).decode("utf-8"); pos += 4 + n
        elif etype == 0x03:
            val, pos = _parse_document(buf, pos)
        elif etype == 0x04:
            val, pos = _parse_document(buf, pos, as_list=True)
        elif etype == 0x05:
            (n,) = struct.unpack_from("<i", buf, pos)
            val = Binary(bytes(buf[pos + 5:pos + 5 + n])); pos += 5 + n
        elif etype == 0x08:
            val = buf[pos] != 0; pos += 1
        elif etype == 0x0A:
            val = None
        elif etype == 0x10:
            (val,) = struct.unpack_from("<i", buf, pos); pos += 4
        elif etype == 0x12:
            (val,) = struct.unpack_from("<q", buf, pos); pos += 8
        elif etype == 0x09 or etype == 0x11:
            (val,) = struct.unpack_from("<q", buf, pos); pos += 8
        else:
            raise ValueError(f"unsupported BSON element type 0x{etype:02x} for key {name!r}")
        if as_list:
            out.append(val)
        else:
            out[name] = val
    return out, end


def parse(data: bytes) -> dict:
    """Raw BSON document -> nested dict / list / scalars / :class:`Binary` (no Julia interpretation)."""
    doc, end = _parse_document(memoryview(data), 0)
    if end != len(data):
        raise ValueError("trailing bytes after the BSON document")
    return doc


def _emit(name: str, val, out: bytearray):
    key = name.encode("utf-8") + b"\x00"
    if isinstance(val, bool):
        out += b"\x08" + key + (b"\x01" if val else b"\x00")
    elif isinstance(val, (int, np.integer)):
        out += b"\x12" + key + struct.pack("<q", int(val))
    elif isinstance(val, (float, np.floating)):
        out += b"\x01" + key + struct.pack("<d", float(val))
    elif isinstance(val, str):
        b = val.encode("utf-8") + b"\x00"
        out += b"\x02" + key + struct.pack("<i", len(b)) + b
    elif isinstance(val, (bytes, bytearray)):
        out += b"\x05" + key + struct.pack("<i", len(val)) + b"\x00" + bytes(val)
    elif val is None:
        out += b"\x0A" + key
    elif isinstance(val, dict):
        out += b"\x03" + key + _document(val)
    elif isinstance(val, (list, tuple)):
        out += b"\x04" + key + _document({str(i): v for i, v in enumerate(val)})
    else:
        raise TypeError(f"cannot encode {type(val)} as BSON")


def _document(d: dict) -> bytes:
    body = bytearray()
    for k, v in d.items():
        _emit(k, v, body)
    return struct.pack("<i", len(body) + 5) + bytes(body) + b"\x00"


# ---- BSON.jl lowering <-> Python values ---------------------------------------------------------------------------
class JuliaStruct:
    """A Julia struct BSON.jl could not map to a Python value: its type path and field values."""

    def __init__(self, type_name: list, fields: list, params: list | None = None):
        self.type_name, self.fields, self.params = type_name, fields, params or []

    def __repr__(self):
        return f"JuliaStruct({'.'.join(self.type_name)}, {len(self.fields)} fields)"


class JuliaSymbol(str):
    pass


def _type_name(t) -> list:
    if isinstance(t, dict) and t.get("tag") == "datatype":
        return list(t["name"])
    raise ValueError(f"not a lowered DataType: {t!r}")


def _raise(v, backrefs, cache):
    if isinstance(v, list):
        return [_raise(x, backrefs, cache) for x in v]
    if not isinstance(v, dict):
        return v
    tag = v.get("tag")
    if tag == "backref":
        i = int(v["ref"]) - 1
        if i not in cache:
            cache[i] = _raise(backrefs[i], backrefs, cache)
        return cache[i]
    if tag == "array":
        size = [int(s) for s in v["size"]]
        data = v["data"]
        if isinstance(data, Binary):
            name = _type_name(v["type"])
            if name[-1] not in _JL_DTYPES:
                raise ValueError(f"array of unsupported bits type {'.'.join(name)}")
            a = np.frombuffer(bytes(data), dtype=_JL_DTYPES[name[-1]])
            return a.reshape(size, order="F")     # Julia arrays are column-major: shape == Julia's size(x)
        items = [_raise(x, backrefs, cache) for x in data]
        out = np.empty(len(items), dtype=object)
        out[:] = items
        return out.reshape(size, order="F")
    if tag == "tuple":
        return tuple(_raise(x, backrefs, cache) for x in v["data"])
    if tag == "symbol":
        return JuliaSymbol(v["name"])
    if tag == "datatype":
        return JuliaStruct(["Core", "DataType"], [list(v["name"])], [_raise(p, backrefs, cache) for p in v.get("params", [])])
    if tag == "struct":
        t = v["type"]
        data = v["data"]
        fields = [_raise(x, backrefs, cache) for x in data] if isinstance(data, list) else [data]
        return JuliaStruct(_type_name(t), fields, [_raise(p, backrefs, cache) for p in t.get("params", [])])
    return {k: _raise(x, backrefs, cache) for k, x in v.items() if k != "_backrefs"}


def load(path: str) -> dict:
    """``BSON.load(path)``: the top-level dictionary with Julia values raised to numpy arrays (``shape == size(x)``,
    Fortran order), tuples, lists, :class:`JuliaSymbol` and :class:`JuliaStruct`."""
    doc = parse(open(path, "rb").read())
    backrefs = doc.get("_backrefs", [])
    return _raise(doc, backrefs, {})


def _lower(v):
    if isinstance(v, np.ndarray) and v.dtype != object:
        name = _NP_JL[v.dtype]
        return {"tag": "array", "type": {"tag": "datatype", "name": ["Core", name], "params": []},
                "size": [int(s) for s in v.shape], "data": np.asfortranarray(v).tobytes(order="F")}
    if isinstance(v, tuple):
        return {"tag": "tuple", "data": [_lower(x) for x in v]}
    if isinstance(v, list):
        return [_lower(x) for x in v]
    if isinstance(v, JuliaStruct):
        return {"tag": "struct", "type": {"tag": "datatype", "name": list(v.type_name), "params": []},
                "data": [_lower(x) for x in v.fields]}
    if isinstance(v, JuliaSymbol):
        return {"tag": "symbol", "name": str(v)}
    if isinstance(v, dict):
        return {k: _lower(x) for k, x in v.items()}
    return v


def save(path: str, **named_values):
    """``BSON.@save path a b ...`` for numpy arrays (``shape`` = Julia size), tuples, lists, scalars and JuliaStructs."""
    with open(path, "wb") as f:
        f.write(_document({k: _lower(v) for k, v in named_values.items()}))


# ---- the reference's two artefacts --------------------------------------------------------------------------------
def load_data_bson(path: str):
    """``data.bson`` of the reference (``model_train.jl:84-99``) in this package's array convention.

    Returns ``(latent_data [N, T, z], u0s [N, z], ps [N, p], frames [T, N, H*W])``: ``frames`` is what the training
    script builds from ``high_dim_data`` (``model_train.jl:100-110``: stack time, stack samples, ``reshape(:, T, N)``) in
    the torch layout ``[T, B, P]`` of Julia's ``(P, B, T)`` batches."""
    data = load(path)["data"]
    latent, u0s, ps, high = data
    latent = np.stack([np.asarray(a, dtype=np.float32).T for a in latent])                     # (z,T) -> [T,z]
    u0s = np.stack([np.asarray(u, dtype=np.float64).reshape(-1) for u in u0s])
    ps = np.stack([np.asarray(p, dtype=np.float64).reshape(-1) for p in ps])
    # a frame is an (h, w) matrix; reshape(train_data, :, T, N) flattens it column-major
    frames = np.stack([np.stack([np.asarray(fr, dtype=np.float32).reshape(-1, order="F") for fr in traj]) for traj in high])  # [N,T,P]
    return latent, u0s, ps, np.ascontiguousarray(frames.transpose(1, 0, 2))


def _collect_arrays(v, out):
    if isinstance(v, np.ndarray) and v.dtype != object:
        out.append(v)
    elif isinstance(v, np.ndarray):
        for x in v.reshape(-1, order="F"):
            _collect_arrays(x, out)
    elif isinstance(v, (list, tuple)):
        for x in v:
            _collect_arrays(x, out)
    elif isinstance(v, JuliaStruct):
        for x in v.fields:
            _collect_arrays(x, out)
    elif isinstance(v, dict):
        for x in v.values():
            _collect_arrays(x, out)


def load_flux_params(path: str, key: str = "weights") -> list:
    """The arrays of a saved ``Flux.params(model)`` (``model_train.jl:212-217``) in parameter order.

    ``Zygote.Params`` has the fields ``order::Buffer`` (whose ``data`` vector lists the arrays in functor-traversal order)
    and ``params::IdSet`` (the same arrays; BSON.jl stores them as back-references).  The ``order`` field comes first, so
    the first occurrence of every array is its position in the traversal."""
    v = load(path)[key]
    seen, out = set(), []
    arrays: list = []
    if isinstance(v, JuliaStruct) and v.type_name[-1] == "Params":
        _collect_arrays(v.fields[0], arrays)      # order.data
    else:
        _collect_arrays(v, arrays)
    for a in arrays:
        if id(a) not in seen:
            seen.add(id(a))
            out.append(a)
    return out


def flux_param_order(model) -> list:
    """This package's parameters in the order ``Flux.params`` visits the reference model: encoder (feature extractor,
    pattern extractor, latent_in) then decoder (latent_out, diffeq, reconstructor), every layer's arrays in Flux's field
    order -- ``Dense``: weight, bias; ``RNN``: Wi, Wh, b, state0; ``LSTM``: Wi, Wh, b, state0 = (h, c).  Returns
    ``[(tensor, julia_shape)]`` where ``julia_shape`` is the size the Julia array has."""
    from .model import LSTM, RNN, Dense
    out = []

    def visit(m):
        if isinstance(m, Dense):
            out.append((m.weight, tuple(m.weight.shape)))                 # (out, in) in both languages
            out.append((m.bias, tuple(m.bias.shape)))
        elif isinstance(m, RNN):
            out.extend([(m.Wi, tuple(m.Wi.shape)), (m.Wh, tuple(m.Wh.shape)), (m.b, tuple(m.b.shape)),
                        (m.state0, (m.state0.shape[0], 1))])              # state0 is an (out, 1) matrix in Flux
        elif isinstance(m, LSTM):
            out.extend([(m.Wi, tuple(m.Wi.shape)), (m.Wh, tuple(m.Wh.shape)), (m.b, tuple(m.b.shape)),
                        (m.h0, (m.h0.shape[0], 1)), (m.c0, (m.c0.shape[0], 1))])
        else:
            for c in m.children():
                visit(c)
    visit(model)
    return out


def assign_flux_params(model, arrays: list):
    """Copy reference-trained weights (``load_flux_params``) into a model of this package; shapes must match one to one."""
    import torch
    slots = flux_param_order(model)
    if len(slots) != len(arrays):
        raise ValueError(f"the file holds {len(arrays)} arrays, the model has {len(slots)} parameter arrays")
    with torch.no_grad():
        for i, ((tensor, jshape), a) in enumerate(zip(slots, arrays)):
            if tuple(a.shape) != tuple(jshape) and tuple(a.shape) != tuple(tensor.shape):
                raise ValueError(f"parameter {i}: file has size {a.shape}, the model expects {jshape}")
            # a Julia (out, in) matrix read in Fortran order has the same [out, in] indexing as the torch weight
            tensor.copy_(torch.from_numpy(np.array(a, order="C")).reshape(tensor.shape))
