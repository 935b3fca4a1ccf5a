"""CPU tier: reader for the reference's on-disk artefacts (BSON.jl files: data.bson, best_model_weights.bson).
No Julia-written file exists in this image, so the files are produced by the matching writer, byte-checked against
hand-assembled BSON where the wire format is concerned."""
import os
import struct

import numpy as np
import pytest
import torch


def _bson(ldeq):
    from importlib import import_module
    return import_module(ldeq.__name__ + ".bson_io")


def test_wire_format_against_hand_assembled_bytes(ldeq):
    b = _bson(ldeq)
    # {"hello": "world"} from bsonspec.org, and a document with a double, an int64, a bool, a binary and a nested array
    assert b.parse(b"\x16\x00\x00\x00\x02hello\x00\x06\x00\x00\x00world\x00\x00") == {"hello": "world"}
    body = (b"\x01x\x00" + struct.pack("<d", 1.5) + b"\x12n\x00" + struct.pack("<q", -7) + b"\x08f\x00\x01" +
            b"\x05d\x00" + struct.pack("<i", 3) + b"\x00abc" +
            b"\x04a\x00" + (lambda inner: struct.pack("<i", len(inner) + 5) + inner + b"\x00")(b"\x100\x00" + struct.pack("<i", 4) + b"\x0A1\x00"))
    doc = struct.pack("<i", len(body) + 5) + body + b"\x00"
    got = b.parse(doc)
    assert got == {"x": 1.5, "n": -7, "f": True, "d": b"abc", "a": [4, None]} and isinstance(got["d"], b.Binary)
    with pytest.raises(ValueError):
        b.parse(doc[:-2])


def test_bsonjl_array_lowering_is_column_major(ldeq, tmp_path):
    b = _bson(ldeq)
    a = np.arange(6, dtype=np.float32).reshape(2, 3)           # Julia: 2x3 matrix [0 1 2; 3 4 5]
    p = str(tmp_path / "a.bson")
    b.save(p, a=a, tup=(a.astype(np.float64), [np.int64(3), 2.5, "s"]), sym=b.JuliaSymbol("relu"))
    raw = b.parse(open(p, "rb").read())
    assert raw["a"]["tag"] == "array" and raw["a"]["type"] == {"tag": "datatype", "name": ["Core", "Float32"], "params": []}
    assert raw["a"]["size"] == [2, 3]
    assert np.frombuffer(raw["a"]["data"], np.float32).tolist() == [0, 3, 1, 4, 2, 5]      # column-major bytes
    back = b.load(p)
    assert np.array_equal(back["a"], a) and back["a"].dtype == np.float32
    assert isinstance(back["tup"], tuple) and np.array_equal(back["tup"][0], a) and back["tup"][1] == [3, 2.5, "s"]
    assert back["sym"] == "relu"


def test_backrefs_are_resolved_once(ldeq):
    b = _bson(ldeq)
    arr = {"tag": "array", "type": {"tag": "datatype", "name": ["Core", "Float64"], "params": []}, "size": [2],
           "data": np.array([1.0, 2.0]).tobytes()}
    doc = b._document({"_backrefs": [arr], "x": {"tag": "backref", "ref": 1}, "y": {"tag": "tuple", "data": [{"tag": "backref", "ref": 1}]}})
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".bson", delete=False) as f:
        f.write(doc)
    got = b.load(f.name)
    os.unlink(f.name)
    assert got["x"] is got["y"][0] and got["x"].tolist() == [1.0, 2.0]


def test_data_bson_layout(ldeq, tmp_path):
    """data = (latent_data, u0s, ps, high_dim_data) as create_data.jl:30-55 builds it."""
    b = _bson(ldeq)
    rng = np.random.default_rng(0)
    N, T, H, W = 5, 7, 4, 3
    latent = [rng.standard_normal((2, T)).astype(np.float32) for _ in range(N)]
    u0s = [rng.standard_normal(2) for _ in range(N)]
    ps = [rng.uniform(1, 2, (1, 1)) for _ in range(N)]
    high = [[rng.random((H, W)).astype(np.float32) for _ in range(T)] for _ in range(N)]
    p = str(tmp_path / "data.bson")
    b.save(p, data=(latent, u0s, ps, high))
    lat, u0, pp, frames = b.load_data_bson(p)
    assert lat.shape == (N, T, 2) and u0.shape == (N, 2) and pp.shape == (N, 1) and frames.shape == (T, N, H * W)
    assert np.array_equal(lat[3, :, 1], latent[3][1]) and np.array_equal(u0[2], u0s[2]) and pp[4, 0] == ps[4][0, 0]
    # reshape(train_data, :, T, N) vectorises a frame column-major: pixel (i, j) -> i + j*H
    assert frames[6, 1, 2 + 1 * H] == high[1][6][2, 1]


def test_flux_params_roundtrip_into_the_default_model(ldeq, tmp_path):
    """weights = Flux.params(model) (model_train.jl:212-217): a Zygote.Params whose `order` buffer lists the arrays in
    functor-traversal order and whose `params` IdSet repeats them as back-references."""
    b = _bson(ldeq)
    torch.manual_seed(0)
    mt = ldeq.GOKU_basic()
    enc, dec = ldeq.default_layers(mt, 784, ldeq.Pendulum())
    model = ldeq.LatentDiffEqModel(mt, enc, dec)
    slots = b.flux_param_order(model)
    assert sum(t.numel() for t, _ in slots) == 503387 and len({t.data_ptr() for t, _ in slots}) == len(slots)
    # first arrays in Flux order: the feature extractor's Dense(784, 200) weight then bias (GOKU.jl:209-214)
    assert slots[0][1] == (200, 784) and slots[1][1] == (200,)
    rng = np.random.default_rng(1)
    arrays = [rng.standard_normal(js).astype(np.float32) for _, js in slots]
    # what BSON.jl writes: arrays once in _backrefs, Params(order = Buffer(data, freeze), params = IdSet(IdDict(...)))
    refs = [b._lower(a) for a in arrays]
    order = {"tag": "struct", "type": {"tag": "datatype", "name": ["Zygote", "Buffer"], "params": []},
             "data": [[{"tag": "backref", "ref": i + 1} for i in range(len(arrays))], False]}
    idset = {"tag": "struct", "type": {"tag": "datatype", "name": ["Base", "IdSet"], "params": []},
             "data": [{"tag": "struct", "type": {"tag": "datatype", "name": ["Base", "IdDict"], "params": []},
                       "data": [[x for i in range(len(arrays)) for x in ({"tag": "backref", "ref": i + 1}, None)], len(arrays), 0]}]}
    doc = {"_backrefs": refs, "weights": {"tag": "struct", "type": {"tag": "datatype", "name": ["Zygote", "Params"], "params": []},
                                           "data": [order, idset]}}
    p = str(tmp_path / "best_model_weights.bson")
    open(p, "wb").write(b._document(doc))
    got = b.load_flux_params(p)
    assert len(got) == len(arrays) and all(np.array_equal(g, a) for g, a in zip(got, arrays))
    b.assign_flux_params(model, got)
    assert np.array_equal(model.encoder.feature_extractor[0].weight.detach().numpy(), arrays[0])
    with pytest.raises(ValueError):
        b.assign_flux_params(model, got[:-1])
