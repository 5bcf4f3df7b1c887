"""Caffe binary weights front end (caffemodel.py): wire-format round trip, Caffe's BatchNorm factor convention, shape
checks, and that a plan built from a .caffemodel is byte-identical to the plan built from the same weights in memory
(reference role: trtexec --deploy=<prototxt> --model=<caffemodel>, models/setup.py:53-55)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tensorrt_laboratory_b200 import builder, caffemodel, graph, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mini_net():
    L = [dict(name="conv1", type="Convolution", bottoms=["data"], tops=["conv1"], num_output=64, kernel_size=3, pad=1, stride=1, bias_term=False),
         dict(name="bn1", type="BatchNorm", bottoms=["conv1"], tops=["conv1"]),
         dict(name="scale1", type="Scale", bottoms=["conv1"], tops=["conv1"], bias_term=True),
         dict(name="relu1", type="ReLU", bottoms=["conv1"], tops=["conv1"]),
         dict(name="pool", type="Pooling", bottoms=["conv1"], tops=["pool"], pool="AVE", kernel_size=16, stride=1, pad=0),
         dict(name="fc", type="InnerProduct", bottoms=["pool"], tops=["fc"], num_output=10, bias_term=True),
         dict(name="prob", type="Softmax", bottoms=["fc"], tops=["prob"])]
    return {"name": "mini", "input": "data", "input_dims": [1, 3, 16, 16], "layers": L}


def test_round_trip_and_bn_factor():
    net = _mini_net()
    w = weights.random_weights(net, 7)
    for factor in (1.0, 0.999, 37.5):
        got = caffemodel.load_caffemodel(caffemodel.save_caffemodel(net, w, bn_factor=factor), net)
        assert set(got) == set(w)
        for name in w:
            for f in w[name]:
                np.testing.assert_allclose(got[name][f], w[name][f], rtol=3e-7, atol=0, err_msg=f"{name}.{f} factor {factor}")
    # factor 1: bit-exact
    got = caffemodel.load_caffemodel(caffemodel.save_caffemodel(net, w), net)
    assert all(np.array_equal(got[n][f], w[n][f]) for n in w for f in w[n])


def test_bn_zero_factor_gives_zero_statistics():
    net = _mini_net()
    w = weights.random_weights(net, 7)
    buf = caffemodel.save_caffemodel(net, w, bn_factor=0.0)
    got = caffemodel.load_caffemodel(buf, net)
    assert not got["bn1"]["mean"].any() and not got["bn1"]["var"].any()


def test_reader_sees_layers_in_file_order_and_legacy_dims():
    net = _mini_net()
    w = weights.random_weights(net, 1)
    layers = caffemodel.read_layers(caffemodel.save_caffemodel(net, w))
    assert [L["name"] for L in layers] == [L["name"] for L in net["layers"]]
    assert layers[0]["type"] == "Convolution" and layers[0]["blobs"][0].shape == (64, 3, 3, 3)
    # a pre-BlobShape blob: num/channels/height/width = fields 1..4, data = field 5
    data = np.arange(2 * 3 * 1 * 1, dtype="<f4")
    legacy = b"".join(caffemodel._vi((f << 3) | 0) + caffemodel._vi(v) for f, v in ((1, 2), (2, 3), (3, 1), (4, 1))) + caffemodel._ld(5, data.tobytes())
    assert caffemodel._blob(legacy).shape == (2, 3, 1, 1)


def test_shape_mismatch_and_missing_layer_are_errors():
    net = _mini_net()
    w = weights.random_weights(net, 1)
    buf = caffemodel.save_caffemodel(net, w)
    other = _mini_net()
    other["layers"][0]["num_output"] = 32
    with pytest.raises(ValueError, match="num_output"):
        caffemodel.load_caffemodel(buf, other)
    other = _mini_net()
    other["layers"][5]["name"] = "fc1000"
    with pytest.raises(ValueError, match="no layer"):
        caffemodel.load_caffemodel(buf, other)
    v1 = caffemodel._ld(2, b"\x0a\x01x")
    with pytest.raises(ValueError, match="V1"):
        caffemodel.read_layers(v1)


def test_plan_from_caffemodel_is_identical(tmp_path):
    """tools/build_engine.py --prototxt ... --caffemodel ...  ==  plan built from the in-memory weights."""
    net = graph.resnet_caffe(50)
    w = weights.random_weights(net, 5)
    mp = tmp_path / "rn50.caffemodel"
    mp.write_bytes(caffemodel.save_caffemodel(net, w))
    out = tmp_path / "rn50.plan"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "build_engine.py"), "--model", "resnet50", "--caffemodel", str(mp),
                        "--precision", "fp16", "--batch", "2", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = builder.build_plan(graph.lower(net, w), builder.PREC_FP16, 2)
    assert out.read_bytes() == want


def test_corrupt_files_are_value_errors_never_crashes():
    """Random truncations and byte flips of a valid file either load (payload bytes changed) or raise ValueError."""
    import random
    net = _mini_net()
    buf = caffemodel.save_caffemodel(net, weights.random_weights(net, 1))
    rnd = random.Random(0)
    outcomes = {"ok": 0, "ValueError": 0}
    for t in range(400):
        b = bytearray(buf)
        if t % 2:
            b = b[:rnd.randrange(1, len(b))]
        else:
            for _ in range(3):
                b[rnd.randrange(len(b))] = rnd.randrange(256)
        try:
            caffemodel.load_caffemodel(bytes(b), net)
            outcomes["ok"] += 1
        except ValueError:
            outcomes["ValueError"] += 1
    assert outcomes["ValueError"] > 100 and sum(outcomes.values()) == 400


def test_build_tool_int8_from_caffemodel(tmp_path):
    """tools/build_engine.py --caffemodel ... --precision int8: the plan it writes is the INT8 plan of the same weights."""
    net = graph.resnet_caffe(50)
    w = weights.random_weights(net, 5)
    mp = tmp_path / "rn50.caffemodel"
    mp.write_bytes(caffemodel.save_caffemodel(net, w))
    out = tmp_path / "rn50_i8.plan"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "build_engine.py"), "--model", "resnet50", "--caffemodel", str(mp),
                        "--precision", "int8", "--batch", "2", "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from tensorrt_laboratory_b200 import quantize
    low = quantize.quantize_lowered(graph.lower(net, w), weights.synthetic_input(8, seed=4321))
    assert out.read_bytes() == builder.build_plan(low, builder.PREC_INT8, 2)
