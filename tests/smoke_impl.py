"""__graft_entry__.smoke(): one small invocation of the hot path on cuda:0, checked against the oracle."""
import numpy as np


def run():
    from oracle.caffe_forward import lowered_forward_f16emu
    from tensorrt_laboratory_b200 import builder, capi, graph
    from tests import helpers

    if capi.device_count() < 1:
        raise RuntimeError("smoke(): no CUDA device (the product path has no CPU fallback)")
    info = capi.device_info(0)
    capi.check(capi.load().b2_device_set(0))
    # 1. the reference's known-answer model through the fp32 engine
    net, w, xs, ys = helpers.load_mnist_golden()
    low = graph.lower(net, w)
    got = list(helpers.run_engine(low, xs[0], builder.PREC_FP32).values())[0]
    assert np.abs(got - ys[0]).max() < 1.5e-3 and int(got.argmax()) == 2, "MNIST known-answer failed"
    # 2. one ResNet bottleneck-sized 3x3 convolution + fused residual/ReLU on the tcgen05 path
    _, _, low = helpers.conv_case(64, 56, 56, 64, 3, 1, 1, relu=True, residual=True)
    x = np.random.default_rng(0).standard_normal((2, 64, 56, 56), dtype=np.float32)
    ref = lowered_forward_f16emu(low, x)
    got = list(helpers.run_engine(low, x, builder.PREC_FP16).values())[0].reshape(2, -1)
    err = helpers.rel_err(got, ref)
    assert err <= 2.0 ** -9, f"tcgen05 conv parity failed: rel err {err:.3e}"
    print(f"smoke ok on {info['name']} (sm_{info['cc'][0]}{info['cc'][1]}): mnist argmax=2, conv rel err {err:.2e}")
