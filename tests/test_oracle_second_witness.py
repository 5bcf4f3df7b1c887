"""Two witnesses for the ops the reference's MNIST golden vectors do not pin.

``oracle/numpy_ops.py`` (plain numpy, float64, loops over taps / windows) and ``oracle/caffe_forward.py`` (torch) are
independent restatements of the same Caffe semantics; they must agree on BatchNorm+Scale, their folding into the
convolution, ceil-mode MAX pooling with clipped windows, global AVE pooling, Eltwise SUM and Softmax -- op by op and on a
whole bottleneck network shaped like the reference's ResNet deploy files."""
import numpy as np
import pytest
import torch

from oracle import numpy_ops as N
from oracle.caffe_forward import _pool_out, caffe_forward, lowered_forward_f16emu
from tensorrt_laboratory_b200 import graph, weights


def _mini_resnet(c_in=3, hw=29, width=8, classes=10):
    """conv 3x3/2 +BN+Scale+ReLU, max pool 3x3/2 (ceil mode with a clipped last window), one bottleneck with a projection
    shortcut (Eltwise SUM + ReLU), global AVE pool, InnerProduct, Softmax -- the ResNet-50 deploy vocabulary in small."""
    L = []

    def conv(name, bottom, cout, k, s, p, bias):
        L.append(dict(name=name, type="Convolution", bottoms=[bottom], tops=[name], num_output=cout, kernel_size=k, pad=p, stride=s,
                      bias_term=bias))
        L.append(dict(name="bn" + name, type="BatchNorm", bottoms=[name], tops=[name], use_global_stats=True, eps=1e-5))
        L.append(dict(name="scale" + name, type="Scale", bottoms=[name], tops=[name], bias_term=True))

    def relu(name, blob):
        L.append(dict(name=name, type="ReLU", bottoms=[blob], tops=[blob]))

    conv("conv1", "data", width, 3, 2, 1, True)
    relu("conv1_relu", "conv1")
    L.append(dict(name="pool1", type="Pooling", bottoms=["conv1"], tops=["pool1"], pool="MAX", kernel_size=3, stride=2, pad=0))
    conv("res2a_branch1", "pool1", 4 * width, 1, 1, 0, False)
    conv("res2a_branch2a", "pool1", width, 1, 1, 0, False)
    relu("res2a_branch2a_relu", "res2a_branch2a")
    conv("res2a_branch2b", "res2a_branch2a", width, 3, 1, 1, False)
    relu("res2a_branch2b_relu", "res2a_branch2b")
    conv("res2a_branch2c", "res2a_branch2b", 4 * width, 1, 1, 0, False)
    L.append(dict(name="res2a", type="Eltwise", bottoms=["res2a_branch1", "res2a_branch2c"], tops=["res2a"], operation="SUM"))
    relu("res2a_relu", "res2a")
    h1 = (hw + 2 - 3) // 2 + 1
    hp = _pool_out(h1, 3, 0, 2, True)
    L.append(dict(name="pool5", type="Pooling", bottoms=["res2a"], tops=["pool5"], pool="AVE", kernel_size=hp, stride=1, pad=0))
    L.append(dict(name="fc", type="InnerProduct", bottoms=["pool5"], tops=["fc"], num_output=classes, bias_term=True))
    L.append(dict(name="prob", type="Softmax", bottoms=["fc"], tops=["prob"]))
    return dict(name="mini", input="data", input_dims=[1, c_in, hw, hw], layers=L)


def test_whole_network_two_oracles_agree():
    net = _mini_resnet()
    w = weights.random_weights(net, 5)
    x = np.random.default_rng(1).standard_normal((3, 3, 29, 29)).astype(np.float32)
    a = caffe_forward(net, w, x, dtype=torch.float64)
    b = N.forward(net, w, x)
    np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-12)
    # the fused / folded graph the ENGINE executes (graph.lower: BN+Scale folded, ReLU / Eltwise fused) evaluates to the
    # same function -- checked against the numpy oracle's UNFUSED evaluation
    fused = lowered_forward_f16emu(graph.lower(net, w), x, round16=False)
    np.testing.assert_allclose(fused, b, rtol=1e-5, atol=1e-7)


def test_bn_scale_fold_matches_the_unfused_chain():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 5, 9, 9))
    W = rng.standard_normal((7, 5, 3, 3)) * 0.2
    b = rng.standard_normal(7) * 0.1
    mean, var = rng.standard_normal(7) * 0.1, rng.uniform(0.5, 1.5, 7)
    gamma, beta = rng.uniform(0.8, 1.2, 7), rng.standard_normal(7) * 0.1
    chain = N.scale(N.batchnorm(N.conv2d(x, W, b, 1, 1), mean, var), gamma, beta)
    Wf, bf = N.fold_bn_scale(W, b, mean, var, gamma, beta)
    np.testing.assert_allclose(N.conv2d(x, Wf, bf, 1, 1), chain, rtol=1e-12, atol=1e-12)
    t = torch.from_numpy
    ref = torch.nn.functional.conv2d(t(x), t(W), t(b), stride=1, padding=1)
    ref = (ref - t(mean).view(1, -1, 1, 1)) / torch.sqrt(t(var).view(1, -1, 1, 1) + 1e-5) * t(gamma).view(1, -1, 1, 1) + t(beta).view(1, -1, 1, 1)
    np.testing.assert_allclose(chain, ref.numpy(), rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("size,k,s,p", [(112, 3, 2, 0), (15, 3, 2, 0), (14, 3, 2, 1), (7, 2, 2, 0), (8, 3, 3, 1)])
def test_ceil_mode_max_pool(size, k, s, p):
    """112 -> 56 is pool1 of the reference's ResNets: ceil((112-3)/2)+1 = 56, the last window is clipped to 2 pixels."""
    x = np.random.default_rng(size).standard_normal((2, 3, size, size))
    net = dict(name="p", input="data", input_dims=[1, 3, size, size],
               layers=[dict(name="pool", type="Pooling", bottoms=["data"], tops=["pool"], pool="MAX", kernel_size=k, stride=s, pad=p)])
    a = caffe_forward(net, {}, x.astype(np.float32), dtype=torch.float64)
    b = N.maxpool(x.astype(np.float32), k, s, p)
    assert b.shape[2] == N.pool_out_size(size, k, p, s) == _pool_out(size, k, p, s, True)
    np.testing.assert_array_equal(a.reshape(b.shape), b)
    if (size, k, s, p) == (112, 3, 2, 0):
        assert b.shape[2] == 56
        np.testing.assert_array_equal(b[:, :, 55, 55], x.astype(np.float32)[:, :, 110:112, 110:112].max(axis=(2, 3)))


def test_avgpool_eltwise_softmax():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((4, 6, 7, 7)).astype(np.float32)
    y = rng.standard_normal((4, 6, 7, 7)).astype(np.float32)
    t = torch.from_numpy
    np.testing.assert_allclose(N.avgpool_global(x), torch.nn.functional.avg_pool2d(t(x).double(), 7).numpy(), rtol=1e-12, atol=1e-14)
    np.testing.assert_array_equal(N.eltwise_sum(x, y), (t(x).double() + t(y).double()).numpy())
    z = rng.standard_normal((4, 1000, 1, 1)) * 30
    np.testing.assert_allclose(N.softmax(z), torch.softmax(t(z), dim=1).numpy(), rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(N.softmax(z).reshape(4, -1).sum(1), 1.0, rtol=1e-12)
