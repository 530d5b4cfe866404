"""CPU: how the planner classifies graphs (llo.plan_preview needs no GPU).

The lowering decisions that matter for the hot path -- every argument's
ade::CoordMap folded into an integer coordinate view, nothing left on the
per-element generic-map kernels -- are host logic and are checked here; the
numerics of the same graphs are checked on the GPU (tests/test_gpu_graph.py)."""
import numpy as np
import workloads
from cortenn_b200 import age, llo

F4 = np.dtype(np.float32)
F8 = np.dtype(np.float64)


def var(shape, label, dtype=np.float32):
    rng = np.random.default_rng(len(label))
    return llo.variable(rng.uniform(0.5, 2.0, shape).astype(dtype), label)


def test_matmul_and_its_gradients_have_no_generic_arguments():
    # llo::matmul (llo/src/helper.cpp:67-97) and grad_prod (:18-33)
    x, w = var((96, 64), "x"), var((64, 32), "w")
    m = age.matmul(x, w)
    st = llo.plan_preview([m], F4)
    assert st["generic"] == 0 and st["views"] >= 2
    loss = age.reduce_sum0(age.pow(m, llo.scalar(2.0, [96, 32])))
    st = llo.plan_preview([llo.derive(loss, w), llo.derive(loss, x)], F4)
    assert st["generic"] == 0
    assert st["shared"] > 0          # the two gradients share the forward pass


def test_convolution_and_its_kernel_gradient_have_no_generic_arguments():
    # llo::convolution (llo/src/helper.cpp:102-202): two-source integer coordinate maps
    img, ker = var((2, 8, 8, 3), "img", np.float64), var((3, 3, 3, 4), "ker", np.float64)
    root = age.convolution(img, ker)
    assert llo.plan_preview([root], F8)["generic"] == 0
    assert llo.plan_preview([llo.derive(root, ker)], F8)["generic"] == 0


def test_scatter_maps_stay_on_the_generic_path():
    # the image gradient of a convolution pushes through a two-source map (col2im): it is not
    # a view of the output coordinates, and the planner must say so instead of mis-addressing
    # (DESIGN.md section 6 lists its lowering as open)
    img, ker = var((2, 8, 8, 3), "img", np.float64), var((3, 3, 3, 4), "ker", np.float64)
    root = age.convolution(img, ker)
    assert llo.plan_preview([llo.derive(root, img)], F8)["generic"] >= 1


def test_mlp_gradient_graph_is_fully_lowered_and_deterministic():
    # BASELINE config 5 (down-scaled batch): nothing generic, and the classification of the
    # same graph is the same on every call (the CUDA-graph replay keys on it)
    X, Y, params = workloads.mlp_data(64, sizes=(16, 32, 32, 10))
    loss, pvars, _ = workloads.mlp_graph(age, llo, X, Y, params, global_batch=64)
    grads = [llo.derive(loss, p) for p in pvars]
    roots, report = llo.optimize([loss] + grads, llo.OPT_ALL)
    assert report["div_cancel"] > 0      # what lets the matmul gradients lower to GEMMs
    first = llo.plan_preview(list(roots), F4)
    assert first["generic"] == 0
    assert first["shared"] > 0
    assert llo.plan_preview(list(roots), F4) == first
