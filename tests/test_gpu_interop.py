"""GPU: zero-copy device interop of the Python surface.  The
reference's binding only moves host numpy arrays (llo/python/llo.cpp:36-147); here results
leave as __cuda_array_interface__ / DLPack without a copy and leaves accept CUDA tensors of
other libraries without a host bounce.  torch is the other library (plumbing, not product)."""
import numpy as np
import pytest

import util
from util import age, llo, oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

F4 = np.dtype(np.float32)


def test_result_is_visible_to_torch_without_a_copy():
    x = np.random.default_rng(41).uniform(0.5, 2, (33, 65)).astype(np.float32)
    root = age.exp(llo.variable(x, "x"))
    res = llo.evaluate_device(root, F4)
    cai = res.__cuda_array_interface__
    assert cai["shape"] == (33, 65) and cai["typestr"] == "<f4" and cai["data"][0] == res.ptr()
    t = torch.as_tensor(res, device="cuda")
    assert t.data_ptr() == res.ptr() and tuple(t.shape) == (33, 65)
    np.testing.assert_array_equal(t.cpu().numpy(), res.numpy(F4))
    d = torch.from_dlpack(res)
    assert d.data_ptr() == res.ptr() and d.dtype == torch.float32
    want = oracle.evaluate(root, F4)
    assert np.abs(d.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()
    # the DLPack tensor keeps the block alive after the result handle is gone
    del res, t
    torch.cuda.synchronize()
    np.testing.assert_array_equal(d.cpu().numpy().shape, (33, 65))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.int32, torch.int64])
def test_leaves_accept_cuda_tensors(dtype):
    g = torch.Generator(device="cuda").manual_seed(5)
    t = (torch.rand((17, 40), generator=g, device="cuda") * 8 + 1).to(dtype)
    host = t.cpu().numpy()
    v = llo.variable(t, "t")                      # device -> device, no host bounce
    assert list(v.shape()) == [17, 40]
    np_dtype = np.dtype(host.dtype)
    np.testing.assert_array_equal(llo.evaluate(v, np_dtype), host)
    root = age.add(v, v)
    np.testing.assert_array_equal(llo.evaluate(root, np_dtype), host + host)
    # assign from another CUDA tensor, produced on torch's stream just before
    t2 = t * 3
    v.assign(t2)
    np.testing.assert_array_equal(llo.evaluate(root, np_dtype), (t2 + t2).cpu().numpy())
    # the host mirror is refreshed on demand (pbm / serialize read Variable::data())
    np.testing.assert_array_equal(llo.evaluate(v, np_dtype), t2.cpu().numpy())


def test_device_assign_checks_shape_and_dtype_like_the_host_path():
    v = llo.variable(np.zeros((4, 5), dtype=np.float32), "v")
    with pytest.raises(RuntimeError, match="incompatible shaped"):
        v.assign(torch.zeros((5, 4), device="cuda"))
    with pytest.raises(RuntimeError, match="incompatible types"):
        v.assign(torch.zeros((4, 5), device="cuda", dtype=torch.float64))
    with pytest.raises(RuntimeError, match="C-contiguous"):
        v.assign(torch.zeros((5, 4), device="cuda").t())
