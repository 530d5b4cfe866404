import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _ensure_built():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "_cortenn_b200_build", os.path.join(ROOT, "cortenn_b200", "build.py"))
    product_build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(product_build)
    if not (os.path.exists(product_build.cuda_lib_path()) and
            os.path.exists(product_build.host_lib_path())):
        product_build.build_all()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_oracle
    if not (os.path.exists(build_oracle.capi_path()) and
            os.path.exists(build_oracle.module_path())):
        build_oracle.build()


_ensure_built()


def gpu_available():
    try:
        import cortenn_b200
        return bool(cortenn_b200.llo.available())
    except Exception:
        return False


@pytest.fixture(scope="session")
def ctx():
    """C-ABI context on cuda:0 (GPU tests only)"""
    from cortenn_b200 import capi
    context = capi.Context(int(os.environ.get("LOCAL_RANK", "0")))
    yield context
    context.close()
