import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    import __graft_entry__ as g
    g.build_host()
    g.build_oracle()
    yield


@pytest.fixture(scope="session")
def hr(_built):
    import hanamaru_renderer_b200
    return hanamaru_renderer_b200


@pytest.fixture(scope="session")
def assets(hr):
    return hr.AssetStore.from_pack()


@pytest.fixture(scope="session")
def oracle(_built):
    from oracle_ffi import Oracle
    return Oracle("det")


@pytest.fixture(scope="session")
def oracle_glibc(_built):
    from oracle_ffi import Oracle
    return Oracle("glibc")


_scene_cache = {}


@pytest.fixture(scope="session")
def get_scene(hr, assets):
    def get(name):
        if name not in _scene_cache:
            _scene_cache[name] = hr.build_scene(name, assets)
        return _scene_cache[name]
    return get


@pytest.fixture(scope="session")
def core(hr):
    """The CUDA library + a device.  GPU tests fail (not skip) if either is missing."""
    import __graft_entry__ as g
    g.build_core()
    from hanamaru_renderer_b200 import _ffi
    _ffi.core()
    assert hr.device_count() >= 1, "GPU test without a CUDA device"
    return _ffi.core()


_dev_cache = {}


@pytest.fixture(scope="session")
def get_device_scene(hr, core, get_scene):
    def get(name):
        if name not in _dev_cache:
            _dev_cache[name] = hr.DeviceScene(get_scene(name), 0)
        return _dev_cache[name]
    return get
