import importlib.util
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver with -m gpu on the GPU box)")


def load_package():
    """import the hyphen-named product package literf-radiancefields_b200/ as module `lrf_b200`"""
    if "lrf_b200" in sys.modules:
        return sys.modules["lrf_b200"]
    spec = importlib.util.spec_from_file_location("lrf_b200", os.path.join(ROOT, "literf-radiancefields_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lrf_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def pkg():
    return load_package()


@pytest.fixture(scope="session")
def orc():
    import libs
    return libs.oracle()
