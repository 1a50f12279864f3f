"""
TEST INFRASTRUCTURE - not part of the product path.

Makes the *unmodified* reference package importable in the build container so
its own code can pin the oracle.  The reference fails to
import only because of third-party packages that are absent here (plotoptix,
tkinter, skyfield, tzlocal); placeholder modules are injected into sys.modules
before `import moonrtx....`.

`/root/reference` does not exist on the GPU box, so nothing in `-m gpu` tests,
`smoke()` or `bench.py` may call this module; only `oracle/make_golden.py`
(run here, fixtures committed under tests/golden/) and the `not gpu` tests that
skip when the reference is absent use it.
"""

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MOONRTX_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "moonrtx", "data_loader.py"))


class _Placeholder:
    """Stands in for any class/function of a stubbed package."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Placeholder()

    def __getattr__(self, name):
        return _Placeholder()


def _module(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _any(attr):                 # any other symbol; dunders stay absent so `inspect` keeps working
        if attr.startswith("__") and attr.endswith("__"):
            raise AttributeError(attr)
        return _Placeholder

    m.__getattr__ = _any
    sys.modules[name] = m
    return m


# The array the stubbed plotoptix.utils.read_image hands back; set by callers.
_read_image_result = {}


def set_read_image(path: str, array) -> None:
    """Register what `read_image(path)` returns (a uint16 (H, W) array, which
    the reference reinterprets in place as int16, data_loader.py:215)."""
    _read_image_result[os.path.abspath(path)] = array


def _read_image(path, normalized=False):
    return _read_image_result.get(os.path.abspath(path))


def install_stubs() -> None:
    if "plotoptix" in sys.modules and getattr(sys.modules["plotoptix"], "_moonb200_stub", False):
        return
    p = _module("plotoptix", __version__="0.19.2", _moonb200_stub=True,
                TkOptiX=_Placeholder, NpOptiX=_Placeholder)
    p.__path__ = []
    _module("plotoptix.utils", read_image=_read_image,
            get_gpu_architecture=lambda *a, **k: None)
    _module("plotoptix.materials", m_diffuse={}, m_flat={})
    _module("plotoptix.enums", GpuArchitecture=_Placeholder)
    _module("plotoptix.install", download_file_from_google_drive=lambda *a, **k: None)
    if "tkinter" not in sys.modules:
        try:
            import tkinter  # noqa: F401
        except Exception:
            t = _module("tkinter", BOTH="both", X="x", Y="y", LEFT="left", RIGHT="right",
                        END="end", TOP="top", BOTTOM="bottom", W="w", E="e", N="n", S="s",
                        NORMAL="normal", DISABLED="disabled", WORD="word", NW="nw")
            t.__path__ = []
            for sub in ("font", "filedialog", "ttk", "messagebox", "simpledialog"):
                _module("tkinter." + sub)
    s = _module("skyfield")
    s.__path__ = []
    for sub in ("almanac", "api", "positionlib", "framelib", "trigonometry", "timelib",
                "units", "constants", "earthlib", "functions", "searchlib", "errors"):
        _module("skyfield." + sub)
    _module("tzlocal", get_localzone=lambda: None, get_localzone_name=lambda: "UTC")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def import_reference(name: str):
    """import_reference('data_loader') -> the reference's moonrtx.data_loader module."""
    if not reference_available():
        raise ImportError(f"reference tree not present at {REFERENCE_ROOT}")
    install_stubs()
    import importlib
    return importlib.import_module("moonrtx." + name)
