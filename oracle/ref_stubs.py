"""Import the reference's OWN hot-path modules in place (TEST-ONLY, build container only).

/root/reference cannot be imported as-is: cutouts.py:6 needs ``resize_right``,
losses.py:2 ``pytorch_msssim``, utils/functional.py:1,6,10,11 ``clip`` / ``anvil`` /
``tqdm.notebook`` / ``IPython``, utils/image_utils.py:3,5,16-22 ``pyimgur`` /
``firebase_admin``.  This module installs inert stand-ins for
those third-party imports (``resize_right.resize`` -> ``oracle.resize_right.resize``)
and puts /root/reference on ``sys.path`` so ``clip_diffusion.cutouts``, ``.losses``,
``.utils.functional`` and ``.config`` execute UNMODIFIED.  Nothing is copied.

/root/reference does not exist on the GPU box; ``available()`` says whether this
harness can be used, and only tests/golden generation depends on it.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CLIPGUIDE_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "clip_diffusion"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Inert()

    def __getattr__(self, name):
        return _Inert()


def install():
    """Returns (cutouts, losses, functional, config) reference modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from oracle import resize_right as _rr

    _stub("resize_right", resize=_rr.resize)
    _stub("pytorch_msssim", MS_SSIM=_Inert)
    _stub("clip", load=_Inert(), tokenize=_Inert())
    anvil = _stub("anvil", BlobMedia=_Inert)
    anvil.server = _stub("anvil.server", callable=lambda f=None, *a, **k: f, task_state={}, connect=_Inert())
    _stub("IPython", display=_Inert())
    import tqdm

    _stub("tqdm.notebook", trange=tqdm.trange, tqdm=tqdm.tqdm)
    _stub("pyimgur", Imgur=_Inert)
    fb = _stub("firebase_admin", initialize_app=_Inert(), _apps=[True])
    fb.credentials = _stub("firebase_admin.credentials", Certificate=_Inert)
    fb.storage = _stub("firebase_admin.storage", bucket=_Inert())
    _stub("cv2")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for k in [k for k in sys.modules if k == "clip_diffusion" or k.startswith("clip_diffusion.")]:
        del sys.modules[k]
    mods = []
    for name in ("clip_diffusion.cutouts", "clip_diffusion.losses", "clip_diffusion.utils.functional", "clip_diffusion.config"):
        mods.append(importlib.import_module(name))
    return tuple(mods)
