"""`alias_free_activation.torch`: the reference's torch implementation, aliased -- not re-implemented.

bigvgan.py:19 imports `alias_free_activation.torch.act.Activation1d` unconditionally, and the
reference's own act.py / resample.py import `alias_free_activation.torch.{resample,filter}`, yet the
reference ships those files flat in BigVGAN/alias_free_activation/.  This package finds that
directory (env AFA_REFERENCE_AFA_DIR, or any later sys.path entry holding an
`alias_free_activation/act.py`) and loads the three files unmodified as our sub-modules.
Without the reference tree importing a sub-module raises ImportError: this repository ships no
torch/CPU implementation of the path.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_reference_dir():
    cand = os.environ.get("AFA_REFERENCE_AFA_DIR")
    if cand and os.path.isfile(os.path.join(cand, "act.py")):
        return cand
    for p in sys.path:
        d = os.path.join(p or ".", "alias_free_activation")
        if os.path.isfile(os.path.join(d, "act.py")) and os.path.abspath(d) != os.path.dirname(_HERE):
            return d
    return None


def _alias(name: str, ref_dir: str):
    full = f"{__name__}.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(ref_dir, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


_ref = _find_reference_dir()
if _ref is not None:
    filter = _alias("filter", _ref)        # noqa: A001  (order matters: resample imports filter, act imports resample)
    resample = _alias("resample", _ref)
    act = _alias("act", _ref)
