"""TEST INFRASTRUCTURE ONLY — imports the reference's own scripts/utils.py (when /root/reference
is mounted, i.e. in the build container, never on the GPU box) with the two absent third-party
modules replaced by the CPU oracle:

    ripser.ripser          -> oracle.rips.ripser            
    persim.wasserstein     -> oracle.wasserstein_ref.wasserstein (Appendix A.2)

Everything else in utils.py (scipy / numpy code) then runs unmodified, so its functions can be
used to validate the restatements in oracle/ and to generate tests/golden/*.npz
(tests/golden/make_golden.py).  Nothing is copied: the module is executed from where it lies.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("TDA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "scripts", "utils.py"))


def load_utils():
    if not available():
        raise FileNotFoundError(REFERENCE_ROOT)
    from . import rips, wasserstein_ref
    saved = {k: sys.modules.get(k) for k in ("ripser", "persim")}
    m_r = types.ModuleType("ripser")
    m_r.ripser = rips.ripser
    m_p = types.ModuleType("persim")
    m_p.wasserstein = wasserstein_ref.wasserstein
    sys.modules["ripser"], sys.modules["persim"] = m_r, m_p
    try:
        spec = importlib.util.spec_from_file_location(
            "_reference_utils", os.path.join(REFERENCE_ROOT, "scripts", "utils.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
