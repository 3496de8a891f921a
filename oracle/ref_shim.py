"""TEST INFRASTRUCTURE ONLY -- container-side shim that makes the *real* reference importable.

`/root/reference` exists only in the authoring container (never on the GPU box), so this module
is used solely by `oracle/make_golden.py` to generate the committed fixtures under
`tests/golden/`.  Nothing in the product package, `bench.py`, `smoke()` or the `-m gpu` tests
imports it.

What it does:
  * exposes `/root/reference` as the package `InverseProblemWithDiffusionModel` through a
    symlink in a temp dir on `sys.path` (every intra-repo import of the reference is absolute);
  * inserts empty stand-ins for `matplotlib`, `matplotlib.pyplot` and `SimpleITK`, which
    `helpers/utils.py:2,7` imports but the hot path never calls;
  * turns the plotting helpers the samplers call unconditionally into no-ops
    (`ALD_optimizers.py:203,307,416,556-581`).
"""
import os
import sys
import tempfile
import types

REFERENCE_ROOT = "/root/reference"
PKG = "InverseProblemWithDiffusionModel"


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "ncsn"))


def install():
    """Make `import InverseProblemWithDiffusionModel...` work; idempotent."""
    if not available():
        raise RuntimeError("reference tree not present (this shim only works in the authoring container)")
    if PKG in sys.modules:
        return
    for name in ("matplotlib", "matplotlib.pyplot", "SimpleITK"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    alias_dir = tempfile.mkdtemp(prefix="ipdm_ref_alias_")
    os.symlink(REFERENCE_ROOT, os.path.join(alias_dir, PKG))
    sys.path.insert(0, alias_dir)
    import importlib

    ald = importlib.import_module(PKG + ".ncsn.models.ALD_optimizers")
    ald.vis_images = lambda *a, **k: None
    ald.vis_multi_channel_signal = lambda *a, **k: None


def load_config(name: str):
    """`ncsn/configs/<name>.yml` as the nested Namespace the reference uses (helpers/utils.py:173-191)."""
    install()
    import importlib

    utils = importlib.import_module(PKG + ".helpers.utils")
    cfg = utils.load_yml_file(os.path.join(REFERENCE_ROOT, "ncsn", "configs", name + ".yml"))
    import torch

    cfg.device = torch.device("cpu")
    return cfg
