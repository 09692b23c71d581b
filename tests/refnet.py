"""Load the reference's OWN ``models/networks.py`` (staged under oracle/_ref by oracle/build_ref.py) with this repository's
three modules switched in for ``.IPSR_model``, ``.InnerCos`` and ``.InnerCos2`` (models/networks.py:11-13) -- the
three-import switch of INTEGRATION.md, done without editing a file: the reference's source is executed as the module
``ipsr_refnet.networks`` inside a synthetic package whose sibling modules are ours."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def staged_networks_path():
    from oracle import ref_runner
    root = ref_runner.reference_dir()
    return None if root is None else os.path.join(root, "models", "networks.py")


def load_networks_with_dropin():
    """Returns the reference's networks module wired to deepinpainting_b200's IPSR_model / InnerCos / InnerCos2."""
    if "ipsr_refnet.networks" in sys.modules:
        return sys.modules["ipsr_refnet.networks"]
    path = staged_networks_path()
    if path is None:
        raise FileNotFoundError("oracle/_ref is not staged (python oracle/build_ref.py where /root/reference exists)")
    import deepinpainting_b200.models  # noqa: F401  (the package re-exports the classes under the submodules' names)
    mod_shift = sys.modules["deepinpainting_b200.models.IPSR_model"]
    mod_cos = sys.modules["deepinpainting_b200.models.InnerCos"]
    mod_cos2 = sys.modules["deepinpainting_b200.models.InnerCos2"]
    pkg = types.ModuleType("ipsr_refnet")
    pkg.__path__ = []                                       # a package without a directory: siblings come from sys.modules
    sys.modules["ipsr_refnet"] = pkg
    sys.modules["ipsr_refnet.IPSR_model"] = mod_shift
    sys.modules["ipsr_refnet.InnerCos"] = mod_cos
    sys.modules["ipsr_refnet.InnerCos2"] = mod_cos2
    spec = importlib.util.spec_from_file_location("ipsr_refnet.networks", path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ipsr_refnet.networks"] = mod
    spec.loader.exec_module(mod)                            # the reference's file, unmodified
    return mod


class Opt:
    """The hot-path options of the reference's option object (app.py:1-60)."""
    threshold = 5 / 16.0
    fixed_mask = 1
    shift_sz = 1
    stride = 1
    mask_thred = 1
    triple_weight = 1
    strength = 1
    skip = 0
