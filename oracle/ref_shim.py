"""TEST INFRASTRUCTURE ONLY — imports the unmodified reference (``/root/reference/code``).

This module exists so that ``tests/golden/make_golden.py`` (run in the build container, where
``/root/reference`` is mounted) can execute the reference itself and freeze its outputs into
``tests/golden/*.pt``.  Nothing in the product path, ``bench.py`` or the ``-m gpu`` tests imports it:
``/root/reference`` does not exist on the GPU box.

The reference does not import as-is under the installed ``transformers`` 5.x:
  * ``arguments.py:10``  ``from transformers.utils import cached_property``  -> gone
  * ``trainer.py:12-13`` ``from transformers import AdamW``                 -> removed
  * ``dataset.py:8``     ``import h5py``                                    -> not installed
so three names are patched in before the import.  ``AdamW`` is supplied by the restatement in
``oracle/map_oracle.py`` (HF 4.26.1 semantics).
"""
import functools
import os
import sys
import types

REFERENCE_CODE = "/root/reference/code"


def available() -> bool:
    return os.path.isdir(REFERENCE_CODE)


def import_reference():
    """Returns a namespace with the reference modules: layers, models, trainer, nce, arguments."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REFERENCE_CODE)
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    import transformers
    import transformers.utils

    from . import map_oracle

    if not hasattr(transformers.utils, "cached_property"):
        transformers.utils.cached_property = functools.cached_property
    transformers.AdamW = map_oracle.HFAdamW
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")

    # The reference uses top-level module names (layers, models, nce, ...) that collide with nothing in
    # this repo (our package is namespaced), so a plain sys.path insert is enough.
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    import arguments as ref_arguments
    import layers as ref_layers
    import models as ref_models
    import nce as ref_nce
    import trainer as ref_trainer

    ns = types.SimpleNamespace(
        arguments=ref_arguments, layers=ref_layers, models=ref_models, nce=ref_nce, trainer=ref_trainer
    )
    return ns
