"""Drop-in for the reference's ``plankassembly/models.py`` (ref: trainer_complete.py:15 `from plankassembly.models import
build_model`).

Put this directory (``shim/``) in front of the reference checkout on ``sys.path`` / ``PYTHONPATH``:

    PYTHONPATH=/path/to/plankassembly_b200_repo/shim:/path/to/plankassembly_b200_repo:/path/to/PlankAssembly \\
        python trainer_complete.py fit --config configs/train_complete.yaml

``plankassembly`` is a namespace package in the reference (it ships no ``__init__.py``) and this directory has none
either, so ``plankassembly.datasets`` / ``plankassembly.metric`` still resolve to the reference's files while
``plankassembly.models`` resolves here: the B200 kernels behind the same ``build_model`` / ``PlankModel`` surface.
"""
from plankassembly_b200.models import PlankModel, build_model  # noqa: F401

__all__ = ['PlankModel', 'build_model']
