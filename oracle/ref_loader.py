"""Load the UNMODIFIED reference modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing at
run time there may call this; it is used by ``oracle/make_golden.py`` and by the CPU
tests that (when the tree is present) compare the oracle against the live reference.

hstu.py / sasrec.py import only torch -> loaded by file path.  rqvae.py pulls ``gin``
and (through genrec/__init__) ``sentence_transformers``; neither is installed here, so
two inert stub modules are registered first.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("GENREC_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "genrec", "models", "hstu.py"))


def _by_path(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_hstu():
    return _by_path("_ref_genrec_hstu", "genrec/models/hstu.py")


def ref_sasrec():
    return _by_path("_ref_genrec_sasrec", "genrec/models/sasrec.py")


def _install_stubs():
    if "gin" not in sys.modules:
        gin = types.ModuleType("gin")

        def configurable(*a, **k):
            if len(a) == 1 and callable(a[0]) and not k:
                return a[0]
            return lambda f: f

        gin.configurable = configurable
        gin.constants_from_enum = lambda c=None, **k: c if c is not None else (lambda x: x)
        gin.parse_config = lambda *a, **k: None
        gin.REQUIRED = object()
        sys.modules["gin"] = gin
    if "sentence_transformers" not in sys.modules:
        st = types.ModuleType("sentence_transformers")
        st.SentenceTransformer = type("SentenceTransformer", (), {})
        sys.modules["sentence_transformers"] = st


def ref_genrec_package():
    """Import the whole reference ``genrec`` package under the name ``_refpkg`` is not possible
    (absolute imports), so it is imported as ``genrec`` from a temporary sys.path entry and
    immediately removed from sys.modules again (our own ``genrec`` shim has the same name)."""
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "genrec" or k.startswith("genrec.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        pkg = importlib.import_module("genrec")
        rq = importlib.import_module("genrec.models.rqvae")
        hstu_data = importlib.import_module("genrec.data.amazon_hstu")
        sas_data = importlib.import_module("genrec.data.amazon_sasrec")
        return types.SimpleNamespace(pkg=pkg, rqvae=rq, amazon_hstu=hstu_data, amazon_sasrec=sas_data)
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "genrec" or k.startswith("genrec.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def ref_transformer():
    """The reference's genrec.modules.transformer module (T5Attention, ...)."""
    return _ref_module("genrec.modules.transformer")


def ref_tiger():
    """The reference's genrec.models.tiger module (Tiger, build_trie), imported like ref_genrec_package()."""
    return _ref_module("genrec.models.tiger")


def _ref_module(name: str):
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "genrec" or k.startswith("genrec.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REF_ROOT)
    try:
        return importlib.import_module(name)
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if k == "genrec" or k.startswith("genrec.")]:
            del sys.modules[k]
        sys.modules.update(saved)
