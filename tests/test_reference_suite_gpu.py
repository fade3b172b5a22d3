"""The reference's OWN unit tests for the hot path, run against this package.

The compiled reference in ``baseline/_ref`` (installed from /root/reference by ``__graft_entry__.build()``; it
travels to the GPU box) ships its test modules.  This harness imports them with the name ``pypmc`` resolved to
``pypmc_b200`` -- ``from pypmc.density.gauss import *`` inside ``gauss_test.py`` then binds THIS package's classes --
and runs their ``unittest`` cases unmodified: golden numbers, hand-computed update tables, error contracts and all.
Only the reference's pure-Python test helpers (``tools/_probability_densities.py``, the ``*_test.py`` modules
themselves) are loaded from the reference tree; none of its compiled hot-path modules is imported.

Cases outside the path (VBMerge, plotting, ...) are listed in ``OUT_OF_SCOPE`` with the reason; every
other case must pass.
"""
import importlib
import importlib.abc
import importlib.util
import os
import sys
import unittest

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

REF_PKG = os.path.join(ROOT, "baseline", "_ref", "pypmc")

#: reference modules that may be loaded from the reference tree itself: pure-Python helpers of its tests
PURE_PYTHON_HELPERS = {"pypmc.tools._probability_densities"}

#: test modules of the reference that exercise the hot path
MODULES = ["pypmc.density.base_test", "pypmc.density.gauss_test", "pypmc.density.student_t_test",
           "pypmc.density.mixture_test", "pypmc.mix_adapt.pmc_test", "pypmc.mix_adapt.variational_test",
           "pypmc.sampler.importance_sampling_test", "pypmc.tools.convergence_test", "pypmc.tools.regularize_test",
           "pypmc.tools.linalg_test"]

#: (class or class.method) -> why it is not expected to pass here
OUT_OF_SCOPE = {
    "TestVBMerge": "VBMerge (variational.pyx:1035-1218) loops over components, not samples: out of scope (SURVEY 2)",
}


class _Alias(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """``pypmc[.x.y]`` -> ``pypmc_b200[.x.y]``; ``*_test`` modules and the whitelisted helpers from baseline/_ref."""

    def find_spec(self, name, path=None, target=None):
        if name != "pypmc" and not name.startswith("pypmc."):
            return None
        return importlib.util.spec_from_loader(name, self, is_package=True)

    def create_module(self, spec):
        name = spec.name
        if name.endswith("_test") or name in PURE_PYTHON_HELPERS:
            src = os.path.join(REF_PKG, *name.split(".")[1:]) + ".py"
            mod = type(sys)(name)
            mod.__file__ = src
            mod.__package__ = name.rpartition(".")[0]
            return mod
        return importlib.import_module("pypmc_b200" + name[len("pypmc"):])      # ImportError if we do not have it

    def exec_module(self, module):
        src = getattr(module, "__file__", "")
        if src.startswith(REF_PKG) and (module.__name__.endswith("_test") or module.__name__ in PURE_PYTHON_HELPERS):
            with open(src) as fh:
                exec(compile(fh.read(), src, "exec"), module.__dict__)


@pytest.fixture(scope="module")
def reference_tests():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.isdir(REF_PKG):
        pytest.skip("baseline/_ref is not installed (run __graft_entry__.build() where /root/reference exists)")
    saved = {k: v for k, v in sys.modules.items() if k == "pypmc" or k.startswith("pypmc.")}
    for k in saved:
        del sys.modules[k]
    finder = _Alias()
    sys.meta_path.insert(0, finder)
    try:
        yield
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k in sys.modules if k == "pypmc" or k.startswith("pypmc.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def _cases(suite):
    for item in suite:
        if isinstance(item, unittest.TestSuite):
            yield from _cases(item)
        else:
            yield item


@pytest.mark.parametrize("modname", MODULES)
def test_reference_test_module(reference_tests, modname):
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = importlib.import_module(modname)
        # the classes this module's tests exercise must be OURS
        for attr in ("Gauss", "StudentT", "MixtureDensity", "gaussian_pmc", "GaussianInference", "ImportanceSampler"):
            obj = getattr(mod, attr, None)
            if obj is not None:
                assert obj.__module__.startswith("pypmc_b200."), (attr, obj.__module__)
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        selected, skipped = unittest.TestSuite(), []
        for case in _cases(suite):
            cls, meth = type(case).__name__, case._testMethodName
            reason = OUT_OF_SCOPE.get(cls + "." + meth, OUT_OF_SCOPE.get(cls))
            if reason:
                skipped.append((cls + "." + meth, reason))
            else:
                selected.addTest(case)
        result = unittest.TestResult()
        selected.run(result)
    problems = ["%s:\n%s" % (case.id(), "\n".join(tb.strip().splitlines()[-25:])) for case, tb in result.failures + result.errors]
    print("%s: ran %d of the reference's cases, %d failed, %d out of scope" % (modname, result.testsRun, len(problems), len(skipped)))
    assert result.testsRun > 0
    assert not problems, "\n".join(problems)
