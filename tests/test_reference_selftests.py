"""Run the reference's OWN unit tests (reference tests/test_gantts.py, 5 tests) on top of the
nnmnkwii restatement, in the build container where /root/reference exists.  This anchors
oracle/nnmnkwii_port.py on the only tests the reference holds for the path.
Skipped on the GPU box (no reference tree there)."""
import importlib.util
import os
import sys

import pytest

from oracle import reference_loader


@pytest.mark.skipif(not reference_loader.available(), reason="reference tree not present")
def test_reference_unit_tests_pass_on_the_port():
    ref = reference_loader.load()
    stubs = reference_loader._stub_modules()
    names = list(stubs) + ["gantts", "gantts.models", "gantts.seqloss", "gantts.multistream"]
    saved = {k: sys.modules.get(k) for k in names}
    try:
        sys.modules.update(stubs)
        import types
        pkg = types.ModuleType("gantts")
        pkg.models, pkg.seqloss, pkg.multistream = ref.models, ref.seqloss, ref.multistream
        sys.modules.update({"gantts": pkg, "gantts.models": ref.models,
                            "gantts.seqloss": ref.seqloss, "gantts.multistream": ref.multistream})
        path = os.path.join(reference_loader.REFERENCE_ROOT, "tests", "test_gantts.py")
        spec = importlib.util.spec_from_file_location("_ref_test_gantts", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ran = 0
        for name in sorted(dir(mod)):
            if name.startswith("test_"):
                getattr(mod, name)()
                ran += 1
        assert ran == 5
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
