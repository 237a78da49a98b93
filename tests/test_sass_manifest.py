"""Device code that has passed `pytest -m gpu` on a B200 must not change between two GPU runs without notice.

profiles/validated_sass.json holds the SHA-256 of the SASS instruction stream of every GPU-validated kernel
(tools/sass_manifest.py).  This container has no GPU, so an edit that silently changes the code of a validated kernel
(a shared helper, a compiler flag) could only be caught at the next GPU run; this test catches it at build time.  After
an INTENDED change, re-validate on the GPU and refresh the manifest with `python tools/sass_manifest.py --write <kernel>`.
Kernels that never ran on a GPU (frame_window) are deliberately not listed."""
import importlib.util
import json
import os
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MANIFEST = os.path.join(ROOT, "profiles", "validated_sass.json")

have_tools = (shutil.which("cuobjdump") or os.path.exists("/usr/local/cuda/bin/cuobjdump")) and \
             (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))


@pytest.mark.skipif(not have_tools, reason="needs nvcc and cuobjdump")
def test_validated_kernels_are_unchanged():
    spec = importlib.util.spec_from_file_location("sass_manifest", os.path.join(ROOT, "tools", "sass_manifest.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    from fdtd_em_solver_b200 import _lib
    want = json.load(open(MANIFEST))
    assert want["nvcc_flags"] == _lib.NVCC_FLAGS
    got = tool.kernel_hashes(_lib.build())
    changed = [v["kernel"] for k, v in want["kernels"].items() if got.get(k) != v["sha256"]]
    assert not changed, "device code changed since its last GPU validation: %s" % changed
    for family in ("step_stream", "stepK_stream", "step_exact", "step_resident"):
        assert any(family in v["kernel"] for v in want["kernels"].values())
    assert not any("frame_window" in v["kernel"] for v in want["kernels"].values())
