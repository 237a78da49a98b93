"""Device code that has passed `pytest -m gpu` on a B200 must not change between two GPU runs without notice.

profiles/validated_sass.json holds the SHA-256 of the SASS instruction stream of every GPU-validated kernel
(tools/sass_manifest.py).  This container has no GPU, so an edit that silently changes the code of a validated kernel
(a shared helper, a compiler flag) could only be caught at the next GPU run; this test catches it at build time.  After
an INTENDED change, re-validate on the GPU and refresh the manifest with `python tools/sass_manifest.py --write <kernel>`.
Kernels that have not run on a GPU yet are deliberately not listed."""
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


@pytest.mark.skipif(not have_tools, reason="needs nvcc and cuobjdump")
def test_no_update_kernel_contains_a_fused_multiply_add():
    """Bit-exactness against numpy rests on every a + b*c of the update being TWO roundings.  The sources say so
    (__dadd_rn / __dmul_rn, -fmad=false), but what runs is SASS: no kernel that advances fields -- validated or still
    experimental -- may contain an FMA of any width.  (The coefficient builders are exempt: their DFMA/FFMA are the
    Newton steps of the correctly rounded IEEE division, pinned against the oracle's division by the GPU tests.)"""
    import re
    import subprocess
    from fdtd_em_solver_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    text = subprocess.run([cuobjdump, "-sass", _lib.build()], capture_output=True, text=True, check=True).stdout
    fused, cur, kernels = {}, None, set()
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels.add(cur)
            continue
        # (HFMA2 is not scanned: "HFMA2 R, -RZ, RZ, 0, 0" is the compiler's register-zeroing idiom; nothing here is fp16)
        if cur and re.search(r"\b(DFMA|FFMA2?)\b", line):
            fused[cur] = fused.get(cur, 0) + 1
    assert len(kernels) > 20
    offenders = [k for k in fused if "synth_coeffs" not in k and "paint_coeffs" not in k]
    assert not offenders, offenders
    assert any("synth_coeffs" in k for k in fused)            # the scan does see FMAs where they are allowed
