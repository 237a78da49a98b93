"""SHA-256 of the SASS instruction stream of every kernel in libfdtd2d.so (addresses and encodings stripped).

`python tools/sass_manifest.py` prints the manifest of the current build; `--write` refreshes
profiles/validated_sass.json, the list of kernels whose device code has passed `pytest -m gpu` on a B200.
tests/test_sass_manifest.py compares the build against it, so device code cannot change between two GPU runs
without somebody noticing (host-side edits and new kernels do not touch the hashes of the listed ones)."""
import hashlib, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MANIFEST = os.path.join(ROOT, "profiles", "validated_sass.json")


def kernel_hashes(lib_path):
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    text = subprocess.run([cuobjdump, "-sass", lib_path], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = hashlib.sha256()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if cur and m:
            out[cur].update((m.group(1).strip() + "\n").encode())
    return {k: h.hexdigest() for k, h in out.items()}


def demangle(names):
    cxxfilt = shutil.which("c++filt")
    if not cxxfilt:
        return {n: n for n in names}
    res = subprocess.run([cxxfilt], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, res))


if __name__ == "__main__":
    sys.path.insert(0, ROOT)
    from fdtd_em_solver_b200 import _lib
    hashes = kernel_hashes(_lib.build())
    names = demangle(sorted(hashes))
    if "--write" in sys.argv:
        keep = [a for a in sys.argv[1:] if not a.startswith("--")]          # substrings of kernels to record (default: all)
        rec = {k: dict(kernel=names[k], sha256=v) for k, v in sorted(hashes.items())
               if not keep or any(s in names[k] for s in keep)}
        old = json.load(open(MANIFEST)) if os.path.exists(MANIFEST) else {}
        old.setdefault("kernels", {}).update(rec)
        # kernels that no longer exist in the build (renamed / re-templated / removed) cannot be "unchanged": drop them
        old["kernels"] = {k: v for k, v in old["kernels"].items() if k in hashes}
        old["nvcc_flags"] = _lib.NVCC_FLAGS
        json.dump(old, open(MANIFEST, "w"), indent=1, sort_keys=True)
        print("recorded %d kernels in %s" % (len(rec), MANIFEST))
    else:
        for k, v in sorted(hashes.items()):
            print(v[:16], names[k])
