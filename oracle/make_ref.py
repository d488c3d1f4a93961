"""Make the UNMODIFIED reference travel to the GPU box (test / benchmark infrastructure, not product code).

    python -m oracle.make_ref            # /root/reference/cooking_zoo -> oracle/_ref/cooking_zoo

`/root/reference` exists only in the build container.  The reference is pure Python, so "building" it for the box is a
file copy: every .py and .json of the package (the rendering assets under environment/game/graphics are skipped —
nothing on the step path opens them) goes to the git-ignored `oracle/_ref/`, which gpurun ships with the snapshot like
the repo's own built `.so` files.  Nothing is edited: `MANIFEST.json` records the SHA-256 of every copied file next to
the SHA-256 of its source, and `verify()` re-checks the copy.  Consumers: `bench.py --impl reference` and the
`cpu_baseline` leg (kind "reference"), and the live lockstep tests (`-m reference`), all through
`oracle/ref_loader.py` behind the import stubs of `oracle/refshim/`.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SOURCE = os.environ.get("CZ_REFERENCE_SOURCE", "/root/reference")
KEEP = (".py", ".json")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def make_ref(source=SOURCE, dest=DEST):
    """Copy the package; returns the manifest.  Raises when the source tree is absent."""
    pkg = os.path.join(source, "cooking_zoo")
    if not os.path.isdir(pkg):
        raise RuntimeError(f"reference package not found under {source}")
    out_pkg = os.path.join(dest, "cooking_zoo")
    if os.path.isdir(out_pkg):
        shutil.rmtree(out_pkg)
    files = {}
    for root, _, names in os.walk(pkg):
        for name in sorted(names):
            if not name.endswith(KEEP):
                continue
            src = os.path.join(root, name)
            rel = os.path.relpath(src, source)
            dst = os.path.join(dest, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            files[rel] = {"sha256": _sha(dst), "source_sha256": _sha(src)}
    manifest = {"source": source, "files": files, "skipped": "everything but *.py / *.json (rendering assets)"}
    with open(os.path.join(dest, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return manifest


def verify(dest=DEST):
    """True when every file listed in the manifest is present and byte-identical to what was copied."""
    path = os.path.join(dest, "MANIFEST.json")
    if not os.path.exists(path):
        return False
    manifest = json.load(open(path))
    for rel, h in manifest["files"].items():
        p = os.path.join(dest, rel)
        if not os.path.exists(p) or _sha(p) != h["sha256"] or h["sha256"] != h["source_sha256"]:
            return False
    return bool(manifest["files"])


if __name__ == "__main__":
    m = make_ref()
    print(f"copied {len(m['files'])} files to {DEST}; verify: {verify()}")
    sys.exit(0 if verify() else 1)
