#!/usr/bin/env python
"""Stage an UNMODIFIED copy of the reference under baseline/_ref (git-ignored, not gpurun-ignored) so that the
drop-in tests and the CPU reference arm can run the reference's own files on the GPU box, where /root/reference
does not exist.  Nothing from the reference enters the repo's history: only the sha256 manifest
(tests/golden/reference_manifest.json) is committed, and the tests check the staged files against it.

    python tools/stage_reference.py            # /root/reference -> baseline/_ref
"""
import hashlib
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CROWDNAV_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(REPO, "baseline", "_ref")
SKIP_DIRS = {".git", "figures", "__pycache__", "ORCA_no_rand", "SF_no_rand", "my_model", "datasets"}
SKIP_EXT = {".png", ".gif", ".jpg", ".mp4", ".pyc"}


def main():
    if not os.path.isdir(SRC):
        print("no reference at", SRC)
        return 1
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    manifest = {}
    for root, dirs, files in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if d not in SKIP_DIRS)
        rel = os.path.relpath(root, SRC)
        for f in sorted(files):
            if os.path.splitext(f)[1].lower() in SKIP_EXT:
                continue
            s = os.path.join(root, f)
            d = os.path.join(DST, rel, f)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            os.chmod(d, 0o644)
            manifest[os.path.normpath(os.path.join(rel, f))] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=0, sort_keys=True)
    gold = os.path.join(REPO, "tests", "golden", "reference_manifest.json")
    json.dump(manifest, open(gold, "w"), indent=0, sort_keys=True)
    total = sum(os.path.getsize(os.path.join(DST, k)) for k in manifest)
    print("staged %d files, %.1f MB -> %s; manifest -> %s" % (len(manifest), total / 2 ** 20, DST, gold))
    return 0


if __name__ == "__main__":
    sys.exit(main())
