"""Copies the reference's own pytest files, byte for byte, into tests/golden/ref_tests/ so that they run - unmodified -
against this repository's `rogue_gym` package on the GPU box (where /root/reference does not exist).

Run in the build container only:
    python tests/golden/fetch_ref_tests.py
Source: kngwyu/rogue-gym @ c78608b, python/tests/{data,test_ff_env,test_st_env,test_rogue_env,test_parallel}.py.
MANIFEST.json records the sha256 of every file as copied; tests/golden/ref_tests/conftest.py (ours) supplies what the
files expect from their environment: the `gym` module (a stand-in, neither gym nor gymnasium is installed), `data`
on the import path, the `gpu` marker, and the list of cases whose golden data is stale in the reference itself."""
import hashlib
import json
import os
import shutil

SRC = "/root/reference/python/tests"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_tests")
FILES = ["data.py", "test_ff_env.py", "test_st_env.py", "test_rogue_env.py", "test_parallel.py"]

os.makedirs(DST, exist_ok=True)
manifest = {"source": "kngwyu/rogue-gym @ c78608b python/tests/", "files": {}}
for f in FILES:
    shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    manifest["files"][f] = hashlib.sha256(open(os.path.join(DST, f), "rb").read()).hexdigest()
with open(os.path.join(DST, "MANIFEST.json"), "w") as fh:
    json.dump(manifest, fh, indent=1)
print("copied", ", ".join(FILES))
