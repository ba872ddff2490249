#!/usr/bin/env python3
"""Stage the UNMODIFIED reference hot path into the git-ignored oracle/_ref/ so that it travels to the GPU box
(the box has no /root/reference) and bench.py can time the real Python `PGTGEnv` on the box's host cores.

TEST / BENCH INFRASTRUCTURE ONLY. Copies, byte for byte, the files of the path --
pgtg/{environment,map,parser,map_generator,map_tiles_data,constants}.py -- plus the reference's fixed test map.
Nothing under oracle/_ref/ is tracked by git or imported by the product; the stand-ins for the four absent
third-party imports stay in oracle/shims/.

    python oracle/stage_ref.py            (also run by __graft_entry__.build() when /root/reference exists)
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("PGTG_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["pgtg/environment.py", "pgtg/map.py", "pgtg/parser.py", "pgtg/map_generator.py", "pgtg/map_tiles_data.py", "pgtg/constants.py",
         "tests/test_data/map_with_all_directions.json"]


def stage() -> bool:
    if not os.path.isdir(os.path.join(SRC, "pgtg")):
        return False
    manifest = {}
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump(dict(source=SRC, sha256=manifest), f, indent=1)
    return True


if __name__ == "__main__":
    ok = stage()
    print("staged" if ok else f"{SRC} not present: nothing staged", DST)
    sys.exit(0)
