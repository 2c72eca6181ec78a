"""Recipe for oracle/_ref: the UNMODIFIED reference package, staged for the CPU baseline.

    python oracle/build_ref.py            # build container only (needs /root/reference)

The reference (yliess86/BayeFormers) is a pure-Python package without setup.py / pyproject.toml, so
`pip install --target baseline/_ref /root/reference` has nothing to build (recorded in DESIGN.md); its
"build" is a verbatim copy of the package directory.  The copy goes to oracle/_ref/bayeformers, which is
listed in .gitignore (never part of the history) but not in .gpurunignore, so it travels to the GPU box
where `bench.py --impl reference` and the `cpu_baseline` leg import it: the CPU arm then runs the
reference's own code (`cpu_baseline.kind == "reference"`), not a port.  Nothing in the product package or
in the GPU tests reads it.  Each file's sha256 is written to oracle/_ref/MANIFEST.json so a run can state
exactly which reference it timed.
"""
import hashlib
import json
import os
import shutil
import sys

SRC = "/root/reference/bayeformers"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref", "bayeformers")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"{SRC} not present (GPU box?): keeping whatever oracle/_ref already holds")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    manifest = {}
    for root, _, files in os.walk(DST):
        for f in sorted(files):
            p = os.path.join(root, f)
            manifest[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(HERE, "_ref", "MANIFEST.json"), "w"), indent=1)
    print(f"staged {len(manifest)} files of the reference under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
