"""Recipe that BUILDS the reference for the reference arm of bench.py (``--impl reference``), from the sources where they
lie under /root/reference, into ``oracle/_ref/`` (git-ignored build output; it travels to the GPU box with the snapshot,
like the repo's own built ``.so`` files - /root/reference itself does not exist there).

The reference's hot path is Python: "building" it is byte-compiling ``src/**/*.py`` into sourceless ``.pyc`` modules
(``py_compile``, legacy layout so that the import system loads them without sources).  No reference source is copied
into the repository and nothing of it is modified; bench.py imports the resulting package and drives the reference's own
``MapleCLIPSeg`` / ``COOPCRIS`` classes on the host cores.

    python oracle/build_ref.py          (``__graft_entry__.build()`` runs it when /root/reference is present)
"""
from __future__ import annotations

import os
import py_compile
import shutil
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def build(verbose: bool = False) -> str | None:
    src_root = os.path.join(REF, "src")
    if not os.path.isdir(src_root):
        return None
    if os.path.isdir(OUT):
        shutil.rmtree(OUT)
    n = 0
    for d, _, files in os.walk(src_root):
        for f in files:
            if not f.endswith(".py"):
                continue
            rel = os.path.relpath(os.path.join(d, f), REF)
            dst = os.path.join(OUT, rel + "c")               # pkg/mod.py -> pkg/mod.pyc (sourceless import layout)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            py_compile.compile(os.path.join(d, f), cfile=dst, dfile=os.path.join("<reference>", rel), doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
            n += 1
    # the same modules once more as ONE archive (zipimport loads sourceless .pyc members): file-sync tools that drop *.pyc on the
    # way to the GPU box (the round-1 / round-2 reference arm silently fell back to the oracle port there) leave a .zip alone
    import zipfile

    with zipfile.ZipFile(os.path.join(OUT, "reference_src.zip"), "w", zipfile.ZIP_STORED) as z:
        for d, _, files in os.walk(os.path.join(OUT, "src")):
            z.writestr(os.path.relpath(d, OUT) + "/", b"")      # explicit directory entries: zipimport needs them for packages without __init__
            for f in files:
                if f.endswith(".pyc"):
                    full = os.path.join(d, f)
                    z.write(full, os.path.relpath(full, OUT))
    with open(os.path.join(OUT, "BUILT_FROM"), "w") as fh:
        fh.write(f"{REF}/src ({n} modules, python {sys.version.split()[0]})\n")
    if verbose:
        print(f"[build_ref] {n} modules -> {OUT}")
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
