#!/usr/bin/env python3
"""Check that a source change left the device code of existing kernels untouched.

    python tools/sass_identity.py <git-rev>      # e.g. the last commit whose kernels ran on a GPU

Builds csrc/ of <git-rev> into a scratch directory with the flags of build.py, disassembles both
libraries with cuobjdump and compares every kernel's instruction stream (encodings are dropped, so
line-info changes do not count; a trailing defaulted `false` template argument is ignored so that a
kernel may gain an opt-in template flag).  Exit status 1 if any kernel of <git-rev> changed or vanished.
Used when there is no GPU at hand: host-side refactors and new opt-in instantiations can then be
committed without touching what was validated.
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "boltzmann-fourier-spectral-method_b200"
sys.path.insert(0, ROOT)


def kernels(so):
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    d = {}
    for f in out.split("Function : ")[1:]:
        name = f.split("\n")[0].strip()
        body = [re.sub(r"/\* 0x[0-9a-f]* \*/", "", l).rstrip() for l in f.split("\n")
                if re.match(r"\s*/\*[0-9a-f]{4,5}\*/", l)]
        name = re.sub(r"_GLOBAL__N__[0-9a-f]+_\d+_(\w+?)_cu_[0-9a-f]{8}", r"_GLOBAL__N_\1_cu", name)  # file hash
        name = re.sub(r"ELb0EEEv", "EEEv", name)               # trailing defaulted `false`
        name = re.sub(r"(k_plane_gain_wsILi\d+ELi\d+)ELi0EEEv", r"\1EEEv", name)  # ... or defaulted RCFG = 0
        d[name] = body
    return d


def main(rev):
    import bfsm_b200
    build = bfsm_b200.submodule("build")
    cur = build.build_library()
    with tempfile.TemporaryDirectory() as tmp:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, PKG + "/csrc", "include"],
                             capture_output=True, check=True).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
        ref = os.path.join(tmp, "ref.so")
        subprocess.run([build._nvcc()] + build.NVCC_FLAGS + ["-o", ref, "bfsm_capi.cu"],
                       cwd=os.path.join(tmp, PKG, "csrc"), check=True, capture_output=True)
        old, new = kernels(ref), kernels(cur)
    changed = [k for k, v in old.items() if new.get(k) != v]
    for k in changed:
        print("CHANGED" if k in new else "MISSING", k[:110])
    print(f"{len(old) - len(changed)} of {len(old)} kernels of {rev} are instruction-identical; "
          f"{len(set(new) - set(old))} new kernels")
    return 1 if changed else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "HEAD"))
