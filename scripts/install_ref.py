"""Installs the UNMODIFIED upstream reference into the git-ignored ``baseline/_ref/`` so that ``bench.py --impl reference``
can time the reference's own ``UnitSpeech.forward`` (unitspeech/unitspeech.py:387-391) on the GPU box's host cores.

``/root/reference`` exists only in the build container; ``baseline/_ref/`` is git-ignored but NOT gpurun-ignored, so it
travels to the GPU box with the snapshot (like the built .so).  Nothing under ``baseline/_ref`` is tracked, edited or imported
by the product: only ``bench.py``'s CPU legs and the oracle-validation tests load it, through ``oracle/ref_shim.py``.

Recipe:
  1. the documented ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of the reference>``
     is attempted first (from a copy under /tmp: the source tree is read-only).  The reference's setup.py declares
     ``py_modules=["unitspeech"]`` although ``unitspeech`` is a directory without ``__init__.py``, so the build either fails
     or installs no module; the outcome is written to ``baseline/_ref/INSTALL.json``.
  2. if ``unitspeech/unitspeech.py`` did not land, the 11 files of the decoder/vocoder path are
     copied verbatim, keeping the reference's directory shape (``unitspeech`` is a namespace package).

Run from ``__graft_entry__.build()`` whenever ``/root/reference`` is present; a no-op otherwise.
"""

from __future__ import annotations

import filecmp
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = os.environ.get("UNITSPEECH_REFERENCE_SRC", "/root/reference")
DEST = os.path.join(ROOT, "baseline", "_ref")

# what `import unitspeech.unitspeech` / `unitspeech.vocoder.models` need (everything else those
# modules import at top level -- librosa, matplotlib, phonemizer, conf, speaker/unit encoders -- is stubbed by ref_shim)
FILES = [
    "unitspeech/unitspeech.py",
    "unitspeech/base.py",
    "unitspeech/util.py",
    "unitspeech/vocoder/models.py",
    "unitspeech/vocoder/activations.py",
    "unitspeech/vocoder/env.py",
    "unitspeech/vocoder/xutils.py",
    "unitspeech/vocoder/alias_free_torch/__init__.py",
    "unitspeech/vocoder/alias_free_torch/act.py",
    "unitspeech/vocoder/alias_free_torch/filter.py",
    "unitspeech/vocoder/alias_free_torch/resample.py",
]


def _installed() -> bool:
    return all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)


def _try_pip() -> dict:
    tmp = tempfile.mkdtemp(prefix="usb_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns("*.wav", "*.npy", "*.ipynb", "logs", "evaluation", "notebooks",
                                                                     "resources", ".git", "DUMMY"),
                        symlinks=True, ignore_dangling_symlinks=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", DEST, src]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        return {"cmd": " ".join(cmd[2:]), "rc": r.returncode, "tail": (r.stdout + r.stderr)[-600:]}
    except Exception as exc:  # noqa: BLE001
        return {"cmd": "pip install", "rc": -1, "tail": repr(exc)}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def install(verbose: bool = True, try_pip: bool = True) -> bool:
    """Returns True when baseline/_ref holds the reference's decoder path."""
    if not os.path.isfile(os.path.join(REF_SRC, "unitspeech", "unitspeech.py")):
        if verbose:
            print(f"install_ref: {REF_SRC} not present; keeping baseline/_ref as it is "
                  f"({'installed' if _installed() else 'absent'})", file=sys.stderr)
        return _installed()
    up_to_date = _installed() and all(filecmp.cmp(os.path.join(REF_SRC, f), os.path.join(DEST, f), shallow=False) for f in FILES)
    if up_to_date and os.path.isfile(os.path.join(DEST, "INSTALL.json")):
        return True
    os.makedirs(DEST, exist_ok=True)
    record = {"source": REF_SRC, "pip": _try_pip() if try_pip else None}
    pip_landed = os.path.isfile(os.path.join(DEST, "unitspeech", "unitspeech.py"))
    record["pip_installed_decoder"] = pip_landed
    if not pip_landed:
        for f in FILES:
            dst = os.path.join(DEST, f)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(REF_SRC, f), dst)     # verbatim, unmodified
        record["copied_files"] = FILES
    with open(os.path.join(DEST, "INSTALL.json"), "w") as f:
        json.dump(record, f, indent=1)
    if verbose:
        print(f"install_ref: reference decoder path installed under {DEST} "
              f"({'pip' if pip_landed else 'verbatim copy of %d files' % len(FILES)}; pip rc {record['pip']['rc'] if record['pip'] else 'n/a'})",
              file=sys.stderr)
    return _installed()


if __name__ == "__main__":
    ok = install()
    sys.exit(0 if ok else 1)
