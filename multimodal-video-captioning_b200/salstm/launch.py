"""Run an UNMODIFIED script of the reference (src/train.py, a notebook export, ...) over the B200 modules:

    python /path/to/multimodal-video-captioning_b200/salstm/launch.py /path/to/reference/src/train.py --dataset MSVD --gpu 0

Why a launcher: `python src/train.py` puts the script's own directory (`src/`) at sys.path[0], AHEAD of anything
in PYTHONPATH, so `from models import ...` / `from losses import ...` (train.py:12-13) would silently resolve to the
reference's own CPU/ATen modules.  This launcher executes the script with

    sys.path = [<this package>, <what was there>, <script dir>, <script dir>/..]

so that `models` and `losses` resolve to the B200 implementations, everything else the script imports
(`get_loader`, `pycocoevalcap`, ...) to the reference's own files, and `losses.NLPScore` forwards to the
reference's pycocoevalcap wrapper.  The script itself is not touched.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 2
    script = os.path.abspath(argv[0])
    pkg = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.dirname(script)
    here = os.path.dirname(os.path.abspath(__file__))
    rest = [p for p in sys.path if os.path.abspath(p or ".") not in (pkg, src, here)]
    sys.path[:] = [pkg] + rest + [src, os.path.dirname(src)]
    import models  # noqa: F401  (fail early, and loudly, if the package or libmvc_b200.so is not importable)
    from salstm import cabi
    cabi.load()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
