"""Run the reference's UNMODIFIED main.py with the B200 scorer behind its MVS2 interface.

    python -m mvs_b200.launcher [--reference-root DIR] -- -img_p dinoRing/ -par_p dinoRing/dinoR_par.txt -t png -scale 10

The reference does ``from MVS2 import *`` (main.py:5); Python resolves that from sys.modules
first, so registering this package's MVS2 module under that name swaps the MVS stage without
touching or copying any reference file.  matplotlib / mpl_toolkits / pyntcloud (plots and the
PLY writer, absent in this image) are stubbed when missing; SFM.py, BundleAdjustment.py,
GlobalSet.py and utils.py are imported from the reference directory as they are.
"""
import argparse
import os
import runpy
import sys


def install_shim():
    from . import MVS2 as shim
    sys.modules["MVS2"] = shim
    return shim


def stub_missing_viz():
    from unittest.mock import MagicMock
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm", "mpl_toolkits", "mpl_toolkits.mplot3d", "pyntcloud"]:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = MagicMock()


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference-root", default=os.environ.get("MVS_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="arguments of the reference's main.py (after --)")
    a = ap.parse_args(argv)
    rest = a.rest[1:] if a.rest and a.rest[0] == "--" else a.rest
    main_py = os.path.join(a.reference_root, "main.py")
    if not os.path.exists(main_py):
        raise SystemExit(f"{main_py} not found: pass --reference-root")
    stub_missing_viz()
    install_shim()
    sys.path.insert(0, a.reference_root)
    sys.argv = [main_py] + rest
    runpy.run_path(main_py, run_name="__main__")


if __name__ == "__main__":
    main()
