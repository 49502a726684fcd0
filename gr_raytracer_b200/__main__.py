"""`python -m gr_raytracer_b200 <reference command line>`: forwards to the C++ CLI (`grrt`), which mirrors
reference src/cli/cli.rs.  Texture paths in the bundled scenes are relative to the repository root."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    exe = os.path.join(HERE, "grrt")
    if not os.path.exists(exe):
        from . import build
        build.build()
    args = sys.argv[1:]
    if not any(a.startswith("--resource-root") for a in args):
        args = ["--resource-root", os.path.dirname(HERE)] + args
    sys.exit(subprocess.call([exe] + args))


if __name__ == "__main__":
    main()
