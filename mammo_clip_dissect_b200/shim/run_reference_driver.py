"""Run an UNMODIFIED reference driver on top of the B200 scoring path:

    python -m mammo_clip_dissect_b200.shim.run_reference_driver /path/to/concept_vit/describe_clip_neurons.py [driver args]

The shim directory goes first on sys.path (so `import similarity` and `import CLIP_og_utils` pick up the
drop-ins); runpy.run_path does not prepend the script's own directory.  Extra directories listed in
MCD_EXTRA_PATH (os.pathsep separated) are inserted right after the shim -- tests use it for stub `clip` /
`data_utils` modules when the real models and datasets are not available.
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script, rest = argv[0], argv[1:]
    here = os.path.dirname(os.path.abspath(__file__))
    os.environ.setdefault("MCD_REFERENCE_DIR", os.path.dirname(os.path.abspath(script)))
    extra = [p for p in os.environ.get("MCD_EXTRA_PATH", "").split(os.pathsep) if p]
    sys.path[:0] = [here] + extra
    sys.argv = [script] + rest
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
