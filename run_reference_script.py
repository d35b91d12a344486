"""Run one of the reference's UNMODIFIED training scripts against the drop-in modules of this repo:

    python run_reference_script.py /path/to/reference/train_small_graphs.py --model GCN_DGG_00 --data cora ...

``python /path/to/reference/train_x.py`` would put the script's own directory first on ``sys.path`` and import the
reference's ``model.py`` / ``dgm.py``; this launcher puts THIS repo first (its ``model`` / ``dgm`` / ``utils`` shadow
the reference's, ``utils`` re-exporting everything else from the reference's own file) and the script's directory
second.  Equivalent on Python >= 3.11:  PYTHONSAFEPATH=1 PYTHONPATH=<repo>:<reference> python <script> ...
"""
import os
import runpy
import sys


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    script = os.path.abspath(sys.argv[1])
    repo = os.path.dirname(os.path.abspath(__file__))
    ref = os.path.dirname(script)
    sys.path[:] = [repo, ref] + [p for p in sys.path if p and os.path.abspath(p) not in (repo, ref)]
    if os.environ.get("DGGB_PYG_STUB"):      # build container only: torch_geometric is not installable there
        from oracle import ref_loader

        ref_loader.install_pyg_stub()
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
