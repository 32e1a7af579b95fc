"""`util` package of the drop-in.  When the reference tree follows lw-detr_b200/ on sys.path (the INTEGRATION.md set-up:
PYTHONPATH=<repo>/lw-detr_b200:<reference>), the reference's own util/ directory is appended to this package's search
path, so its scripts (`demo/demo.py`, `main.py`) still find `util.get_param_dicts`, `util.box_ops`, ... next to the
`models` package they now get from here."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
