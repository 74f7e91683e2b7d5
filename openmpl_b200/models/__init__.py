"""Model factories selectable by `config.MODEL` the way the reference's `MPL/lib/models/__init__.py:13` exposes
`multiview_mpl`: `eval('models.' + config.MODEL + '.get_multiview_mpl_net')(config, is_train=...)`."""
from . import multiview_mpl_b200  # noqa: F401
